#!/bin/bash
mkdir -p gpurun_out
tag=${1:-q}
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=40 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" > gpurun_out/env_$tag.txt
: > gpurun_out/gen_$tag.txt
for w in c4 c3t c5 c2; do timeout 300 python profiles/prof_general.py $w 20 >> gpurun_out/gen_$tag.txt 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c4_$tag.csv python profiles/prof_general.py c4 3 > gpurun_out/ncu_l4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c3t_$tag.csv python profiles/prof_general.py c3t 3 > gpurun_out/ncu_l5.log 2>&1
cat gpurun_out/env_$tag.txt; tail -4 gpurun_out/pytest_$tag.log; grep ok gpurun_out/gen_$tag.txt
