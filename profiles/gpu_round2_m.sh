#!/bin/bash
mkdir -p gpurun_out
tag=${1:-m}
V=evergreen_b200/variants
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=40 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" > gpurun_out/env_$tag.txt
: > gpurun_out/diag_$tag.txt
timeout 300 python profiles/diag_c2.py 2 16 >> gpurun_out/diag_$tag.txt 2>&1
timeout 900 python profiles/ab_variants.py run 200 > gpurun_out/ab_$tag.txt 2>&1
: > gpurun_out/gen_$tag.txt
for w in c5 c3; do timeout 300 python profiles/prof_general.py $w 20 >> gpurun_out/gen_$tag.txt 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5_$tag.csv python profiles/prof_general.py c5 3 > gpurun_out/ncu_l3.log 2>&1
cat gpurun_out/env_$tag.txt; tail -4 gpurun_out/pytest_$tag.log; grep "bad ticks" gpurun_out/diag_$tag.txt; cat gpurun_out/ab_$tag.txt; grep ok gpurun_out/gen_$tag.txt
