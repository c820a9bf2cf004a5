#!/bin/bash
mkdir -p gpurun_out
tag=${1:-o}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" > gpurun_out/env_$tag.txt
: > gpurun_out/gen_$tag.txt
for w in c5 c3 plain 1m c2; do timeout 300 python profiles/prof_general.py $w 20 >> gpurun_out/gen_$tag.txt 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5_$tag.csv python profiles/prof_general.py c5 3 > gpurun_out/ncu_l3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_alloc" -s 2 -c 2 -f -o gpurun_out/r02_prof_alloc_$tag python profiles/prof_general.py c5 3 > gpurun_out/ncu_f3.log 2>&1
cat gpurun_out/env_$tag.txt; tail -4 gpurun_out/pytest_$tag.log; grep ok gpurun_out/gen_$tag.txt
