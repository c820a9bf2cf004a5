#!/bin/bash
mkdir -p gpurun_out
tag=${1:-u}
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=40 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" > gpurun_out/env_$tag.txt
: > gpurun_out/gen_$tag.txt
for w in c3 1m c5; do timeout 300 python profiles/prof_general.py $w 20 >> gpurun_out/gen_$tag.txt 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c3_$tag.csv python profiles/prof_general.py c3 3 > gpurun_out/ncu_l1.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-shapes --no-delta --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cat gpurun_out/env_$tag.txt; tail -4 gpurun_out/pytest_$tag.log; grep ok gpurun_out/gen_$tag.txt; cat gpurun_out/bench_$tag.json | cut -c1-300
