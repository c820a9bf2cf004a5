#!/bin/bash
mkdir -p gpurun_out
s=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; echo "bench rc=$? wall=$(( $(date +%s) - s ))s" > gpurun_out/env_w.txt
s=$(date +%s); timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_w.json 2> gpurun_out/bench_ref_w.err; echo "ref rc=$? wall=$(( $(date +%s) - s ))s" >> gpurun_out/env_w.txt
cat gpurun_out/env_w.txt; head -c 300 gpurun_out/bench_w.json; echo; head -c 300 gpurun_out/bench_ref_w.json
