#!/bin/bash
mkdir -p gpurun_out
tag=${1:-zz}
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?" > gpurun_out/env_$tag.txt
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/env_$tag.txt
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?" >> gpurun_out/env_$tag.txt
cat gpurun_out/env_$tag.txt; tail -3 gpurun_out/smoke_$tag.log; tail -4 gpurun_out/pytest_$tag.log; grep -a "Elapsed (wall" gpurun_out/bench_$tag.err; head -c 400 gpurun_out/bench_$tag.json
