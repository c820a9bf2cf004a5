#!/bin/bash
# One gpurun call: environment facts, smoke, the GPU test suite in two processes (new paths first), the bench.
# Every leg runs under its own timeout and logs into gpurun_out/, so one hang or failure does not hide the rest.
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; nproc; cat /sys/fs/cgroup/cpu.max; cat /sys/fs/cgroup/memory.max; free -g | head -2; (go version || echo "no go") 2>&1; } > gpurun_out/env.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/env.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -k "cta or general_path or million or upload_device or contexts or config3_each" > gpurun_out/pytest_new.log 2>&1; echo "pytest_new rc=$?" >> gpurun_out/env.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 -k "not (cta or general_path or million or upload_device or contexts or config3_each)" > gpurun_out/pytest_rest.log 2>&1; echo "pytest_rest rc=$?" >> gpurun_out/env.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/env.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> gpurun_out/env.txt
tail -3 gpurun_out/pytest_new.log; tail -3 gpurun_out/pytest_rest.log; cat gpurun_out/env.txt; head -c 1500 gpurun_out/bench.json
