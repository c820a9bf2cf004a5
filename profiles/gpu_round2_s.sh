#!/bin/bash
mkdir -p gpurun_out
tag=${1:-s}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "finder" > gpurun_out/pytest_pf_$tag.log 2>&1; echo "pf rc=$?" > gpurun_out/env_$tag.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "plan_from_finder or finders_device_output or update_tasks" > gpurun_out/sanitizer_pf_$tag.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/env_$tag.txt
cat gpurun_out/env_$tag.txt; tail -30 gpurun_out/pytest_pf_$tag.log; tail -5 gpurun_out/sanitizer_pf_$tag.log
