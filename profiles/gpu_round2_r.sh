#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "finders_device_output" > gpurun_out/pytest_pf_$tag.log 2>&1; echo "pf rc=$?" > gpurun_out/env_$tag.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/env_$tag.txt
cat gpurun_out/env_$tag.txt; tail -30 gpurun_out/pytest_pf_$tag.log; tail -4 gpurun_out/pytest_$tag.log
