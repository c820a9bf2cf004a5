#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_legacy.py -m gpu -q > gpurun_out/pytest_ad.log 2>&1; echo "pytest rc=$?" > gpurun_out/env_ad.txt
timeout 900 python profiles/prof_rows_r02.py > gpurun_out/rows_r02.json 2> gpurun_out/rows_r02.err
cat gpurun_out/env_ad.txt; tail -3 gpurun_out/pytest_ad.log; tail -2 gpurun_out/rows_r02.err; cat gpurun_out/rows_r02.json
