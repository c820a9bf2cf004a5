#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_ae.log 2>&1; echo "smoke rc=$?" > gpurun_out/env_ae.txt
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/pytest_ae.log 2>&1; echo "pytest rc=$?" >> gpurun_out/env_ae.txt
cat gpurun_out/env_ae.txt; tail -2 gpurun_out/smoke_ae.log; tail -3 gpurun_out/pytest_ae.log
