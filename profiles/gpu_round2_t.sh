#!/bin/bash
# soak: repeated resident ticks of general-path / mixed workloads, every tick against the oracle (order-free atomics in
# the unit table must not reach the output), plus racecheck on a small mixed tick
mkdir -p gpurun_out
tag=${1:-t}
: > gpurun_out/soak_$tag.txt
timeout 600 python profiles/diag_c2.py c3 12 >> gpurun_out/soak_$tag.txt 2>&1
timeout 600 python profiles/diag_c2.py mixed 20 >> gpurun_out/soak_$tag.txt 2>&1
timeout 600 python profiles/diag_c2.py c5s 12 >> gpurun_out/soak_$tag.txt 2>&1
timeout 600 python profiles/diag_c2.py 4 12 >> gpurun_out/soak_$tag.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cta_class_boundaries or general_path_shapes or sparse_classes" > gpurun_out/racecheck_$tag.log 2>&1; echo "racecheck rc=$?" > gpurun_out/env_$tag.txt
cat gpurun_out/env_$tag.txt; grep "bad ticks" gpurun_out/soak_$tag.txt; tail -6 gpurun_out/racecheck_$tag.log
