#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/diag_c2.txt
timeout 300 python profiles/diag_c2.py 2 12 >> gpurun_out/diag_c2.txt 2>&1
for v in nofull late tpt1; do timeout 300 python profiles/diag_c2.py 2 12 evergreen_b200/variants/$v.so >> gpurun_out/diag_c2.txt 2>&1; done
cat gpurun_out/diag_c2.txt
