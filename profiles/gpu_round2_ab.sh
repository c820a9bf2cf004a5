#!/bin/bash
mkdir -p gpurun_out
timeout 900 python profiles/diag_c2.py 2 80 > gpurun_out/soak_c2_80.txt 2>&1
tail -2 gpurun_out/soak_c2_80.txt
