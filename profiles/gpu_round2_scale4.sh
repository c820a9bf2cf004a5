#!/bin/bash
mkdir -p gpurun_out
bash profiles/scale.sh "4" > gpurun_out/scale4_r02.txt 2>&1
cat gpurun_out/scale4_r02.txt; tail -3 gpurun_out/scale_4.err
