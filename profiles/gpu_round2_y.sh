#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/gen_y.txt
for w in c3t c5 c4; do
  timeout 300 python profiles/prof_general.py $w 40 >> gpurun_out/gen_y.txt 2>&1
  timeout 300 python profiles/prof_general.py $w 40 evergreen_b200/variants/wocc4.so >> gpurun_out/gen_y.txt 2>&1
done
grep -a "^ok" gpurun_out/gen_y.txt
