#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/gen_x.txt
for w in c3 plain 1m; do
  timeout 300 python profiles/prof_general.py $w 30 >> gpurun_out/gen_x.txt 2>&1
  timeout 300 python profiles/prof_general.py $w 30 evergreen_b200/variants/gocc2.so >> gpurun_out/gen_x.txt 2>&1
done
grep -a "^ok" gpurun_out/gen_x.txt
