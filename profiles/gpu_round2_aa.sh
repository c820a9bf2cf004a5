#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/gen_aa.txt
for w in c3 1m c5; do
  timeout 300 python profiles/prof_general.py $w 30 >> gpurun_out/gen_aa.txt 2>&1
  timeout 300 python profiles/prof_general.py $w 30 evergreen_b200/variants/wl6.so >> gpurun_out/gen_aa.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3_aa.csv python profiles/prof_general.py c3 3 evergreen_b200/variants/wl6.so > gpurun_out/ncu_l1.log 2>&1
grep -a "^ok" gpurun_out/gen_aa.txt
