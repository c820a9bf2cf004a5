#!/bin/bash
# correctness first (whole GPU suite, race diagnostic on every variant), then A/B timings, then profiles
mkdir -p gpurun_out
tag=${1:-f}
V=evergreen_b200/variants
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" > gpurun_out/env_$tag.txt
: > gpurun_out/diag_$tag.txt
timeout 300 python profiles/diag_c2.py 2 10 >> gpurun_out/diag_$tag.txt 2>&1
for v in late pf2 hint pf2hint; do timeout 300 python profiles/diag_c2.py 2 10 $V/$v.so >> gpurun_out/diag_$tag.txt 2>&1; done
timeout 900 python profiles/ab_variants.py run 200 > gpurun_out/ab_$tag.txt 2>&1
: > gpurun_out/gen_$tag.txt
for w in c3 plain 1m; do
  timeout 300 python profiles/prof_general.py $w 20 >> gpurun_out/gen_$tag.txt 2>&1
  timeout 300 python profiles/prof_general.py $w 20 $V/occ4.so >> gpurun_out/gen_$tag.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c3_$tag.csv python profiles/prof_general.py c3 3 > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_plain_$tag.csv python profiles/prof_general.py plain 3 > gpurun_out/ncu_l2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_plan_cta -s 2 -c 1 -f -o gpurun_out/r02_prof_cta_$tag python profiles/prof_cta.py 3 > gpurun_out/ncu_f1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_gtask|k_gscatter|k_gbest|k_gunit|k_gfill" -s 10 -c 7 -f -o gpurun_out/r02_prof_gen_$tag python profiles/prof_general.py c3 3 > gpurun_out/ncu_f2.log 2>&1
cat gpurun_out/env_$tag.txt; tail -5 gpurun_out/pytest_$tag.log; grep "bad ticks" gpurun_out/diag_$tag.txt; cat gpurun_out/ab_$tag.txt; grep ok gpurun_out/gen_$tag.txt
