#!/bin/bash
# iteration call: A/B of library variants, planners' ncu captures, the GPU tests that touch the planners
mkdir -p gpurun_out
tag=${1:-d}
timeout 900 python profiles/ab_variants.py run 200 > gpurun_out/ab_$tag.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=30 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" > gpurun_out/env_$tag.txt
for w in c3 plain 1m; do timeout 300 python profiles/prof_general.py $w 4 > gpurun_out/prof_gen_${w}_$tag.txt 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3_$tag.csv python profiles/prof_general.py c3 3 > gpurun_out/ncu_l1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_plan_cta -s 2 -c 1 -f -o gpurun_out/r02_prof_cta_$tag python profiles/prof_cta.py 3 > gpurun_out/ncu_f1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_gtask|k_glink|k_gunit|k_gbest" -s 8 -c 4 -f -o gpurun_out/r02_prof_gen_$tag python profiles/prof_general.py c3 3 > gpurun_out/ncu_f2.log 2>&1
cat gpurun_out/ab_$tag.txt gpurun_out/env_$tag.txt; tail -3 gpurun_out/pytest_$tag.log; cat gpurun_out/prof_gen_*_$tag.txt | grep ok
