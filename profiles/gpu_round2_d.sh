#!/bin/bash
# quick iteration call: A/B of library variants, the on-chip planner's ncu capture, targeted tests
mkdir -p gpurun_out
tag=${1:-d}
timeout 900 python profiles/ab_variants.py run 200 > gpurun_out/ab_$tag.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -k "cta or persist or device_side or c2_full or config_parity or mixed or route or all_routes or dag" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" > gpurun_out/env_$tag.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_plan_cta -s 2 -c 1 -f -o gpurun_out/r02_prof_cta_$tag python profiles/prof_cta.py 3 > gpurun_out/ncu_f1.log 2>&1
for w in c3 plain; do timeout 300 python profiles/prof_general.py $w 4 > gpurun_out/prof_gen_${w}_$tag.txt 2>&1; done
cat gpurun_out/ab_$tag.txt gpurun_out/env_$tag.txt; tail -3 gpurun_out/pytest_$tag.log; cat gpurun_out/prof_gen_*_$tag.txt | grep ok
