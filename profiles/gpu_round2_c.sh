#!/bin/bash
# GPU call: whole GPU test suite, then timings / launch lists / ncu captures of the planners, then the bench.
mkdir -p gpurun_out
tag=${1:-c}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" > gpurun_out/env_$tag.txt
timeout 900 python profiles/ab_variants.py run 200 > gpurun_out/ab_$tag.txt 2>&1
timeout 300 python profiles/prof_cta.py 6 > gpurun_out/prof_cta_$tag.txt 2>&1
for w in c3 1m plain; do timeout 300 python profiles/prof_general.py $w 4 > gpurun_out/prof_gen_${w}_$tag.txt 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3_$tag.csv python profiles/prof_general.py c3 3 > gpurun_out/ncu_l1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_plan_cta -s 2 -c 1 -f -o gpurun_out/r02_prof_cta_$tag python profiles/prof_cta.py 3 > gpurun_out/ncu_f1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_gtask|k_gbest|k_gunit|k_gscatter" -s 4 -c 6 -f -o gpurun_out/r02_prof_gen_$tag python profiles/prof_general.py c3 2 > gpurun_out/ncu_f2.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?" >> gpurun_out/env_$tag.txt
cat gpurun_out/ab_$tag.txt; cat gpurun_out/env_$tag.txt; tail -3 gpurun_out/pytest_$tag.log; cat gpurun_out/prof_cta_$tag.txt gpurun_out/prof_gen_*_$tag.txt | grep ok; head -c 1200 gpurun_out/bench_$tag.json
