#!/bin/bash
# GPU call 2: legacy tests, bench (both arms), launch lists and ncu captures of the two second-generation planners.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_legacy.py -m gpu -q --maxfail=10 > gpurun_out/pytest_legacy.log 2>&1; echo "pytest_legacy rc=$?" > gpurun_out/env2.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/env2.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> gpurun_out/env2.txt
timeout 300 python profiles/prof_cta.py 6 > gpurun_out/prof_cta.txt 2>&1
timeout 300 python profiles/prof_general.py c3 4 > gpurun_out/prof_gen_c3.txt 2>&1
timeout 300 python profiles/prof_general.py 1m 4 > gpurun_out/prof_gen_1m.txt 2>&1
timeout 300 python profiles/prof_general.py plain 4 > gpurun_out/prof_gen_plain.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3.csv python profiles/prof_general.py c3 3 > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cta.csv python profiles/prof_cta.py 3 > gpurun_out/ncu_l2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_plan_cta -s 2 -c 1 -f -o gpurun_out/r02_prof_cta python profiles/prof_cta.py 3 > gpurun_out/ncu_f1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_gtask|k_gscatter|k_gplace|k_gcomplex|k_ghist" -s 14 -c 8 -f -o gpurun_out/r02_prof_gen python profiles/prof_general.py c3 2 > gpurun_out/ncu_f2.log 2>&1
cat gpurun_out/env2.txt; tail -2 gpurun_out/pytest_legacy.log; cat gpurun_out/prof_cta.txt gpurun_out/prof_gen_*.txt | grep ok; head -c 3000 gpurun_out/bench.json
