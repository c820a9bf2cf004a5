#!/bin/bash
# Round-2 evidence run: environment, smoke, the bench (both arms), ncu launch list + one full capture of the roofline
# kernel on the bench workload, sanitizer over the tests that touch this round's kernels.
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,driver_version --format=csv; nproc; cat /sys/fs/cgroup/cpu.max; cat /sys/fs/cgroup/memory.max; free -g | head -2; (go version || echo "no go toolchain") 2>&1; nvcc --version | tail -2; } > gpurun_out/r02_box_env.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_box_env.txt
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?" >> gpurun_out/r02_box_env.txt
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "bench_ref rc=$?" >> gpurun_out/r02_box_env.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 2 --warmup 1 --distros 400 --no-shapes --no-cpu-baseline --no-delta --e2e-steps 1 > gpurun_out/ncu_b1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_gtask" -s 2 -c 1 -f -o gpurun_out/r02_prof_gtask_bench python bench.py --steps 2 --warmup 1 --distros 400 --no-shapes --no-cpu-baseline --no-delta --e2e-steps 1 > gpurun_out/ncu_b2.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_legacy.py -m gpu -q -x -k "general_path_shapes or sparse_classes or dependency_fan or group_versions_big or cta_class_boundaries or update_tasks or mixed_parity or breakdown or legacy or dag or persist" > gpurun_out/sanitizer_final.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/r02_box_env.txt
cat gpurun_out/r02_box_env.txt; tail -3 gpurun_out/sanitizer_final.log; head -c 600 gpurun_out/bench_final.json; echo; head -c 400 gpurun_out/bench_ref_final.json
