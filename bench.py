#!/usr/bin/env python3
"""bench.py -- the scheduler hot path on synthetic ticks of BASELINE.json's shape.

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU algorithm on the host cores)

A "step" is one scheduler tick: tunable planner + DistroQueueInfo + utilization
host allocator over every distro of the workload (configs[1]: 1k distros x 10k
tasks each per GPU).  `value` = tasks ranked per second with the tick's inputs
already resident in HBM; `e2e` = the same through the public API
(Engine.plan_and_alloc_batch: pinned HOST buffers in, host buffers out, H2D and
D2H inside the timed region).  N > 1: one process per GPU (torchrun), distros
sharded whole by LPT, one NCCL all-gather of the per-distro result vector per
step, max-over-ranks timing; weak scaling (each GPU owns its own 1k distros).

The reference is Go and this image has no Go toolchain, so the reference arm
times oracle/evg_oracle.cpp (a C++ restatement of the reference's algorithm,
"port") on all host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tasks scheduled/sec over N distros; host-allocator decisions/sec at 1/2/4/8 GPU"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def workload(rank: int, world: int, distros_per_gpu: int, tasks_per_distro: int):
    """configs[1] per GPU.  All N*distros_per_gpu distros form one tick; LPT assigns whole distros."""
    from evergreen_b200 import dist, synth
    D = world * distros_per_gpu
    sizes = np.full(D, tasks_per_distro, dtype=np.int64)
    shards = dist.lpt_partition(sizes, world)
    mine = shards.members[rank]
    w = synth.make(sizes[mine], synth.SEED_BASE + 2 + 1000 * rank,
                   name=f"C2: {len(mine)} distros x {tasks_per_distro} tasks each (rank {rank}/{world})",
                   n_hosts=5 * len(mine))
    return w, shards


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = device_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(device_index), "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def bind_to_gpu_node(torch, local_rank: int):
    """Run this rank on the CPUs next to its GPU (NVML's affinity mask for the device with the same PCI bus id), so
    pinned host buffers are first-touched on the NUMA node the GPU's PCIe link hangs off.  Returns the previous
    affinity set (to restore for the host-core baseline leg) or None when anything about it is unavailable."""
    try:
        import pynvml
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        n = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = {i for i in range(n) if (int(mask[i // 64]) >> (i % 64)) & 1} & prev
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return prev
    except Exception:
        return None


def usable_cores():
    """Host cores this process may actually burn: the affinity mask capped by the cgroup CPU quota (the GPU boxes
    show 128 logical CPUs under a 16-CPU quota; 128 runnable threads there only add throttling)."""
    import math
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    quota = None
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]  # cgroup v2
        if q != "max":
            quota = float(q) / float(p)
    except Exception:
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())  # cgroup v1
            p = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / p
        except Exception:
            pass
    if quota:
        n = max(1, min(n, int(math.ceil(quota))))
    return n, (os.cpu_count() or n), quota


def cores_note():
    n, logical, quota = usable_cores()
    return f"{n} threads = usable host cores ({logical} logical CPUs" + (f", cgroup CPU quota {quota:g}" if quota else "") + ")"


def cpu_baseline(w, n_distros: int, threads: int):
    """Time the oracle (C++ port of the reference algorithm) on a bounded sample: the first n distros."""
    from oracle import oracle as O
    sel = list(range(min(n_distros, w.distros.n_distros)))
    job = O.SoAJob(w.tasks, w.distros, w.hosts, sel)
    n_tasks = int(job.tasks.n)
    job.run(w.now, threads)  # warm-up pass (page faults, thread pool)
    passes, t0 = 0, time.perf_counter()
    while passes < 3 or time.perf_counter() - t0 < 10.0:  # about 10 s of CPU work
        job.run(w.now, threads)
        passes += 1
    dt = (time.perf_counter() - t0) / passes
    return {"value": n_tasks / dt, "unit": "tasks/s", "cores": threads, "kind": "port",
            "sample": f"first {len(sel)} distros ({n_tasks} tasks) of the workload, {passes} passes after one warm-up, "
                      f"{dt:.2f} s each, {cores_note()}; oracle/evg_oracle.cpp (C++17 restatement of the Go path; "
                      f"no Go toolchain in the image)",
            "decisions_per_s": len(sel) / dt}, job


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    w, _ = workload(0, 1, args.distros, args.tasks_per_distro)
    threads = usable_cores()[0]
    from oracle import oracle as O
    O.build()
    sel = list(range(min(args.ref_sample, w.distros.n_distros)))
    job = O.SoAJob(w.tasks, w.distros, w.hosts, sel)
    n_tasks = int(job.tasks.n)
    for _ in range(min(args.warmup, 1)):
        job.run(w.now, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        job.run(w.now, threads)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = n_tasks / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "tasks/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": w.name, "distros": w.distros.n_distros, "tasks_per_distro": args.tasks_per_distro,
                   "sample_distros": len(sel), "parallelism": f"{threads} host threads, one distro per work item"},
        "decisions_per_s": len(sel) / dt,
        "cpu_baseline": {"value": val, "unit": "tasks/s", "cores": threads, "kind": "port",
                         "sample": f"each step = first {len(sel)} distros ({n_tasks} tasks) of the workload; {cores_note()}; "
                                   "oracle/evg_oracle.cpp, C++17 restatement of the Go reference (Go toolchain absent)"},
        "e2e": {"value": val, "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--distros", type=int, default=1000, help="distros per GPU (configs[1]: 1000)")
    ap.add_argument("--tasks-per-distro", type=int, default=10_000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-sample", type=int, default=300, help="distros per reference/cpu_baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from evergreen_b200 import dist as edist
    from evergreen_b200 import scheduler

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; evergreen_b200 has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    prev_affinity = bind_to_gpu_node(torch, local_rank)
    # NCCL / torchrun may write banners to fd 1; keep stdout clean for the single JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w, shards = workload(rank, world, args.distros, args.tasks_per_distro)
    D_local = w.distros.n_distros
    D_total = world * args.distros
    # a dedicated (non-default) stream: kernels, NCCL and the timing events all live on it
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng = scheduler.Engine(local_rank, stream.cuda_stream)
    # allocator results go straight into the all-gather send buffer
    # N > 1: two buffer sets, the all-gather of tick k on its own stream under tick k+1's planner
    pg = edist.PipelinedGather(shards, dev)
    gather = pg.slots[0]
    send = gather.send
    eng.bind_result_buffer(send.data_ptr(), shards.max_shard)
    eng.upload(w.tasks, w.distros, w.hosts)
    tick = [0]

    def step():
        k = tick[0]
        tick[0] += 1
        if world == 1:
            eng.run(w.now)
            return
        pg.before_tick(k, stream)
        eng.bind_result_buffer(pg.send(k).data_ptr(), shards.max_shard)
        eng.run(w.now)
        pg.launch(k, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    launches_per_step = eng.last_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sort_ms = total_ms = 0.0
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    pg.drain(stream)  # every tick's gathered result is complete inside the timed region
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    total_ms, sort_ms = eng.last_timing_ms()  # last step's own CUDA-event split (same stream)
    kern_ms = eng.kernel_timing_ms(min(args.steps, 128))  # the dominant kernel, every timed step
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    tasks_total = world * args.distros * args.tasks_per_distro
    value = tasks_total / (ms_per_step * 1e-3)

    # ---- end to end through the public API: host buffers in and out, every step ----
    def pinned_like(a):
        v = a.view(np.int32) if a.dtype == np.uint32 else a  # torch pins signed views; same bytes
        p = torch.from_numpy(v).pin_memory().numpy()
        return p.view(a.dtype)
    for name, _ in w.tasks.COLUMNS:
        setattr(w.tasks, name, pinned_like(getattr(w.tasks, name)))
    for name, _ in w.hosts.COLUMNS:
        setattr(w.hosts, name, pinned_like(getattr(w.hosts, name)))
    h2d = w.tasks.nbytes() + w.distros.nbytes() + w.hosts.nbytes()
    eng.bind_result_buffer(send.data_ptr(), shards.max_shard)
    po, ao = eng.plan_and_alloc_batch(w.tasks, w.distros, w.hosts, w.now)  # warm-up (buffers sized)
    d2h = po.nbytes() + ao.nbytes()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        po, ao = eng.plan_and_alloc_batch(w.tasks, w.distros, w.hosts, w.now)
        if world > 1:
            gather.gather()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = tasks_total / float(t.item())
    clocks = sampler.stop() if sampler else None  # sampled from before the timed loop to the end of the e2e loop
    new_hosts_checksum = int(ao.result["new_hosts"].astype(np.int64).sum())

    line = None
    if rank == 0:
        peak, peak_src = load_peaks()
        alg_bytes = w.algorithmic_bytes()  # per GPU per step
        step_s = ms_per_step * 1e-3
        # the dominant kernel is the on-chip planner: its algorithmic bytes are the task/edge/group terms
        kern_bytes = 60 * w.tasks.n_tasks + 4 * w.tasks.n_edges + 96 * w.distros.n_groups
        kern_s = float(np.mean(kern_ms)) * 1e-3
        achieved = kern_bytes / kern_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per task from the committed ncu --set full capture
        if os.path.exists(tpath):
            try:
                traffic = float(json.load(open(tpath))["dram_bytes_per_task"]) * w.tasks.n_tasks
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "tasks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": f"configs[1]: {args.distros} distros x {args.tasks_per_distro} tasks each per GPU, "
                                   "10% of tasks in task groups, 5 hosts per distro",
                       "distros_total": D_total, "tasks_total": tasks_total, "global_batch": tasks_total,
                       "parallelism": f"distro-sharded x{world} (LPT), 1 all-gather of 16 B/distro per step"
                                      + (", issued on a second stream under the next tick's planner (double-buffered)" if world > 1 else ""),
                       "l2": "inputs (480 MB SoA per GPU) exceed the 126 MB L2; no flush needed"},
            "decisions_per_s": D_total / step_s,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "tasks/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": float(t.item()) * 1e3, "api": "Engine.plan_and_alloc_batch (evg_plan_and_alloc_batch), pinned host columns",
                    "cpu_affinity": "GPU-local NUMA node" if prev_affinity else "unbound"},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "k_plan_smem<1024,12,1> (on-chip planner, one CTA per distro), CUDA events on the launching "
                                   "stream around every launch of the timed region",
                         "kernel_ms": kern_s * 1e3, "algorithmic_bytes_per_launch": int(kern_bytes),
                         "kernel_share_of_step": kern_s / step_s,
                         "whole_tick": {"achieved": alg_bytes / step_s / 1e9, "frac": alg_bytes / step_s / 1e9 / peak,
                                        "algorithmic_bytes_per_step": int(alg_bytes)}},
            "checksum_new_hosts": new_hosts_checksum,
        }
        if not args.no_cpu_baseline and world == 1:
            if prev_affinity:
                os.sched_setaffinity(0, prev_affinity)  # the baseline gets every host core back
            cb, _ = cpu_baseline(w, args.ref_sample, usable_cores()[0])
            line["cpu_baseline"] = cb
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
