#!/usr/bin/env python3
"""Round-2 rows at scale, timed through the public API (wall clock of the C-ABI call: host buffers in and out, so H2D /
D2H are inside).  Prints one JSON line.
  L     evg_prioritize_legacy_batch   500 distros x 20 000 tasks (1e7), three lists per distro
  f.3   evg_dag_rebuild_batch         200 queues x 10 000 items (2e6), ~1 dependency per item, 25 % in task groups
  f.4   evg_download_queue            the persisted slice of 48 x 100 000-task queues (10 000 ranks each) vs evg_download
  f.1   evg_plan_from_finder          3000 distros, ~8e6 candidates (finder + predicate + compaction + upload), then the tick"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from evergreen_b200 import _lib as L  # noqa: E402
from evergreen_b200 import model as M  # noqa: E402
from evergreen_b200 import scheduler, soa, synth  # noqa: E402

NOW = synth.NOW_NS
eng = scheduler.Engine(0)
out = {}


def best(fn, reps=3):
    t = 1e9
    r = None
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); t = min(t, time.perf_counter() - t0)
    return t, r


# ---- row L
rnd = np.random.default_rng(5)
D, per = 500, 20000
n = D * per
prio = rnd.choice([0, 1, 5, 100, 101, 500], size=n).astype(np.int64)
req = rnd.choice([L.EVG_LF_REQ_SYSTEM, L.EVG_LF_REQ_PATCH, L.EVG_LF_REQ_OTHER], size=n, p=[0.45, 0.5, 0.05]).astype(np.uint32)
table = soa.LegacyTable(
    priority=prio, ingest_ns=(NOW - rnd.integers(0, 10 ** 6, n) * M.SECOND).astype(np.int64),
    expected_ns=rnd.integers(1, 100, n).astype(np.int64) * M.MINUTE, num_dependents=rnd.integers(0, 4, n).astype(np.int32),
    revision_order=rnd.integers(0, 1000, n).astype(np.int32), project_id=np.zeros(n, np.int32), tg_rank=np.full(n, -1, np.int32),
    tg_pair_id=np.full(n, -1, np.int32), task_group_order=np.zeros(n, np.int32),
    presort_rank=np.tile(np.arange(per, dtype=np.int32), D),
    flags=req | np.where(rnd.random(n) < 0.1, L.EVG_LF_GENERATE, 0).astype(np.uint32),
    task_off=(np.arange(D + 1, dtype=np.int64) * per),
    list_mode=np.tile(np.array([L.EVG_LEGACY_MODE_INGEST, L.EVG_LEGACY_MODE_INGEST, L.EVG_LEGACY_MODE_REVISION], np.uint8), D))
eng.prioritize_legacy_batch(table)
t, (order, count, status) = best(lambda: eng.prioritize_legacy_batch(table))
out["legacy_prioritizer"] = {"tasks": n, "distros": D, "ms": t * 1e3, "tasks_per_s": n / t, "kept": int(count.sum()), "status_ok": bool((status == 0).all())}
del table, order

# ---- f.3
Dq, per = 200, 10000
N = Dq * per
item_off = np.arange(Dq + 1, dtype=np.int64) * per
deg = rnd.choice([0, 0, 1, 1, 2, 3], size=N)
dep_off = np.zeros(N + 1, np.int64); np.cumsum(deg, out=dep_off[1:])
E = int(dep_off[-1])
dep_item = rnd.integers(0, per, E).astype(np.int32)  # queue-local item indices (both directions: cycles happen)
grouped = rnd.random(N) < 0.25
group_id = np.where(grouped, rnd.integers(0, 40, N), -1).astype(np.int32)
group_index = rnd.integers(0, 9, N).astype(np.int32)
group_off = np.arange(Dq + 1, dtype=np.int64) * 40
eng.dag_rebuild_batch(item_off, group_off, dep_off, dep_item, group_id, group_index)
t, res = best(lambda: eng.dag_rebuild_batch(item_off, group_off, dep_off, dep_item, group_id, group_index))
out["dag_rebuild"] = {"items": N, "queues": Dq, "edges": E, "ms": t * 1e3, "items_per_s": N / t, "cycle_items": int(res[2].sum())}

# ---- f.4
w = synth.config(3, 0.0048, each=True)
eng.upload(w.tasks, w.distros, w.hosts)
eng.run(w.now)
eng.download(); eng.download_queue(task_off=w.distros.task_off)
t_full, _ = best(lambda: eng.download())
t_q, (off, items) = best(lambda: eng.download_queue(task_off=w.distros.task_off))
out["persisted_queue"] = {"tasks": w.n_tasks, "distros": w.distros.n_distros, "download_all_ms": t_full * 1e3, "download_queue_ms": t_q * 1e3,
                          "items": int(items.shape[0]), "bytes": int(items.nbytes)}
del w

# ---- f.1 -> planner
D, P = 3000, 64
sizes = rnd.integers(0, 5400, D)
off = np.zeros(D + 1, np.int64); np.cumsum(sizes, out=off[1:])
T = int(off[-1])
sched = (rnd.integers(0, 256, T) | 0x0F * (rnd.random(T) < 0.85)).astype(np.uint8)
project = rnd.integers(-1, P, T).astype(np.int32)
pflags = (rnd.integers(0, 16, P) | 1).astype(np.uint8)
nvalid = np.where(rnd.random(D) < 0.3, rnd.integers(1, 6, D), 0)
voff = np.zeros(D + 1, np.int64); np.cumsum(nvalid, out=voff[1:])
vidx = rnd.integers(-1, P, int(voff[-1])).astype(np.int32)
finder = rnd.integers(0, 3, D).astype(np.uint8)
n_dep = rnd.integers(0, 3, T)
doff = np.zeros(T + 1, np.int64); np.cumsum(n_dep, out=doff[1:])
Ed = int(doff[-1])
distro_of = np.repeat(np.arange(D), sizes)
ref = (off[distro_of][np.repeat(np.arange(T), n_dep)] + rnd.integers(0, 1 << 30, Ed) % np.maximum(sizes[distro_of][np.repeat(np.arange(T), n_dep)], 1)).astype(np.int32)
kind = rnd.integers(0, 3, Ed).astype(np.uint8)
ref[kind == 1] %= 4000
deps = soa.DepsTable(doff, kind, ref, rnd.integers(0, 4, Ed).astype(np.uint8), rnd.integers(0, 3, T).astype(np.uint8),
                     (rnd.random(T) < 0.1).astype(np.uint8), rnd.integers(0, 3, 4000).astype(np.uint8))
table = soa.RunnableTable(off, sched, project, pflags, voff, vidx, finder, deps)
wc = synth.make(sizes, 91, zipf_priority=True, tg_frac=0.1, met_dep_frac=0.03, includes_dependencies=True)
wc.tasks.flags &= ~np.uint32(L.EVG_TF_DEPS_MET)
fin = np.where(rnd.random(Ed) < 0.5, NOW - rnd.integers(0, 10 ** 12, Ed), M.ZERO_TIME).astype(np.int64)
eng.plan_from_finder(table, wc.tasks, wc.distros, None, fin, NOW)
t, (runnable, count) = best(lambda: eng.plan_from_finder(table, wc.tasks, wc.distros, None, fin, NOW))
ms = []
for _ in range(5):
    eng.run(NOW); ms.append(eng.last_timing_ms()[0])
out["finder_to_planner"] = {"candidates": T, "distros": D, "kept": int(count.sum()), "in_queue_edges": wc.tasks.n_edges,
                            "plan_from_finder_ms": t * 1e3, "candidates_per_s": T / t, "tick_ms": float(np.median(ms))}
print(json.dumps(out))
