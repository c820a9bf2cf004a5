#!/usr/bin/env python3
"""Diagnostic: a synth config through one build of the library, `reps` resident ticks, every tick compared with the
oracle distro by distro; prints where the first differences are.
  python profiles/diag_c2.py [config-number] [reps] [path/to/lib.so]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evergreen_b200 import _lib as L  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if len(sys.argv) > 3:
    lib = C.CDLL(sys.argv[3])
    for name, (res, args) in L.SYMBOLS.items():
        if hasattr(lib, name):
            fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
    L._lib = lib
from evergreen_b200 import scheduler, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

if cfg == "c3":      # 48 distros x 100k tasks, configs[2]'s mix: the general path with every kind of unit
    w = synth.config(3, 0.0048, each=True)
elif cfg == "c5s":   # configs[4] scaled: every route in one tick
    w = synth.config(5, 0.3)
elif cfg == "mixed":
    import numpy as _np
    w = synth.make(_np.array([20, 300, 700, 3000, 9000, 14000, 1, 40000, 1100, 5000, 250, 12288, 12289, 90000]), 77, zipf_priority=True, tg_frac=0.12,
                   met_dep_frac=0.02, unmet_dep_frac=0.03, group_versions_frac=0.3, includes_dependencies=True, n_hosts=120)
else:
    w = synth.config(int(cfg))
job = O.SoAJob(w.tasks, w.distros, w.hosts, None)
ref = job.run(w.now, 16)
toff = w.distros.task_off
eng = scheduler.Engine(0)
eng.upload(w.tasks, w.distros, w.hosts)
shown = 0
bad_ticks = 0
for rep in range(reps):
    eng.run(w.now)
    po, ao = eng.download()
    if np.array_equal(po.order, ref["order"]) and np.array_equal(po.total_value, ref["total_value"]):
        continue
    bad_ticks += 1
    for d in range(w.distros.n_distros):
        a, b = int(toff[d]), int(toff[d + 1])
        go, gv, ro, rv = po.order[a:b], po.total_value[a:b], ref["order"][a:b], ref["total_value"][a:b]
        if np.array_equal(go, ro) and np.array_equal(gv, rv):
            continue
        shown += 1
        if shown > 6:
            break
        ne = np.nonzero((go != ro) | (gv != rv))[0]
        perm = np.array_equal(np.sort(go), np.arange(b - a))
        print(f"tick {rep} distro {d}: n={b - a} off0={a & 3} groups={int(w.distros.group_off[d + 1] - w.distros.group_off[d])} "
              f"differing ranks {ne.size} [{ne[0]}..{ne[-1]}] permutation={perm}")
        # per-task value on both sides
        gval = np.empty(b - a, np.int64); gval[go] = gv
        rval = np.empty(b - a, np.int64); rval[ro] = rv
        wrong = np.nonzero(gval != rval)[0]
        print(f"   tasks whose TotalValue differs: {wrong.size}", wrong[:10])
        for t in wrong[:4]:
            print(f"   task {t} (tile {(t + (a & 3)) // 1024}, slot {(t + (a & 3)) % 1024}, gid {int(w.tasks.group_id[a + t])}): gpu value {gval[t]} at rank "
                  f"{int(np.nonzero(go == t)[0][0]) if perm else -1}, oracle value {rval[t]} at rank {int(np.nonzero(ro == t)[0][0])}")
        if not perm:
            cnt = np.bincount(go.astype(np.int64), minlength=b - a)
            print("   missing tasks", np.nonzero(cnt == 0)[0][:8], "duplicated", np.nonzero(cnt > 1)[0][:8])
        k = int(ne[0])
        for r in range(max(0, k - 1), min(b - a, k + 3)):
            print(f"   rank {r}: gpu task {go[r]} v {gv[r]} | oracle task {ro[r]} v {rv[r]}")
print(f"{os.path.basename(sys.argv[3]) if len(sys.argv) > 3 else 'in-tree'}: bad ticks {bad_ticks} of {reps}")
