#!/usr/bin/env python3
"""Throughput of evg_intern_columns (host C++): `python profiles/intern_bench.py [distros] [tasks_per_distro] [threads...]`.
Strings shaped like Evergreen's: ~60-byte task ids, ~35-byte versions, a task-group key on every tenth task, 7 % of the
tasks with one in-queue dependency."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evergreen_b200 import _lib as L  # noqa: E402
from evergreen_b200.soa import pack_strings  # noqa: E402


def make(D, per, seed=3):
    rng = np.random.default_rng(seed)
    T = D * per
    ids = [f"evergreen_ubuntu2204_test_{i % 977}_patch_{i:012x}_24_09_22_17_03_41" for i in range(T)]
    vers = [f"evergreen_{i // per}_{(i % per) // 400:03d}abcdef0123456789" for i in range(T)]
    gk = [(f"tg{(i % per) // 25}_bv_proj_{vers[i]}" if i % 10 == 0 else "") for i in range(T)]
    ndep = (rng.random(T) < 0.07).astype(np.int64)
    dep_off = np.zeros(T + 1, np.int64)
    np.cumsum(ndep, out=dep_off[1:])
    tgt = [ids[(i // per) * per + int(rng.integers(per))] for i in np.nonzero(ndep)[0]]
    return T, ids, vers, gk, dep_off, tgt


def run(D=200, per=10000, threads=(1, 4, 16)):
    lib = L.load()
    T, ids, vers, gk, dep_off, tgt = make(D, per)
    idb, ido = pack_strings(ids); vb, vo = pack_strings(vers); gb, go = pack_strings(gk); db, do = pack_strings(tgt)
    task_off = np.arange(D + 1, dtype=np.int64) * per
    gmax = np.ones(T, np.int32)
    out = dict(group_id=np.empty(T, np.int32), version_id=np.empty(T, np.int32), group_off=np.zeros(D + 1, np.int64),
               n_versions=np.zeros(D, np.int32), group_max_hosts=np.empty(T, np.int32), group_first=np.empty(T, np.int64),
               dep_off=np.zeros(T + 1, np.int64), dep_idx=np.empty(max(len(tgt), 1), np.int32))
    sc = lambda b, o: L.StrColStruct(L.ptr(b), L.ptr(o))  # noqa: E731
    ins = L.StringColsStruct(T, D, L.ptr(task_off), sc(idb, ido), sc(vb, vo), sc(gb, go), L.ptr(gmax), L.ptr(dep_off), sc(db, do))
    outs = L.InternOutStruct(*[L.ptr(out[k]) for k in ("group_id", "version_id", "group_off", "n_versions", "group_max_hosts",
                                                       "group_first", "dep_off", "dep_idx")])
    nbytes = idb.nbytes + vb.nbytes + gb.nbytes + db.nbytes
    res = []
    for th in threads:
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            rc = lib.evg_intern_columns(C.byref(ins), C.byref(outs), int(th))
            best = min(best, time.perf_counter() - t0)
            assert rc == 0
        res.append({"threads": int(th), "tasks_per_s": T / best, "string_GB_per_s": nbytes / best / 1e9})
    return {"tasks": T, "distros": D, "string_bytes": int(nbytes), "groups": int(out["group_off"][-1]), "edges": int(out["dep_off"][-1]), "runs": res}


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    print(run(*(a[:2] or (200, 10000)), threads=tuple(a[2:]) or (1, 4, 16)))
