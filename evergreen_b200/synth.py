"""Synthetic scheduler ticks in the shapes BASELINE.json names.

Counter-based splitmix64 (the same generator a Go/C++ harness can reproduce):
value k of stream s under seed S is mix64(S + stream_salt(s) + (k+1)*GOLDEN).
Everything is produced column-wise with numpy straight into the SoA tables the
C-ABI takes; oracle.SoAJob rebuilds reference-shaped strings from the same
tables for the CPU side.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib as L
from . import model as M
from .soa import DistroTable, HostSoA, TaskSoA

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
NOW_NS = 1_800_000_000 * 10 ** 9
SEED_BASE = 0xE5E60000


def mix64(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return z


class Rng:
    def __init__(self, seed: int):
        self.seed = np.uint64(seed)
        self.stream = 0

    def u64(self, n: int) -> np.ndarray:
        self.stream += 1
        with np.errstate(over="ignore"):
            salt = mix64(np.array([self.stream], dtype=np.uint64) * np.uint64(0xD1342543DE82EF95))[0]
            k = (np.arange(1, n + 1, dtype=np.uint64) * GOLDEN) + self.seed + salt
        return mix64(k)

    def uniform(self, n: int) -> np.ndarray:
        return (self.u64(n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))

    def integers(self, n: int, lo: int, hi: int) -> np.ndarray:
        """uniform integers in [lo, hi]"""
        span = hi - lo + 1
        return (lo + (self.uniform(n) * span).astype(np.int64)).clip(lo, hi)


@dataclass
class Workload:
    name: str
    now: int
    tasks: TaskSoA
    distros: DistroTable
    hosts: Optional[HostSoA]

    @property
    def n_tasks(self) -> int:
        return self.tasks.n_tasks

    def algorithmic_bytes(self) -> int:
        """60*T + 4*E + 28*H + 96*G + 16*D (compulsory traffic only)."""
        H = self.hosts.n_hosts if self.hosts is not None else 0
        return (60 * self.tasks.n_tasks + 4 * self.tasks.n_edges + 28 * H + 96 * self.distros.n_groups +
                16 * self.distros.n_distros)


def _ranges(off: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """Concatenated index ranges [off[i], off[i+1]) for i in ids, in that order."""
    lens = (off[ids + 1] - off[ids]).astype(np.int64)
    if lens.sum() == 0:
        return np.zeros(0, dtype=np.int64)
    starts = np.repeat(off[ids], lens)
    within = np.arange(int(lens.sum()), dtype=np.int64) - np.repeat(np.cumsum(lens) - lens, lens)
    return starts + within


def take_distros(w: Workload, ids) -> Workload:
    """The sub-tick of the distros `ids` (in that order): what one rank of a distro-sharded job uploads.  Task groups,
    versions and in-queue dependency edges are distro-local, so rows move as they are."""
    ids = np.asarray(ids, dtype=np.int64)
    t, d, h = w.tasks, w.distros, w.hosts
    rows = _ranges(d.task_off, ids)
    cols = {name: getattr(t, name)[rows] for name, _ in t.COLUMNS}
    dep_off = dep_idx = None
    if t.n_edges:
        deg = (t.dep_off[rows + 1] - t.dep_off[rows]).astype(np.int64)
        dep_off = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
        dep_idx = t.dep_idx[_ranges(t.dep_off, rows)] if deg.sum() else np.zeros(0, dtype=np.int32)
    tasks = TaskSoA(**cols, dep_off=dep_off, dep_idx=dep_idx).normalize()
    n = (d.task_off[ids + 1] - d.task_off[ids]).astype(np.int64)
    g = (d.group_off[ids + 1] - d.group_off[ids]).astype(np.int64)
    distros = DistroTable(np.concatenate([[0], np.cumsum(n)]).astype(np.int64), np.concatenate([[0], np.cumsum(g)]).astype(np.int64),
                          d.cfg[ids], d.group_max_hosts[_ranges(d.group_off, ids)]).normalize()
    hosts = None
    if h is not None:
        hr = _ranges(h.host_off, ids)
        hn = (h.host_off[ids + 1] - h.host_off[ids]).astype(np.int64)
        hosts = HostSoA(h.flags[hr], h.group_id[hr], h.expected_ns[hr], h.std_ns[hr], h.start_ns[hr],
                        np.concatenate([[0], np.cumsum(hn)]).astype(np.int64), h.cfg[ids]).normalize()
    return Workload(f"{w.name} [{len(ids)} of {d.n_distros} distros]", w.now, tasks, distros, hosts)


def _zipf_priorities(rng: Rng, n: int, s: float = 1.1, kmax: int = 100) -> np.ndarray:
    ranks = np.arange(1, kmax + 2, dtype=np.float64)
    w = ranks ** (-s)
    cdf = np.cumsum(w) / w.sum()
    return np.searchsorted(cdf, rng.uniform(n)).clip(0, kmax).astype(np.int32)


def make(sizes: np.ndarray, seed: int, *, name: str = "synthetic", now: int = NOW_NS, zipf_priority: bool = False,
         unmet_dep_frac: float = 0.0, met_dep_frac: float = 0.0, tg_frac: float = 0.10,
         group_versions_frac: float = 0.0, custom_factor_frac: float = 0.10, includes_dependencies: bool = False,
         n_hosts: int = 0, providers: Tuple[float, float, float] = (1.0, 0.0, 0.0)) -> Workload:
    """Build one tick. `sizes[d]` = tasks queued on distro d.
    providers = fractions (ephemeral, docker+pool, static)."""
    sizes = np.asarray(sizes, dtype=np.int64)
    D = int(sizes.shape[0])
    rng = Rng(seed)
    task_off = np.zeros(D + 1, dtype=np.int64)
    np.cumsum(sizes, out=task_off[1:])
    T = int(task_off[-1])
    distro_of = np.repeat(np.arange(D, dtype=np.int64), sizes)
    local = np.arange(T, dtype=np.int64) - task_off[distro_of]

    expected = rng.integers(T, 10 * M.SECOND, 2 * M.HOUR)
    queue_basis = now - rng.integers(T, 0, 72 * M.HOUR)
    wait_basis = now - rng.integers(T, 0, 3 * M.HOUR)
    u = rng.uniform(T)
    req = np.where(u < 0.40, L.EVG_TF_REQ_PATCH, np.where(u < 0.45, L.EVG_TF_REQ_MERGE_QUEUE, L.EVG_TF_REQ_OTHER))
    priority = _zipf_priorities(rng, T) if zipf_priority else np.zeros(T, dtype=np.int32)
    numdep = np.floor(np.log(np.maximum(rng.uniform(T), 1e-300)) / np.log(0.3)).astype(np.int32).clip(0, 10000)
    flags = req.astype(np.uint32)
    flags |= np.where(rng.uniform(T) < 0.01, L.EVG_TF_GENERATE, 0).astype(np.uint32)
    flags |= np.where(rng.uniform(T) < 0.01, L.EVG_TF_STEPBACK, 0).astype(np.uint32)
    flags |= np.where(rng.uniform(T) < 0.005, L.EVG_TF_OTHER_DISTRO, 0).astype(np.uint32)

    # versions: ~50 tasks per version per distro
    n_versions = np.maximum(1, sizes // 50).astype(np.int64)
    version = (rng.uniform(T) * n_versions[distro_of]).astype(np.int64).clip(0, None)
    version = np.minimum(version, n_versions[distro_of] - 1)

    # task groups: a random 10% of tasks, chunked into groups in order of appearance
    is_tg = rng.uniform(T) < tg_frac
    tg_idx = np.nonzero(is_tg)[0]
    gid = np.full(T, -1, dtype=np.int64)
    tgo = np.zeros(T, dtype=np.int32)
    group_off = np.zeros(D + 1, dtype=np.int64)
    group_max_hosts = np.zeros(0, dtype=np.int32)
    if tg_idx.shape[0]:
        n = tg_idx.shape[0]
        d_tg = distro_of[tg_idx]
        first_in_distro = np.ones(n, dtype=bool)
        first_in_distro[1:] = d_tg[1:] != d_tg[:-1]
        brk = (rng.uniform(n) < 0.2) | first_in_distro
        # position inside the run since the last break; force a break every 8 members
        start_pos = np.maximum.accumulate(np.where(brk, np.arange(n), 0))
        pos = np.arange(n) - start_pos
        brk |= (pos % 8 == 0)
        start_pos = np.maximum.accumulate(np.where(brk, np.arange(n), 0))
        pos = np.arange(n) - start_pos
        gglobal = np.cumsum(brk) - 1                    # global group number
        gfirst = np.zeros(D + 1, dtype=np.int64)        # groups before each distro
        counts = np.bincount(d_tg[brk], minlength=D)
        np.cumsum(counts, out=gfirst[1:])
        group_off = gfirst.copy()
        gid[tg_idx] = gglobal - gfirst[d_tg]
        tgo[tg_idx] = (pos + 1).astype(np.int32)
        version[tg_idx] = version[tg_idx[start_pos]]    # a group lives in one version
        group_max_hosts = rng.integers(int(gfirst[-1]), 1, 4).astype(np.int32)

    # dependency edges onto other in-queue tasks
    dep_off = None
    dep_idx = None
    deps_met = np.ones(T, dtype=bool)
    if unmet_dep_frac > 0 or met_dep_frac > 0:
        ud = rng.uniform(T)
        has_unmet = (ud < unmet_dep_frac) & (sizes[distro_of] > 1)
        has_met = (ud >= unmet_dep_frac) & (ud < unmet_dep_frac + met_dep_frac) & (sizes[distro_of] > 1)
        n_dep = np.where(has_unmet | has_met, 1 + (rng.uniform(T) < 0.2), 0).astype(np.int64)
        dep_off = np.zeros(T + 1, dtype=np.int64)
        np.cumsum(n_dep, out=dep_off[1:])
        E = int(dep_off[-1])
        owner = np.repeat(np.arange(T, dtype=np.int64), n_dep)
        tgt = (rng.uniform(E) * (sizes[distro_of[owner]] - 1)).astype(np.int64)
        tgt = np.minimum(tgt, sizes[distro_of[owner]] - 2)
        tgt = np.where(tgt >= local[owner], tgt + 1, tgt)  # never depend on yourself
        dep_idx = tgt.astype(np.int32)
        deps_met = ~has_unmet
    else:
        # a few tasks wait on something outside the queue
        deps_met = rng.uniform(T) >= 0.01
    flags |= np.where(deps_met, L.EVG_TF_DEPS_MET, 0).astype(np.uint32)

    tasks = TaskSoA(priority, expected, queue_basis, wait_basis, numdep, tgo, gid.astype(np.int32),
                    version.astype(np.int32), flags, dep_off, dep_idx).normalize()

    cfg = np.zeros(D, dtype=L.DISTRO_CFG_DTYPE)
    custom = rng.uniform(D) < custom_factor_frac
    for f in ("patch_factor", "patch_time_in_queue_factor", "commit_queue_factor", "mainline_time_in_queue_factor",
              "expected_runtime_factor", "generate_task_factor", "stepback_task_factor"):
        cfg[f] = np.where(custom, rng.integers(D, 1, 100), 0)
    cfg["num_dependents_factor"] = np.where(custom, np.round(rng.uniform(D) * 100, 2), 0.0)
    prov_u = rng.uniform(D)
    provider = np.where(prov_u < providers[0], L.EVG_PROVIDER_EPHEMERAL,
                        np.where(prov_u < providers[0] + providers[1], L.EVG_PROVIDER_DOCKER, L.EVG_PROVIDER_STATIC))
    has_pool = provider == L.EVG_PROVIDER_DOCKER
    cfg["target_time_ns"] = np.where(has_pool, M.MAX_DURATION_PER_DISTRO_HOST_WITH_CONTAINERS, M.MAX_DURATION_PER_DISTRO_HOST)
    cfg["group_versions"] = (rng.uniform(D) < group_versions_frac).astype(np.int32)
    cfg["includes_dependencies"] = int(includes_dependencies)
    cfg["n_versions"] = n_versions.astype(np.int32)
    distros = DistroTable(task_off, group_off, cfg, group_max_hosts).normalize()

    hosts = None
    if n_hosts > 0:
        share = sizes.astype(np.float64) / max(1, sizes.sum())
        hcount = np.floor(share * n_hosts).astype(np.int64)
        hcount[: int(n_hosts - hcount.sum())] += 1 if D else 0
        host_off = np.zeros(D + 1, dtype=np.int64)
        np.cumsum(hcount, out=host_off[1:])
        H = int(host_off[-1])
        hd = np.repeat(np.arange(D, dtype=np.int64), hcount)
        running = rng.uniform(H) < 0.7
        found = running & (rng.uniform(H) < 0.98)
        teardown = (~running) & (rng.uniform(H) < 0.03)
        hflags = (np.where(running, L.EVG_HF_RUNNING, 0) | np.where(found, L.EVG_HF_RT_FOUND, 0) |
                  np.where(teardown, L.EVG_HF_TEARDOWN, 0)).astype(np.uint32)
        hexp = np.where(found, rng.integers(H, 10 * M.SECOND, 2 * M.HOUR), 0)
        hstd = np.where(found, hexp // 5, 0)
        elapsed = (rng.uniform(H) * 2.0 * hexp).astype(np.int64)
        hstart = np.where(found, now - elapsed, M.ZERO_TIME)
        ng = (group_off[1:] - group_off[:-1])[hd] if H else np.zeros(0, dtype=np.int64)
        in_group = running & (rng.uniform(H) < 0.05)
        pick = (rng.uniform(H) * np.maximum(ng, 1)).astype(np.int64)
        hgid = np.where(in_group, np.where((ng > 0) & (rng.uniform(H) < 0.8), np.minimum(pick, np.maximum(ng - 1, 0)),
                                           L.EVG_HG_UNQUEUED), L.EVG_HG_NONE).astype(np.int32)
        acfg = np.zeros(D, dtype=L.ALLOC_CFG_DTYPE)
        acfg["future_host_fraction"] = 0.4
        acfg["provider"] = provider
        acfg["disabled"] = (rng.uniform(D) < 0.02).astype(np.int32)
        acfg["minimum_hosts"] = rng.integers(D, 0, 2)
        acfg["maximum_hosts"] = rng.integers(D, 10, 500)
        acfg["round_up"] = (rng.uniform(D) < 0.1).astype(np.int32)
        acfg["waits_over_thresh_feedback"] = (rng.uniform(D) < 0.2).astype(np.int32)
        acfg["has_pool"] = has_pool.astype(np.int32)
        acfg["pool_max_containers"] = np.where(has_pool, 10, 0)
        acfg["parent_found"] = has_pool.astype(np.int32)
        acfg["parent_maximum_hosts"] = np.where(has_pool, rng.integers(D, 5, 50), 0)
        hosts = HostSoA(hflags, hgid, hexp, hstd, hstart, host_off, acfg).normalize()
    return Workload(name, now, tasks, distros, hosts)


def power_law_sizes(rng: Rng, D: int, alpha: float = 1.2, lo: int = 1, hi: int = 1_000_000) -> np.ndarray:
    u = np.maximum(rng.uniform(D), 1e-12)
    return np.floor(lo * u ** (-1.0 / alpha)).clip(lo, min(hi, L.MAX_TASKS_PER_DISTRO)).astype(np.int64)


def config(k: int, scale: float = 1.0, *, each: bool = False) -> Workload:
    """BASELINE.json configs[k-1].  `scale` shrinks the distro count (tests);
    `each` selects the per-distro reading of "N distros x M tasks" for C3/C4."""
    seed = SEED_BASE + k
    if k == 1:
        return make(np.array([1000]), seed, name="C1: 1 distro x 1000 tasks", n_hosts=20)
    if k == 2:
        D = max(1, int(round(1000 * scale)))
        return make(np.full(D, 10_000), seed, name=f"C2: {D} distros x 10k tasks each, uniform expected durations", n_hosts=5 * D)
    if k == 3:
        D = max(1, int(round(10_000 * scale)))
        per = 100_000 if each else 10
        return make(np.full(D, per), seed, name=f"C3: {D} distros x {per} tasks, Zipf priorities, 5% unmet deps",
                    zipf_priority=True, unmet_dep_frac=0.05, met_dep_frac=0.02, includes_dependencies=True, n_hosts=2 * D)
    if k == 4:
        D = max(1, int(round(10_000 * scale)))
        per = 1_000_000 if each else 100
        return make(np.full(D, per), seed, name=f"C4: {D} distros x {per} tasks, 50k-host pool",
                    zipf_priority=True, n_hosts=int(round(50_000 * scale)))
    if k == 5:
        D = max(1, int(round(100_000 * scale)))
        sizes = power_law_sizes(Rng(seed ^ 0x5A5A), D)
        return make(sizes, seed, name=f"C5: {D} distros, power-law queue sizes, mixed providers", zipf_priority=True,
                    unmet_dep_frac=0.03, met_dep_frac=0.01, group_versions_frac=0.2, includes_dependencies=True,
                    n_hosts=D // 2, providers=(0.6, 0.2, 0.2))
    raise ValueError(k)
