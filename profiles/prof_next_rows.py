"""The §8f.1 rows at scale for the ncu launch list: dependency predicate over 4e6 tasks / ~6e6 dependencies and the
task finders' filter over 8e6 candidates in 3000 distros (algorithmic bytes printed for the roofline arithmetic)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from evergreen_b200 import scheduler, soa

rng = np.random.default_rng(9)
eng = scheduler.Engine(0)
D, P = 3000, 64
sizes = rng.integers(0, 5400, D)
off = np.zeros(D + 1, np.int64); np.cumsum(sizes, out=off[1:])
T = int(off[-1])
sched = (rng.integers(0, 256, T) | 0x0F * (rng.random(T) < 0.85)).astype(np.uint8)
project = rng.integers(-1, P, T).astype(np.int32)
pflags = rng.integers(0, 16, P).astype(np.uint8)
nvalid = np.where(rng.random(D) < 0.3, rng.integers(1, 6, D), 0)
voff = np.zeros(D + 1, np.int64); np.cumsum(nvalid, out=voff[1:])
vidx = rng.integers(-1, P, int(voff[-1])).astype(np.int32)
finder = rng.integers(1, 3, D).astype(np.uint8)
n_dep = rng.integers(0, 4, T)
doff = np.zeros(T + 1, np.int64); np.cumsum(n_dep, out=doff[1:])
E = int(doff[-1])
deps = soa.DepsTable(doff, rng.integers(0, 3, E).astype(np.uint8), rng.integers(0, T, E).astype(np.int32),
                     rng.integers(0, 4, E).astype(np.uint8), rng.integers(0, 3, T).astype(np.uint8),
                     (rng.random(T) < 0.1).astype(np.uint8), rng.integers(0, 3, 4000).astype(np.uint8))
deps.dep_ref[deps.dep_kind == 1] %= 4000
table = soa.RunnableTable(off, sched, project, pflags, voff, vidx, finder, deps)
for _ in range(3):
    met = eng.deps_met_batch(deps)
for _ in range(3):
    runnable, count = eng.find_runnable_batch(table)
kept = int(count.sum())
print(f"tasks {T} deps {E} kept {kept}")
print(f"k_deps_met algorithmic bytes: {8 * T + 6 * E + T + 1 * T + 1 * E:d} (dep_off 8/task, kind+ref+want 6/dep, pre 1/task, one state byte per dep, met 1/task)")
print(f"k_runnable algorithmic bytes: {T * (1 + 4 + 1) + 4 * kept:d} (sched 1 + project 4 + met 1 per candidate, 4 per survivor)")

# §8f.2: expected-duration statistics, 8e6 finished-task rows over 200k keys
R, K = 8_000_000, 200_000
key = rng.integers(0, K, R).astype(np.int32)
taken = rng.integers(10 ** 9, 3 * 3600 * 10 ** 9, R)
flags = (rng.random(R) < 0.9).astype(np.uint8)
now = 10 ** 18
start = now - rng.integers(0, 6 * 24 * 3600 * 10 ** 9, R)
rows = soa.DurationRows(key, taken, start, start + taken, flags, K, now - 7 * 24 * 3600 * 10 ** 9, now + 10 ** 15)
for _ in range(3):
    st = eng.expected_durations_batch(rows)
print(f"k_dur_sum + k_dur_dev algorithmic bytes: {2 * R * 29 + 24 * K:d} (29 B/row read by each of the two passes, 24 B/key out)")
