"""The legacy comparator prioritiser through the C-ABI (evg_prioritize_legacy_batch), against the
reference-held vectors and, on larger random queues, against oracle/oracle_legacy.py (literal comparators + a port of
Go's sort.Stable)."""
import json
import os
import random

import numpy as np
import pytest

from evergreen_b200 import _lib as L
from evergreen_b200 import model as M
from evergreen_b200 import scheduler as S
from evergreen_b200 import soa
from oracle import oracle_legacy as OL

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "legacy_prioritizer.json")))
NOW = G["now"]


def mk(d):
    d = dict(d)
    dp = d.pop("duration_prediction", None)
    t = M.Task(**d)
    if dp:
        t.duration_prediction = M.CachedDurationValue(value=dp.get("value", 0), ttl=dp.get("ttl", 0))
    return t


@pytest.mark.parametrize("case", G["orders"], ids=lambda c: c["ref"].split()[-1])
def test_reference_orders(engine, case):
    p = S.CmpBasedTaskPrioritizer(engine=engine, now=NOW)
    got, reasons, err = p.PrioritizeTasks("distro", [mk(t) for t in case["tasks"]], case["versions"])
    assert err is None and reasons == {}
    ids = [t.id for t in got]
    if "want_order" in case:
        assert ids == case["want_order"]
    for a, b in case.get("want_before", []):
        assert ids.index(a) < ids.index(b)


@pytest.mark.parametrize("case", G["comparators"], ids=lambda c: f'{c["comparator"]}@{c["ref"].split(":")[1]}')
def test_comparator_truth_tables_as_two_task_queues(engine, case):
    """A comparator's verdict on (t1, t2) decides the order of the two-task queue {t1, t2} whenever every earlier
    comparator of the chain ties on the pair (true for all the reference's vectors: they vary one field at a time)."""
    t1, t2 = mk(case["t1"]), mk(case["t2"])
    for t in (t1, t2):
        if not t.requester:
            t.requester = M.PATCH_VERSION_REQUESTER  # an empty requester would be dropped by the split
    versions = case["versions"]
    want = [t.id for t in OL.prioritize_tasks([mk_copy(t1), mk_copy(t2)], versions, NOW)]
    got, _, err = S.CmpBasedTaskPrioritizer(engine=engine, now=NOW).PrioritizeTasks("d", [t1, t2], versions)
    assert err is None and [t.id for t in got] == want
    if case["want"] == 1:
        assert want == ["t1", "t2"]
    elif case["want"] == -1:
        assert want == ["t2", "t1"]


def mk_copy(t):
    import copy
    return copy.deepcopy(t)


def test_split_and_merge_shapes(engine):
    p = S.CmpBasedTaskPrioritizer(engine=engine, now=NOW)
    for case in G["splits"]:
        tasks = [mk(t) for t in case["tasks"]]
        got, _, err = p.PrioritizeTasks("d", tasks, {})
        want = OL.prioritize_tasks([mk_copy(t) for t in tasks], {}, NOW)
        assert err is None and [t.id for t in got] == [t.id for t in want]
        n_high = len(case["want"]["high"])
        assert sorted(t.id for t in got[:n_high]) == sorted(case["want"]["high"])
    # merge: every list length combination up to 5, tasks that tie on everything (the presort decides inside a list)
    for nh in range(3):
        for nr in range(5):
            for npt in range(5):
                tasks = ([M.Task(id=f"h{k}", requester=M.REPOTRACKER_VERSION_REQUESTER, priority=101 + k) for k in range(nh)] +
                         [M.Task(id=f"r{k}", requester=M.REPOTRACKER_VERSION_REQUESTER, revision_order_number=k) for k in range(nr)] +
                         [M.Task(id=f"p{k}", requester=M.PATCH_VERSION_REQUESTER, ingest_time=NOW - k) for k in range(npt)])
                got, _, err = p.PrioritizeTasks("d", tasks, {})
                want = OL.prioritize_tasks([mk_copy(t) for t in tasks], {}, NOW)
                assert err is None and [t.id for t in got] == [t.id for t in want], (nh, nr, npt)
    # dropped: unrecognised requester; empty queue
    got, _, err = p.PrioritizeTasks("d", [M.Task(id="x", requester="nonsense"), M.Task(id="y", requester=M.PATCH_VERSION_REQUESTER)], {})
    assert err is None and [t.id for t in got] == ["y"]
    got, _, err = p.PrioritizeTasks("d", [], {})
    assert err is None and got == []


def random_queue(rnd, n, one_project=True, zero_runtimes=False):
    projects = ["proj"] if one_project else ["pa", "pb", "pc"]
    out = []
    for k in range(n):
        req = rnd.choice([M.REPOTRACKER_VERSION_REQUESTER, M.PATCH_VERSION_REQUESTER, M.GITHUB_PR_REQUESTER, M.TRIGGER_REQUESTER,
                          M.GITHUB_MERGE_REQUESTER, M.AD_HOC_REQUESTER, "nonsense" if rnd.random() < 0.05 else M.PATCH_VERSION_REQUESTER])
        tg = rnd.random() < 0.25
        prio = rnd.choice([0, 0, 0, 1, 5, 50, 100, 101, 150, 2 ** 40])
        if prio > M.MAX_TASK_PRIORITY and one_project and req in M.SYSTEM_VERSION_REQUESTER_TYPES:
            req = M.PATCH_VERSION_REQUESTER  # keep the high-priority list free of commit builds: byAge stays on IngestTime there
        out.append(M.Task(id=f"task_{k:05d}_{rnd.randrange(10 ** 6)}", requester=req, project=rnd.choice(projects),
                          version=f"v{rnd.randrange(6)}", build_id=f"build_{rnd.randrange(5)}",
                          task_group=f"tg{rnd.randrange(3)}" if tg else "", task_group_order=rnd.randrange(1, 5) if tg else 0,
                          priority=prio, num_dependents=rnd.choice([0, 0, 1, 2, 7]),
                          generate_task=rnd.random() < 0.1, revision_order_number=rnd.randrange(50), ingest_time=NOW - rnd.randrange(20) * M.HOUR,
                          expected_duration=0 if zero_runtimes and rnd.random() < 0.3 else rnd.randrange(1, 8) * 10 * M.MINUTE))
    return out


@pytest.mark.parametrize("seed", range(6))
def test_random_queues_match_the_literal_oracle(engine, seed):
    """Key-decomposable queues (one project): bit-equal to literal comparators + Go's sort.Stable, several distros per call."""
    rnd = random.Random(900 + seed)
    versions = {f"v{k}": (M.GITHUB_MERGE_REQUESTER if k == 0 else M.PATCH_VERSION_REQUESTER) for k in range(6)}
    batch = [(f"d{k}", random_queue(rnd, n), versions) for k, n in enumerate([0, 1, 2, 37, 400, 1500, 3])]
    res = S.CmpBasedTaskPrioritizer(engine=engine, now=NOW).prioritize_batch(batch)
    for (_, tasks, _), (got, status) in zip(batch, res):
        assert status == L.EVG_LEGACY_OK
        want = OL.prioritize_tasks([mk_copy(t) for t in tasks], versions, NOW)
        assert [t.id for t in got] == [t.id for t in want]


def test_non_decomposable_queues_are_reported(engine):
    rnd = random.Random(77)
    tasks = random_queue(rnd, 300, one_project=False)
    got, reasons, err = S.CmpBasedTaskPrioritizer(engine=engine, now=NOW).PrioritizeTasks("d", tasks, {})
    assert got is None and isinstance(err, S.NotDecomposableError)
    kept = [t for t in tasks if soa.legacy_list_of(t) != 3]
    assert sorted(t.id for t in err.tasks) == sorted(t.id for t in kept)  # still a permutation of the kept tasks
    # zero and non-zero expected durations mixed
    tasks = [M.Task(id=f"t{k}", requester=M.PATCH_VERSION_REQUESTER, expected_duration=0 if k % 2 else M.MINUTE,
                    duration_prediction=M.CachedDurationValue(value=0 if k % 2 else M.MINUTE, ttl=M.HOUR, collected_at=NOW)) for k in range(6)]
    table = soa.marshal_legacy([("d", tasks, {})], None)
    assert int(table.list_mode[1]) == L.EVG_LEGACY_MODE_LITERAL


def test_large_queue_properties(engine):
    """200k tasks in one queue: every list sorted by its key, lists interleaved by the closed form."""
    rnd = np.random.default_rng(5)
    n = 200_000
    prio = rnd.choice([0, 1, 5, 100, 101, 500], size=n).astype(np.int64)
    req = rnd.choice([L.EVG_LF_REQ_SYSTEM, L.EVG_LF_REQ_PATCH, L.EVG_LF_REQ_OTHER], size=n, p=[0.45, 0.5, 0.05]).astype(np.uint32)
    table = soa.LegacyTable(
        priority=prio, ingest_ns=(NOW - rnd.integers(0, 10 ** 6, n) * M.SECOND).astype(np.int64),
        expected_ns=rnd.integers(1, 100, n).astype(np.int64) * M.MINUTE, num_dependents=rnd.integers(0, 4, n).astype(np.int32),
        revision_order=rnd.integers(0, 1000, n).astype(np.int32), project_id=np.zeros(n, np.int32), tg_rank=np.full(n, -1, np.int32),
        tg_pair_id=np.full(n, -1, np.int32), task_group_order=np.zeros(n, np.int32), presort_rank=rnd.permutation(n).astype(np.int32),
        flags=req | np.where(rnd.random(n) < 0.1, L.EVG_LF_GENERATE, 0).astype(np.uint32),
        task_off=np.array([0, n], np.int64), list_mode=np.array([L.EVG_LEGACY_MODE_INGEST, L.EVG_LEGACY_MODE_INGEST, L.EVG_LEGACY_MODE_REVISION], np.uint8))
    order, count, status = engine.prioritize_legacy_batch(table)
    lst = np.where(prio > 100, 0, np.where(req == L.EVG_LF_REQ_SYSTEM, 2, np.where(req == L.EVG_LF_REQ_PATCH, 1, 3)))
    kept = int((lst < 3).sum())
    assert int(count[0]) == kept and int(status[0]) == 0
    o = order[:kept]
    assert (order[kept:] == -1).all() and np.array_equal(np.sort(o), np.nonzero(lst < 3)[0])
    nH, nP, nR = [(lst == k).sum() for k in range(3)]
    assert (lst[o[:nH]] == 0).all()
    m = min(nP, nR)
    assert (lst[o[nH:nH + 2 * m:2]] == 1).all() and (lst[o[nH + 1:nH + 2 * m:2]] == 2).all()
    # inside the patch list: merge-queue bit is 0 everywhere here, so (priority desc, num_dependents desc, generator first, ingest asc, expected desc, presort)
    pl = o[lst[o] == 1]
    gen = (table.flags[pl] & L.EVG_LF_GENERATE) != 0
    key = np.stack([-table.priority[pl], -table.num_dependents[pl].astype(np.int64), (~gen).astype(np.int64), table.ingest_ns[pl],
                    -table.expected_ns[pl], table.presort_rank[pl].astype(np.int64)], axis=1)
    assert (np.lexsort(key.T[::-1]) == np.arange(len(pl))).all()


# ---------------------------------------------------------------- DAG dispatcher rebuild
def _tq(items):
    return M.TaskQueue(distro="d", queue=[M.TaskQueueItem(id=it["id"], group=it.get("group", ""), build_variant=it.get("build_variant", ""),
                                                           project=it.get("project", ""), version=it.get("version", ""),
                                                           group_index=it.get("group_index", 0), dependencies=list(it.get("dependencies", [])))
                                          for it in items])


def test_dag_rebuild_reference_order(engine):
    from oracle import oracle_dag as OD
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "dag_dispatcher.json")))
    (order, n_cycles, units), = S.rebuild_dag_dispatchers([_tq(g["items"])], engine=engine)
    assert order == g["expected_order"] and n_cycles == 0
    assert len(units) == g["n_task_groups"] and {len(v) for v in units.values()} == {g["group_size"]}
    assert units == OD.rebuild(g["items"])[2]


def test_dag_rebuild_random_batches_match_the_oracle(engine):
    from oracle import oracle_dag as OD
    rnd = random.Random(12)
    batches = []
    for n in (0, 1, 2, 50, 700, 3000, 10000):
        items = []
        for k in range(n):
            deps = [str(rnd.randrange(n)) for _ in range(rnd.choice([0, 0, 0, 1, 2, 3]))]
            if rnd.random() < 0.02:
                deps.append("not-in-queue")
            if deps and rnd.random() < 0.1:
                deps.append(deps[0])  # a repeated dependency: parallel lines in the multigraph
            grp = rnd.random() < 0.3
            items.append({"id": str(k), "dependencies": deps, "group": f"g{rnd.randrange(6)}" if grp else "", "build_variant": f"bv{rnd.randrange(2)}",
                          "project": "p", "version": f"v{rnd.randrange(3)}", "group_index": rnd.randrange(-2, 9)})
        batches.append(items)
    res = S.rebuild_dag_dispatchers([_tq(b) for b in batches], engine=engine)
    for items, (order, n_cycles, units) in zip(batches, res):
        want_order, want_cycles, want_units = OD.rebuild(items)
        assert order == want_order and n_cycles == len(want_cycles)
        assert units == want_units
    assert any(n for _, n, _ in res)  # random edges in both directions do form cycles somewhere
