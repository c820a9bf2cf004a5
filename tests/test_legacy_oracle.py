"""The legacy comparator prioritiser's oracle (oracle/oracle_legacy.py) against every known answer the reference's own
tests hold for it (tests/golden/legacy_prioritizer.json, transcribed by make_legacy_golden.py), plus properties of the
Go sort.Stable port.  CPU only."""
import functools
import json
import os
import random

import pytest

from evergreen_b200 import model as M
from oracle import oracle_legacy as OL

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "legacy_prioritizer.json")))
NOW = G["now"]


def mk(d):
    d = dict(d)
    dp = d.pop("duration_prediction", None)
    t = M.Task(**d)
    if dp:
        t.duration_prediction = M.CachedDurationValue(value=dp.get("value", 0), ttl=dp.get("ttl", 0))
    return t


CMPS = {"by_priority": OL.by_priority, "by_num_deps": OL.by_num_deps, "by_age": OL.by_age, "by_runtime": OL.make_by_runtime(NOW),
        "by_task_group_order": OL.by_task_group_order, "by_generate_tasks": OL.by_generate_tasks, "by_commit_queue": OL.by_commit_queue}


@pytest.mark.parametrize("case", G["comparators"], ids=lambda c: f'{c["comparator"]}@{c["ref"].split(":")[1]}')
def test_comparator_truth_tables(case):
    assert CMPS[case["comparator"]](mk(case["t1"]), mk(case["t2"]), case["versions"]) == case["want"], case["ref"]


@pytest.mark.parametrize("case", G["orders"], ids=lambda c: c["ref"].split()[-1])
def test_orders(case):
    got = [t.id for t in OL.prioritize_tasks([mk(t) for t in case["tasks"]], case["versions"], NOW)]
    if "want_order" in case:
        assert got == case["want_order"]
    for a, b in case.get("want_before", []):
        assert got.index(a) < got.index(b)
    assert sorted(got) == sorted(t["id"] for t in case["tasks"])


def test_first_definitive_comparator_wins():
    synth = {"always_equal": lambda a, b, v: 0, "always_more": lambda a, b, v: 1, "always_less": lambda a, b, v: -1,
             "id": lambda a, b, v: (a.id > b.id) - (a.id < b.id)}
    ts = {"t1": M.Task(id="t1"), "t2": M.Task(id="t2")}
    for case in G["chain"]:
        cmps = None if case["comparators"] is None else [synth[n] for n in case["comparators"]]
        for pair, want in case["want"].items():
            a, b = pair.split(",")
            assert OL.task_more_important_than(ts[a], ts[b], {}, cmps) == want, (case["ref"], pair)


def test_split_and_merge():
    for case in G["splits"]:
        high, repo, patch = OL.split_tasks_by_requester([mk(t) for t in case["tasks"]])
        assert [t.id for t in repo] == case["want"]["repotracker"] and [t.id for t in patch] == case["want"]["patch"]
        assert [t.id for t in high] == case["want"]["high"], case["ref"]
    for case in G["merges"]:
        assert OL.merge_tasks(case["high"], case["repotracker"], case["patch"]) == case["want"], case["ref"]
    # an unrecognised requester is dropped (task_prioritizer.go:232-240); ad_hoc is a system requester (globals.go:771)
    high, repo, patch = OL.split_tasks_by_requester([M.Task(id="a", requester="nonsense"), M.Task(id="b", requester=M.AD_HOC_REQUESTER)])
    assert not high and [t.id for t in repo] == ["b"] and not patch


def test_go_stable_sort_is_a_stable_sort():
    rnd = random.Random(5)
    for n in (0, 1, 2, 19, 20, 21, 40, 41, 257, 1000):
        data = [(rnd.randrange(7), k) for k in range(n)]
        want = sorted(data, key=lambda x: x[0])
        OL.go_stable_sort(data, lambda a, b: a[0] < b[0])
        assert data == want


def test_presort_puts_task_groups_first_and_is_reverse_lexical():
    ts = [M.Task(id=f"t{k}", build_id=f"b{k % 3}", task_group="g" if k % 2 else "") for k in range(12)]
    got = OL.group_task_groups(ts)
    keys = [f"{t.build_id}-{t.task_group}-{t.id}" for t in got]
    assert keys == sorted(keys, reverse=True)
