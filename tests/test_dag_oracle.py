"""oracle/oracle_dag.py (restatement of the DAG dispatcher's rebuild and of gonum's topo.SortStabilized) against the one
ordering the reference's tests hold for it, plus structural properties.  CPU only."""
import json
import os
import random

from oracle import oracle_dag as OD

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "dag_dispatcher.json")))


def test_reference_expected_order():
    order, cycles, groups = OD.rebuild(G["items"])
    assert order == G["expected_order"] and cycles == []
    assert len(groups) == G["n_task_groups"] and {len(v) for v in groups.values()} == {G["group_size"]}


def test_cycles_and_self_edges():
    # TestDependencyCycle (model/task_queue_service_test.go:686-714): t0 <-> t1 form a cycle, t2 stays dispatchable
    items = [{"id": "t0", "dependencies": ["t1"]}, {"id": "t1", "dependencies": ["t0"]}, {"id": "t2", "dependencies": []}]
    order, cycles, _ = OD.rebuild(items)
    assert order.count(None) == 1 and "t2" in order and cycles == [["t0", "t1"]]
    # TestSelfEdge (:659-684): a self edge is a one-node component, not a cycle for Tarjan
    order, cycles, _ = OD.rebuild([{"id": "t0", "dependencies": ["t0"]}])
    assert order == ["t0"] and cycles == []
    # a dependency that is not in the queue adds no edge (addEdge :123-126)
    order, _, _ = OD.rebuild([{"id": "a", "dependencies": ["zzz"]}, {"id": "b", "dependencies": []}])
    assert order == ["a", "b"]


def test_random_dags_are_topological_and_stable_for_roots():
    rnd = random.Random(4)
    for n in (1, 2, 10, 200, 2000):
        items = [{"id": str(k), "dependencies": [str(rnd.randrange(k + 1, n)) for _ in range(rnd.choice([0, 0, 1, 2])) if k + 1 < n]} for k in range(n)]
        order, cycles, _ = OD.rebuild(items)
        assert cycles == [] and sorted(order, key=int) == [str(k) for k in range(n)]
        at = {v: p for p, v in enumerate(order)}
        for it in items:
            for dep in it["dependencies"]:
                assert at[dep] < at[it["id"]]  # a dependency is dispatched before its dependent
        if all(not it["dependencies"] for it in items):
            assert order == [str(k) for k in range(n)]  # no edges: the scheduler's order (TestFindNextTaskRespectsQueueOrderForRootTasks)


def test_task_groups_are_stably_sorted_by_group_index():
    items = [{"id": f"i{k}", "group": "g" if k % 2 else "", "build_variant": "bv", "project": "p", "version": "v", "group_index": (7 - k) // 3,
              "dependencies": []} for k in range(12)]
    _, _, groups = OD.rebuild(items)
    (ids,) = groups.values()
    keys = [(items[int(i[1:])]["group_index"], int(i[1:])) for i in ids]
    assert keys == sorted(keys)
