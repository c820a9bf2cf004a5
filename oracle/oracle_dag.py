"""CPU restatement of the DAG dispatcher's rebuild -- test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Follows model/task_queue_service_dependency.go:153-252 (basicCachedDAGDispatcherImpl.rebuild): one graph node per
TaskQueueItem with queueIndex = its position in the scheduler-sorted queue; task groups bucketed by
compositeGroupID(Group, BuildVariant, Project, Version) and each bucket sort.SliceStable'd by GroupIndex (:165-192);
an edge dependency -> item for every TaskQueueItem.Dependencies entry that is itself in the queue (addEdge :118-150);
then topo.SortStabilized(graph, order) with `order` = ascending queueIndex (:204-215).

topo.SortStabilized lives in gonum.org/v1/gonum v0.17.0 (go.mod:62), which is NOT under /root/reference.  Its published
algorithm (graph/topo/tarjan.go: tarjanSCCstabilized + sortedFrom) is restated here:
  * the node list is put in `order` and then REVERSED; the successors of a node likewise (order, then reverse);
  * Tarjan's strongly-connected-components algorithm runs over the nodes in that (descending) sequence and emits
    components in reverse topological order;
  * a component of more than one node is a dependency cycle: it contributes ONE nil placeholder to the result and is
    reported in the Unorderable error (the reference logs the cycle and keeps going, :216-247);
  * the emitted sequence is reversed.
For a DAG the result is the reverse post-order of a depth-first search that visits nodes and successors in
descending queueIndex.  Pinned on the one ordering the reference's tests hold: TestConstructor's expectedOrder
(model/task_queue_service_test.go:548-656; 100 items, two dependency chains 50->45->40->35 and 80->75->70->65) --
tests/golden/dag_dispatcher.json.  PARITY UNPINNED beyond it (no Go toolchain, gonum not vendored): in particular the
position of cycle placeholders rests on the restatement alone.
"""
from __future__ import annotations

import sys
from typing import Dict, List, Optional, Sequence, Tuple


def topo_sort_stabilized(n: int, succ: Sequence[Sequence[int]]) -> Tuple[List[Optional[int]], List[List[int]]]:
    """Nodes are 0..n-1 with key == id (queueIndex).  `succ[v]` = nodes that depend on v (any order, duplicates allowed:
    the multigraph keeps parallel lines but From() yields each neighbour once).  Returns (sorted with None for each
    cyclic component, cycles)."""
    succs = [sorted(set(s), reverse=True) for s in succ]  # order(to); ordered.Reverse(to)
    index = [0] * n
    low = [0] * n
    on_stack = [False] * n
    stack: List[int] = []
    sccs: List[List[int]] = []
    counter = 0
    sys.setrecursionlimit(max(10000, 4 * n + 100))

    def strongconnect(v: int) -> None:
        nonlocal counter
        counter += 1
        index[v] = low[v] = counter
        stack.append(v)
        on_stack[v] = True
        for w in succs[v]:
            if index[w] == 0:
                strongconnect(w)
                low[v] = min(low[v], low[w])
            elif on_stack[w]:
                low[v] = min(low[v], index[w])
        if low[v] == index[v]:
            scc = []
            while True:
                w = stack.pop()
                on_stack[w] = False
                scc.append(w)
                if w == v:
                    break
            sccs.append(scc)

    for v in range(n - 1, -1, -1):  # order(nodes); ordered.Reverse(nodes)
        if index[v] == 0:
            strongconnect(v)
    out: List[Optional[int]] = []
    cycles: List[List[int]] = []
    for scc in sccs:  # sortedFrom
        if len(scc) != 1:
            cycles.append(sorted(scc))
            out.append(None)
        else:
            out.append(scc[0])
    out.reverse()
    cycles.reverse()
    return out, cycles


def rebuild(items: Sequence[dict]):
    """items: TaskQueueItem-like dicts {id, group, build_variant, project, version, group_index, dependencies}.
    Returns (sorted ids with None placeholders, cycles as id lists, task groups {composite id: [item ids by GroupIndex]})."""
    pos = {it["id"]: k for k, it in enumerate(items)}
    succ: List[List[int]] = [[] for _ in items]
    for k, it in enumerate(items):
        for dep in it.get("dependencies", []):
            j = pos.get(dep)
            if j is not None:  # "the depend_on task is not in the DAG so we don't need an edge" (:123-126)
                succ[j].append(k)
    order, cycles = topo_sort_stabilized(len(items), succ)
    groups: Dict[str, List[int]] = {}
    for k, it in enumerate(items):
        if it.get("group", ""):
            gid = f'{it["group"]}_{it.get("build_variant", "")}_{it.get("project", "")}_{it.get("version", "")}'  # compositeGroupID
            groups.setdefault(gid, []).append(k)
    for gid in groups:
        groups[gid].sort(key=lambda k: items[k].get("group_index", 0))  # sort.SliceStable by GroupIndex
    return ([None if v is None else items[v]["id"] for v in order], [[items[v]["id"] for v in c] for c in cycles],
            {g: [items[k]["id"] for k in ks] for g, ks in groups.items()})
