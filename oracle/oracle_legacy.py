"""CPU restatement of the reference's LEGACY comparator prioritiser -- test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(evergreen_b200/) never does.

Follows, statement by statement (paths relative to the evergreen-ci/evergreen checkout):
  scheduler/task_prioritizer.go:80-142   CmpBasedTaskPrioritizer.PrioritizeTasks (lists in the order repotracker,
                                         patch, high priority; setup, sort.Stable, merge)
  scheduler/task_prioritizer.go:159-184  taskMoreImportantThan: first definitive comparator
  scheduler/task_prioritizer.go:214-247  splitTasksByRequester
  scheduler/task_prioritizer.go:251-278  mergeTasks
  scheduler/task_priority_cmp.go:25-208  the seven comparators
  scheduler/setup_funcs.go:72-87         groupTaskGroups (reverse-lexical presort)
and Go's sort.Stable itself (src/sort/zsortinterface.go: insertion-sorted blocks of 20, then symMerge), because the
comparator chain is not a strict weak order in general and only the exact algorithm reproduces the reference there.

Pinned on the known answers the reference's own tests hold: task_priority_cmp_test.go:46-197 (comparator truth
tables), :216-273 (byTaskGroupOrder), :340-343, :406-413 (orders), :437-454, :469-483 (byGenerateTasks,
byCommitQueue), task_prioritizer_test.go:164-196 (split), :223-327 (merge) -- tests/golden/legacy_prioritizer.json.
PARITY UNPINNED beyond those: no large ordering of the Go implementation could be generated here (no Go toolchain).
cacheExpectedDurations (setup_funcs.go:20-67) reorders the list by goroutine completion, which is not reproducible;
like the planner oracle this restatement keeps the input order there (groupTaskGroups re-sorts the list by a key
that contains the unique task id, so the result does not depend on it).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

from evergreen_b200 import model as M


# ---- comparators: (t1, t2, versions) -> -1 / 0 / 1 (task_priority_cmp.go) ---------------------------------
def by_priority(t1: M.Task, t2: M.Task, versions) -> int:  # :25-36
    if t1.priority > t2.priority:
        return 1
    if t1.priority < t2.priority:
        return -1
    return 0


def by_num_deps(t1, t2, versions) -> int:  # :43-54
    if t1.num_dependents > t2.num_dependents:
        return 1
    if t1.num_dependents < t2.num_dependents:
        return -1
    return 0


def tasks_are_commit_builds(t1, t2) -> bool:  # :201-207
    return t1.requester in M.SYSTEM_VERSION_REQUESTER_TYPES and t2.requester in M.SYSTEM_VERSION_REQUESTER_TYPES


def by_age(t1, t2, versions) -> int:  # :73-95
    if tasks_are_commit_builds(t1, t2) and t1.project == t2.project:
        if t1.revision_order_number > t2.revision_order_number:
            return 1
        if t1.revision_order_number < t2.revision_order_number:
            return -1
        return 0
    if t1.ingest_time < t2.ingest_time:  # time.Before
        return 1
    if t2.ingest_time < t1.ingest_time:
        return -1
    return 0


def expected_average(t: M.Task, now: Optional[int]) -> int:
    """Task.FetchExpectedDuration(ctx).Average as the planner oracle resolves it (model/task/task.go:3519-3590)."""
    return M.fetch_expected_duration(t, now)[0] if now is not None else t.expected_duration


def make_by_runtime(now: Optional[int]) -> Callable:
    def by_runtime(t1, t2, versions) -> int:  # :104-123
        one, two = expected_average(t1, now), expected_average(t2, now)
        if one == 0 or two == 0:
            return 0
        if one == two:
            return 0
        return 1 if one > two else -1
    return by_runtime


def by_task_group_order(t1, t2, versions) -> int:  # :132-169
    if t1.task_group == "" and t2.task_group == "":
        return 0
    if t2.task_group == "" and t1.task_group != "":
        return 1
    if t1.task_group == "" and t2.task_group != "":
        return -1
    if t1.task_group == t2.task_group and t1.build_id == t2.build_id:
        if t1.task_group_order > t2.task_group_order:
            return -1
        if t2.task_group_order > t1.task_group_order:
            return 1
    if f"{t1.build_id}-{t1.task_group}" < f"{t2.build_id}-{t2.task_group}":
        return 1
    return -1


def by_generate_tasks(t1, t2, versions) -> int:  # :175-185
    if t1.generate_task == t2.generate_task:
        return 0
    return 1 if t1.generate_task else -1


def by_commit_queue(t1, t2, versions) -> int:  # :191-204
    r1 = versions.get(t1.version, "") if versions else ""
    r2 = versions.get(t2.version, "") if versions else ""
    if r1 == M.GITHUB_MERGE_REQUESTER and r2 != M.GITHUB_MERGE_REQUESTER:
        return 1
    if r1 != M.GITHUB_MERGE_REQUESTER and r2 == M.GITHUB_MERGE_REQUESTER:
        return -1
    return 0


def default_comparators(now: Optional[int] = None) -> List[Callable]:  # task_prioritizer.go:59-67
    return [by_task_group_order, by_commit_queue, by_priority, by_num_deps, by_generate_tasks, by_age, make_by_runtime(now)]


def task_more_important_than(t1, t2, versions, comparators) -> bool:  # task_prioritizer.go:159-184
    for cmp in comparators or []:
        r = cmp(t1, t2, versions)
        if r == -1:
            return False
        if r == 1:
            return True
    return False


# ---- Go's sort.Stable over a Python list with a `less(i, j)` on positions ----------------------------------
def go_stable_sort(data: list, less_items: Callable) -> None:
    """sort.Stable (Go 1.19+ src/sort/zsortinterface.go stable/symMerge/rotate/insertionSort), in place."""
    def less(i, j):
        return less_items(data[i], data[j])

    def swap(i, j):
        data[i], data[j] = data[j], data[i]

    def insertion_sort(a, b):
        for i in range(a + 1, b):
            j = i
            while j > a and less(j, j - 1):
                swap(j, j - 1)
                j -= 1

    def swap_range(a, b, n):
        for i in range(n):
            swap(a + i, b + i)

    def rotate(a, m, b):
        i, j = m - a, b - m
        while i != j:
            if i > j:
                swap_range(m - i, m, j)
                i -= j
            else:
                swap_range(m - i, m + j - i, i)
                j -= i
        swap_range(m - i, m, i)

    def sym_merge(a, m, b):
        if m - a == 1:
            i, j = m, b
            while i < j:
                h = (i + j) >> 1
                if less(h, a):
                    i = h + 1
                else:
                    j = h
            for k in range(a, i - 1):
                swap(k, k + 1)
            return
        if b - m == 1:
            i, j = a, m
            while i < j:
                h = (i + j) >> 1
                if not less(m, h):
                    i = h + 1
                else:
                    j = h
            for k in range(m, i, -1):
                swap(k, k - 1)
            return
        mid = (a + b) >> 1
        n = mid + m
        if m > mid:
            start, r = n - b, mid
        else:
            start, r = a, m
        p = n - 1
        while start < r:
            c = (start + r) >> 1
            if not less(p - c, c):
                start = c + 1
            else:
                r = c
        end = n - start
        if start < m and m < end:
            rotate(start, m, end)
        if a < start and start < mid:
            sym_merge(a, start, mid)
        if mid < end and end < b:
            sym_merge(mid, end, b)

    n = len(data)
    block = 20
    a, b = 0, block
    while b <= n:
        insertion_sort(a, b)
        a = b
        b += block
    insertion_sort(a, n)
    while block < n:
        a, b = 0, 2 * block
        while b <= n:
            sym_merge(a, a + block, b)
            a = b
            b += 2 * block
        m = a + block
        if m < n:
            sym_merge(a, m, n)
        block *= 2


# ---- the prioritiser ---------------------------------------------------------------------------------------
def split_tasks_by_requester(tasks: Sequence[M.Task]) -> Tuple[List[M.Task], List[M.Task], List[M.Task]]:
    """task_prioritizer.go:214-247 -> (high priority, repotracker, patch); anything else is logged and dropped."""
    high, repo, patch = [], [], []
    for t in tasks:
        if t.priority > M.MAX_TASK_PRIORITY:
            high.append(t)
        elif t.requester in M.SYSTEM_VERSION_REQUESTER_TYPES:
            repo.append(t)
        elif M.is_patch_requester(t.requester):
            patch.append(t)
        elif t.requester == M.AD_HOC_REQUESTER:  # unreachable: ad_hoc is a system requester (globals.go:771)
            patch.append(t)
    return high, repo, patch


def group_task_groups(tasks: List[M.Task]) -> List[M.Task]:  # setup_funcs.go:72-87
    keyed = {f"{t.build_id}-{t.task_group}-{t.id}": t for t in tasks}
    keys = [f"{t.build_id}-{t.task_group}-{t.id}" for t in tasks]
    keys.sort(reverse=True)  # sort.Sort(sort.Reverse(sort.StringSlice)): byte-wise, like Python's str order on ASCII ids
    return [keyed[k] for k in keys]


def merge_tasks(high: List, repo: List, patch: List) -> List:  # task_prioritizer.go:251-278
    merged = list(high)
    r = p = 0
    for idx in range(len(repo) + len(patch)):
        if p >= len(patch):
            merged.append(repo[r]); r += 1
        elif r >= len(repo):
            merged.append(patch[p]); p += 1
        elif idx > 0 and (idx + 1) % 2 == 0:
            merged.append(repo[r]); r += 1
        else:
            merged.append(patch[p]); p += 1
    return merged


def prioritize_tasks(tasks: Sequence[M.Task], versions: Optional[Dict[str, str]] = None, now: Optional[int] = None,
                     comparators: Optional[List[Callable]] = None) -> List[M.Task]:
    """CmpBasedTaskPrioritizer.PrioritizeTasks (task_prioritizer.go:80-142).  `versions` maps a version id to its
    Requester (all byCommitQueue reads of model.Version)."""
    cmps = default_comparators(now) if comparators is None else comparators
    high, repo, patch = split_tasks_by_requester(tasks)
    out = []
    for lst in (repo, patch, high):
        lst = group_task_groups(lst)
        go_stable_sort(lst, lambda a, b: task_more_important_than(a, b, versions, cmps))
        out.append(lst)
    return merge_tasks(out[2], out[0], out[1])
