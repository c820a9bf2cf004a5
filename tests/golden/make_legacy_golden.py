#!/usr/bin/env python3
"""Known answers the reference's tests hold for the LEGACY comparator prioritiser, transcribed by hand (each with the
file:line of the assertion) -> tests/golden/legacy_prioritizer.json.

  scheduler/task_priority_cmp_test.go   comparator truth tables, two end-to-end orders
  scheduler/task_prioritizer_test.go    first-definitive chain, split by requester, merge

Task fields use evergreen_b200.model.Task's names; times / durations are ns.  `now` is only used to resolve
FetchExpectedDuration's defaults (no cached prediction, no history -> 10 min)."""
import json
import os

MIN, HOUR = 60 * 10 ** 9, 3600 * 10 ** 9
NOW = 1_800_000_000 * 10 ** 9
REPO, PATCH, MERGE = "gitter_request", "patch_request", "github_merge_request"

cmp_cases = []


def c(line, name, t1, t2, want, versions=None):
    cmp_cases.append({"ref": f"task_priority_cmp_test.go:{line}", "comparator": name, "t1": dict(id="t1", **t1),
                      "t2": dict(id="t2", **t2), "want": want, "versions": versions or {}})


# byPriority :46-68
c(54, "by_priority", {"priority": 2}, {"priority": 2}, 0)
c(60, "by_priority", {"priority": 2}, {"priority": 1}, 1)
c(66, "by_priority", {"priority": 1}, {"priority": 2}, -1)
# byRuntime :70-104 (no durations at all: both default to 10 min)
c(75, "by_runtime", {}, {}, 0)
c(82, "by_runtime", {"expected_duration": 20 * MIN}, {"expected_duration": HOUR}, -1)
c(86, "by_runtime", {"expected_duration": HOUR}, {"expected_duration": 20 * MIN}, 1)
c(90, "by_runtime", {"expected_duration": 20 * MIN}, {"expected_duration": 20 * MIN}, 0)
c(98, "by_runtime", {"expected_duration": 1, "duration_prediction": {"value": 1, "ttl": HOUR}}, {"expected_duration": HOUR}, -1)
c(102, "by_runtime", {"expected_duration": HOUR}, {"expected_duration": 1, "duration_prediction": {"value": 1, "ttl": HOUR}}, 1)
# byNumDeps :106-129
c(112, "by_num_deps", {}, {}, 0)
c(118, "by_num_deps", {"num_dependents": 1}, {}, 1)
c(121, "by_num_deps", {}, {"num_dependents": 1}, -1)
c(127, "by_num_deps", {"num_dependents": 1}, {"num_dependents": 1}, 0)
# byAge, commit builds :131-161
c(141, "by_age", {"requester": REPO}, {"requester": REPO}, 0)
c(147, "by_age", {"requester": REPO, "revision_order_number": 1}, {"requester": REPO}, 1)
c(152, "by_age", {"requester": REPO}, {"requester": REPO, "revision_order_number": 1}, -1)
c(158, "by_age", {"requester": REPO, "revision_order_number": 1, "project": "project"}, {"requester": REPO}, 0)  # different projects, equal (zero) ingest times
# byAge, patches :163-197
c(172, "by_age", {"requester": PATCH}, {"requester": PATCH}, 0)
c(179, "by_age", {"requester": PATCH, "ingest_time": NOW}, {"requester": PATCH}, -1)
c(184, "by_age", {"requester": PATCH, "ingest_time": NOW, "project": "project"}, {"requester": PATCH}, -1)
c(188, "by_age", {"requester": PATCH}, {"requester": PATCH, "ingest_time": NOW, "project": "project"}, 1)
c(193, "by_age", {"requester": PATCH, "ingest_time": NOW, "project": "project"}, {"requester": PATCH, "ingest_time": NOW}, 0)
# byTaskGroupOrder :216-273
TG = "example_task_group"
c(218, "by_task_group_order", {}, {}, 0)
c(222, "by_task_group_order", {"task_group": TG}, {}, 1)
c(227, "by_task_group_order", {}, {"task_group": TG}, -1)
c(234, "by_task_group_order", {"task_group": TG}, {"task_group": "another_task_group"}, -1)
c(243, "by_task_group_order", {"task_group": TG, "build_id": "build_id"}, {"task_group": TG, "build_id": "another_build_id"}, -1)
c(265, "by_task_group_order", {"task_group": TG, "build_id": "build_id", "version": "version_id", "task_group_order": 1},
  {"task_group": TG, "build_id": "build_id", "version": "version_id", "task_group_order": 2}, 1)
c(272, "by_task_group_order", {"task_group": TG, "build_id": "build_id", "version": "version_id", "task_group_order": 2},
  {"task_group": TG, "build_id": "build_id", "version": "version_id", "task_group_order": 1}, -1)
# byGenerateTasks :437-454
c(439, "by_generate_tasks", {"generate_task": True}, {"generate_task": True}, 0)
c(442, "by_generate_tasks", {}, {}, 0)
c(445, "by_generate_tasks", {"generate_task": True}, {}, 1)
c(451, "by_generate_tasks", {}, {"generate_task": True}, -1)
# byCommitQueue :469-483
V = {"v0": MERGE, "v1": PATCH}
c(471, "by_commit_queue", {"version": "v0"}, {"version": "v0"}, 0, V)
c(475, "by_commit_queue", {"version": "v1"}, {"version": "v1"}, 0, V)
c(479, "by_commit_queue", {"version": "v0"}, {"version": "v1"}, 1, V)
c(483, "by_commit_queue", {"version": "v1"}, {"version": "v0"}, -1, V)

orders = [
    {"ref": "task_priority_cmp_test.go:340-343 TestPrioritizeTasksWithSameTaskGroupsAndDifferentBuilds",
     "versions": {"version_1": "", "version_2": ""},
     "tasks": [
         {"id": "task_1", "build_id": "build_1", "display_name": "another_task", "version": "version_1", "requester": PATCH, "task_group": TG, "task_group_order": 2},
         {"id": "task_2", "build_id": "build_2", "display_name": "first_task", "version": "version_1", "requester": PATCH, "task_group": TG, "task_group_order": 1},
         {"id": "task_3", "build_id": "build_2", "display_name": "another_task", "version": "version_1", "requester": PATCH, "task_group": TG, "task_group_order": 2},
         {"id": "task_4", "build_id": "build_1", "display_name": "first_task", "version": "version_1", "requester": PATCH, "task_group": TG, "task_group_order": 1}],
     "want_order": ["task_4", "task_1", "task_2", "task_3"]},
    {"ref": "task_priority_cmp_test.go:400-413 TestTaskGroupsNotOutOfOrderFromOtherComparators",
     "versions": {"version_1": "", "version_2": ""},
     "tasks": [
         {"id": "task_1", "build_id": "build_1", "display_name": "later_task", "version": "version_1", "requester": PATCH, "task_group_order": 2, "task_group": TG, "priority": 4},
         {"id": "task_3", "build_id": "build_1", "display_name": "third_task", "version": "version_1", "requester": PATCH, "priority": 1},
         {"id": "task_2", "build_id": "build_1", "display_name": "earlier_task", "version": "version_1", "requester": PATCH, "task_group": TG, "task_group_order": 1, "priority": 0}],
     "want_before": [["task_2", "task_1"]]},  # the assertion: earlier_task is reached before later_task
]

chain = [  # task_prioritizer_test.go:66-147: taskMoreImportantThan with synthetic comparators over ids t1, t2
    {"ref": "task_prioritizer_test.go:70-80", "comparators": None, "want": {"t1,t2": False, "t2,t1": False}},
    {"ref": "task_prioritizer_test.go:84-96", "comparators": [], "want": {"t1,t2": False, "t2,t1": False}},
    {"ref": "task_prioritizer_test.go:100-114", "comparators": ["always_more"], "want": {"t1,t2": True, "t2,t1": True}},
    {"ref": "task_prioritizer_test.go:118-147", "comparators": ["always_equal", "id", "always_more", "always_less"],
     "want": {"t1,t2": False, "t2,t1": True, "t1,t1": True, "t2,t2": True}},
]

splits = [
    {"ref": "task_prioritizer_test.go:153-175",
     "tasks": [{"id": "t1", "requester": REPO}, {"id": "t2", "requester": PATCH}, {"id": "t3", "requester": PATCH},
               {"id": "t4", "requester": REPO}, {"id": "t5", "requester": REPO}],
     "want": {"repotracker": ["t1", "t4", "t5"], "patch": ["t2", "t3"], "high": []}},
    {"ref": "task_prioritizer_test.go:176-198",
     "tasks": [{"id": "t1", "requester": REPO, "priority": 101}, {"id": "t2", "requester": PATCH, "priority": 101},
               {"id": "t3", "requester": PATCH}, {"id": "t4", "requester": REPO}, {"id": "t5", "requester": REPO}],
     "want": {"repotracker": ["t4", "t5"], "patch": ["t3"], "high": ["t1", "t2"]}},
]

merges = [
    {"ref": "task_prioritizer_test.go:223-231", "high": [], "repotracker": ["t1", "t2", "t3"], "patch": [], "want": ["t1", "t2", "t3"]},
    {"ref": "task_prioritizer_test.go:239-251", "high": ["t4", "t5"], "repotracker": ["t1", "t2", "t3"], "patch": [], "want": ["t4", "t5", "t1", "t2", "t3"]},
    {"ref": "task_prioritizer_test.go:262-270", "high": [], "repotracker": [], "patch": ["t1", "t2", "t3"], "want": ["t1", "t2", "t3"]},
    {"ref": "task_prioritizer_test.go:283-292", "high": [], "repotracker": ["t1", "t2", "t3"], "patch": ["t4", "t5", "t6"], "want": ["t4", "t1", "t5", "t2", "t6", "t3"]},
    {"ref": "task_prioritizer_test.go:303-311", "high": [], "repotracker": ["t1", "t2"], "patch": ["t3", "t4", "t5", "t6"], "want": ["t3", "t1", "t4", "t2", "t5", "t6"]},
    {"ref": "task_prioritizer_test.go:320-327", "high": [], "repotracker": ["t1", "t2", "t3", "t4", "t5"], "patch": ["t6"], "want": ["t6", "t1", "t2", "t3", "t4", "t5"]},
]

out = {"now": NOW, "comparators": cmp_cases, "orders": orders, "chain": chain, "splits": splits, "merges": merges}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "legacy_prioritizer.json")
json.dump(out, open(path, "w"), indent=1)
print(path, len(cmp_cases), "comparator cases")
