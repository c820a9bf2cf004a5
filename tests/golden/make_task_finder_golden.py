#!/usr/bin/env python3
"""Write tests/golden/task_finder.json: what the reference's task-finder tests assert
(scheduler/task_finder_test.go), transcribed by hand -- the Go tests insert into MongoDB and cannot run here.

`asserted` cases carry the exact expectation of the Go test (a length, and ids where the test checks them) and
hold for every finder the suite is instantiated with (legacy, alternate, parallel; :42-67).  `equivalence`
cases are the fixtures of TaskFinderComparisonSuite (:309-358), which asserts only that all finders return the
same set; `expect_ids` there is derived from the finder source and marked so.

Task / project-ref dicts use the field names of evergreen_b200.model; Go zero values apply to omitted fields
(a Go task.Task{} has Activated=false and Status="").
"""
import json
import os

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "task_finder.json")
F = "scheduler/task_finder_test.go"
UND, OK, FAIL = "undispatched", "success", "failed"


def base_tasks():  # SetupTest :70-84
    ts = [dict(id=f"t{i}", status=UND, activated=True, project="exists") for i in range(6)]
    ts[5]["priority"] = -1
    return ts


def go_task(**kw):  # a Go struct literal: zero values unless given
    d = dict(status="", activated=False)
    d.update(kw)
    return d


refs_enabled = [dict(id="exists", enabled=True)]
cases = []

cases.append(dict(name="NoRunnableTasksReturnsEmptySlice", ref=f"{F}:113-118", kind="asserted",
                  tasks=[], project_refs=refs_enabled, expect_len=0))

t = base_tasks(); t[4]["activated"] = False
cases.append(dict(name="InactiveTasksNeverReturned", ref=f"{F}:120-129", kind="asserted",
                  tasks=t + [go_task(id="td1"), go_task(id="td2")], project_refs=refs_enabled, expect_len=4))

deps = [go_task(id="td1"), go_task(id="td2")]
cases.append(dict(name="FilterTasksWhenValidProjectsSet/default", ref=f"{F}:131-136", kind="asserted",
                  tasks=base_tasks() + deps, project_refs=refs_enabled, expect_len=5))
cases.append(dict(name="FilterTasksWhenValidProjectsSet/listed", ref=f"{F}:138-146", kind="asserted",
                  tasks=base_tasks() + deps, project_refs=refs_enabled, valid_projects=["exists"], expect_len=5))
t = base_tasks(); t[0]["project"] = "something_else"; t[1]["project"] = "something_else"
cases.append(dict(name="FilterTasksWhenValidProjectsSet/other-project", ref=f"{F}:148-156", kind="asserted",
                  tasks=t + deps, project_refs=refs_enabled, valid_projects=["exists"], expect_len=3))

t = base_tasks()
td1 = go_task(id="td1", status=FAIL)
td2 = go_task(id="td2", status=UND, depends_on=[dict(task_id="none", status="*", unattainable=True)])
t[0]["depends_on"] = [dict(task_id="td1", status=FAIL)]                                   # matching - runnable
t[1]["depends_on"] = [dict(task_id="td1", status=OK)]                                     # not matching
t[2]["depends_on"] = [dict(task_id="td2", status="*"), dict(task_id="td1", status="*")]   # blocked + "*" - runnable
t[3]["depends_on"] = [dict(task_id="td1", status="*")]                                    # "*" matches any finished
cases.append(dict(name="TasksWithUnsatisfiedDependenciesNeverReturned", ref=f"{F}:159-191", kind="asserted",
                  tasks=t + [td1, td2], project_refs=refs_enabled, expect_len=4, expect_ids=["t0", "t2", "t3", "t4"]))

# the two project tests never insert tasks (no insertTasks call), so the reference only asserts "empty"
cases.append(dict(name="TasksWithDisabledProjectNeverReturned", ref=f"{F}:193-202", kind="asserted",
                  tasks=[], project_refs=[dict(id="exists", enabled=False)], expect_len=0))
cases.append(dict(name="TasksWithProjectDispatchingDisabledNeverReturned", ref=f"{F}:204-213", kind="asserted",
                  tasks=[], project_refs=[dict(id="exists", dispatching_disabled=True)], expect_len=0))
# the same two with the suite's tasks present: what the test names promise (derived from ProjectCanDispatchTask)
cases.append(dict(name="TasksWithDisabledProjectNeverReturned/with-tasks", ref=f"{F}:193-202 + model/project_ref.go:3441-3451",
                  kind="derived", tasks=base_tasks(), project_refs=[dict(id="exists", enabled=False)], expect_len=0))
cases.append(dict(name="TasksWithProjectDispatchingDisabledNeverReturned/with-tasks", ref=f"{F}:204-213 + model/project_ref.go:3453-3455",
                  kind="derived", tasks=base_tasks(), project_refs=[dict(id="exists", enabled=True, dispatching_disabled=True)],
                  expect_len=0))

# TaskFinderComparisonSuite project refs (SetupSuite :228-262) and the static fixture (:368-479)
suite_refs = [dict(id="exists", enabled=True), dict(id="disabled", enabled=False),
              dict(id="patching-disabled", enabled=True, patching_disabled=True),
              dict(id="dispatching-disabled", enabled=True, dispatching_disabled=True)]
static = [
    go_task(id="parent0", status=OK, activated=True),
    go_task(id="parent0-child0", status=UND, activated=True, depends_on=[dict(task_id="parent0", status=FAIL)]),
    go_task(id="parent0-child1", status=UND, activated=True, depends_on=[dict(task_id="parent0", status=OK)]),
    go_task(id="parent1", status=FAIL, activated=True),
    go_task(id="parent1-child1-child1", status=UND, activated=True, depends_on=[dict(task_id="parent1", status=FAIL)]),
    go_task(id="parent0+parent1-child0", status=UND, activated=True,
            depends_on=[dict(task_id="parent0", status=OK), dict(task_id="parent1", status=FAIL)]),
    go_task(id="parent2", status=UND, activated=True),
    go_task(id="foo", status=UND, activated=True, project="disabled"),
    go_task(id="bar", status=UND, activated=True, requester="patch_request", project="patching-disabled"),
    go_task(id="baz", status=UND, activated=True, requester="github_pull_request", project="patching-disabled"),
    go_task(id="runnable", status=UND, activated=True, requester="gitter_request", project="patching-disabled"),
    go_task(id="also-runnable", status=UND, activated=True, requester="gitter_request", project="dispatching-disabled"),
]
# tasks with Project "" have no project ref and are skipped by every finder; "also-runnable" sits in a project with
# dispatching disabled, which ProjectCanDispatchTask refuses (project_ref.go:3453-3455)
cases.append(dict(name="CompareTaskRunnersWithStaticTasks", ref=f"{F}:368-479,309-334", kind="equivalence",
                  tasks=static, project_refs=suite_refs, expect_ids=["runnable"], expect_ids_derived=True))
huge = [go_task(id="hugedeps", status=UND, activated=True, project="exists",
                depends_on=[dict(task_id=f"task{i}", status=OK) for i in range(5)])]
huge += [go_task(id=f"task{i}", status=OK, activated=True, project="exists") for i in range(5)]
huge += [go_task(id="skipped-project00", status=UND, activated=True, project="doesn't exist")]
cases.append(dict(name="CompareTaskRunnersWithHugeTasks", ref=f"{F}:609-650,309-334", kind="equivalence",
                  tasks=huge, project_refs=suite_refs, expect_ids=["hugedeps"], expect_ids_derived=True))

json.dump(dict(source=F, cases=cases), open(OUT, "w"), indent=1)
print("wrote", OUT, len(cases), "cases")
