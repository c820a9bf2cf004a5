#!/usr/bin/env python3
"""Write tests/golden/planner_kats.json: the known answers the reference's own
planner / queue-info / grouping tests assert, transcribed by hand (the Go tests
are not mechanically parseable: nested t.Run closures, helper lambdas).

Each case cites the reference test it restates.  Task dicts use the field names
of evergreen_b200.model.Task; `*_ago` fields are durations before `now` (the
reference builds them with time.Now().Add(-x)); the loader adds 1 us of elapsed
test time, because the reference reads the clock again inside Unit.info
(planner.go:318-322) and e.g. planner_test.go:250 (178, not 179) relies on it.
"""
import json
import os

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "planner_kats.json")
S, MIN, H = 10 ** 9, 60 * 10 ** 9, 3600 * 10 ** 9
P = "scheduler/planner_test.go"

PATCH, GHPR, MERGE, REPOTRACKER = "patch_request", "github_pull_request", "github_merge_request", "gitter_request"

# ---- unit.sortingValueBreakdown known answers (RankExpectedValues, planner_test.go:199-396) ----
unit_values = [
    dict(name="SingleTask", ref=f"{P}:200-205", tasks=[dict(id="foo")], total=180),
    dict(name="MultipleTasks", ref=f"{P}:206-212", tasks=[dict(id="foo"), dict(id="bar")], total=181),
    dict(name="MergeQueue", ref=f"{P}:213-218", tasks=[dict(id="foo", requester=MERGE)], total=2413),
    dict(name="Patches/CLI", ref=f"{P}:220-226", tasks=[dict(id="foo", requester=PATCH)], settings=dict(patch_factor=10), total=22),
    dict(name="Patches/Github", ref=f"{P}:227-233", tasks=[dict(id="foo", requester=GHPR)], settings=dict(patch_factor=10), total=22),
    dict(name="Priority", ref=f"{P}:235-240", tasks=[dict(id="foo", priority=10)], total=1970),
    dict(name="TimeInQueuePatch", ref=f"{P}:241-246", tasks=[dict(id="foo", requester=PATCH, activated_ago=H)], total=73),
    dict(name="TimeInQueueMainline", ref=f"{P}:247-252", tasks=[dict(id="foo", requester=REPOTRACKER, activated_ago=H)], total=178),
    dict(name="LifeTimePatch", ref=f"{P}:253-258", tasks=[dict(id="foo", requester=PATCH, ingest_ago=10 * H)], total=613),
    dict(name="LifeTimeMainlineNew", ref=f"{P}:259-264", tasks=[dict(id="foo", requester=REPOTRACKER, ingest_ago=10 * MIN)], total=179),
    dict(name="LifeTimeMainlineOld", ref=f"{P}:265-270", tasks=[dict(id="foo", requester=REPOTRACKER, ingest_ago=7 * 24 * H)], total=12),
    dict(name="NumDependents", ref=f"{P}:271-276", tasks=[dict(id="foo", num_dependents=2)], total=182),
    dict(name="NumDependentsWithFactor", ref=f"{P}:277-283", tasks=[dict(id="foo", num_dependents=2)], settings=dict(num_dependents_factor=10), total=200),
    dict(name="NumDependentsWithFractionFactor", ref=f"{P}:284-290", tasks=[dict(id="foo", num_dependents=2)], settings=dict(num_dependents_factor=0.5), total=181),
    dict(name="NumDependentsInGroupedUnit", ref=f"{P}:291-322",
         tasks=[dict(id="build-debug", num_dependents=22, priority=99)] + [dict(id=f"test-task-{i}") for i in range(22)],
         fields=dict(num_dependents_impact=22, initial_priority_impact=100)),
    dict(name="GenerateTask", ref=f"{P}:381-387", tasks=[dict(id="foo", generate_task=True)], settings=dict(generate_task_factor=10), total=1791),
    dict(name="TaskGroup", ref=f"{P}:388-395", tasks=[dict(id=i, task_group="tg1") for i in ("foo", "bar", "baz")], total=719),
    dict(name="RankCachesValue", ref=f"{P}:397-406", tasks=[dict(id="foo", priority=100)], total=18080),
]

# ---- whole-plan cases: PrepareTasksForPlanning(...).Export ----
dep = lambda *ids: [dict(task_id=i) for i in ids]
plans = [
    dict(name="TaskPlan/NoChange", ref=f"{P}:416-422", tasks=[dict(id="foo"), dict(id="bar")], order=["foo", "bar"]),
    dict(name="TaskPlan/ChangeOrder", ref=f"{P}:423-429", tasks=[dict(id="foo"), dict(id="bar", priority=10)], order=["bar", "foo"]),
    dict(name="PrepareTaskPlan/Noop", ref=f"{P}:484-486", tasks=[], n_units=0, order=[]),
    dict(name="PrepareTaskPlan/TaskGroupsGrouped", ref=f"{P}:487-496",
         tasks=[dict(id="one", task_group="first"), dict(id="two", task_group="first"), dict(id="three")], n_units=2, n_out=3),
    dict(name="PrepareTaskPlan/VersionsGrouped", ref=f"{P}:497-510", group_versions=True,
         tasks=[dict(id="one", version="first"), dict(id="two", version="first"), dict(id="three", version="second")],
         n_units=2, n_out=3),
    dict(name="PrepareTaskPlan/VersionsAndTaskGroupsGrouped", ref=f"{P}:511-530", group_versions=True,
         tasks=[dict(id="three", version="second"), dict(id="four", version="second"), dict(id="five", version="second"),
                dict(id="one", version="first", task_group="one"), dict(id="two", version="first", task_group="one"),
                dict(id="extra", version="first", priority=1)],
         n_units=3, n_out=6, head_task_groups=["one", "one"]),
    dict(name="PrepareTaskPlan/DependenciesGrouped", ref=f"{P}:531-548",
         tasks=[dict(id="one", depends_on=dep("two")), dict(id="three"), dict(id="two"), dict(id="other", depends_on=dep("two"))],
         n_units=4, n_out=4, last="three", head_set=["one", "two", "other"]),
    dict(name="PrepareTaskPlan/ExternalDependenciesIgnored", ref=f"{P}:549-558",
         tasks=[dict(id="one", depends_on=dep("missing")), dict(id="three"), dict(id="two", depends_on=dep("missing"))],
         n_units=3, n_out=3),
    dict(name="DependencyTaskScheduledFirst", ref=f"{P}:324-380", group_versions=True,
         tasks=[dict(id="build-debug", version="v1", num_dependents=20, priority=99, activated_ago=10 * MIN),
                dict(id="independent-test", version="v1", activated_ago=10 * MIN)] +
               [dict(id=f"test-kube-{i}", version="v1", depends_on=dep("build-debug"), activated_ago=10 * MIN) for i in range(20)],
         before=[["build-debug", "independent-test"]]),
    dict(name="TestDistroAliases/VerifyPrimaryQueue/Tunable", ref="scheduler/distro_alias_test.go:22-61",
         tasks=[dict(id="other", distro_id="one", priority=200, version="foo"),
                dict(id="one", distro_id="one", priority=2000, version="foo")], distro_id="one", order=["one", "other"]),
    dict(name="TestDistroAliases/DistroAlias/Tunable", ref="scheduler/distro_alias_test.go:22-36,93-112",
         tasks=[dict(id="other", distro_id="one", priority=200, version="foo"),
                dict(id="one", distro_id="one", priority=2000, version="foo")], distro_id="two", order=["one", "other"],
         secondary_queue=True),
]

# ---- TaskList comparator (planner_test.go:435-481): one unit, in-unit order ----
task_lists = [
    dict(name="TaskList/NoChange", ref=f"{P}:436-443", tasks=[dict(id="second"), dict(id="first")], order=["second", "first"]),
    dict(name="TaskList/TaskGroupOrder", ref=f"{P}:444-452",
         tasks=[dict(id="second", task_group_order=2), dict(id="first", task_group_order=1)], order=["first", "second"]),
    dict(name="TaskList/NumDependents", ref=f"{P}:453-459", tasks=[dict(id="second"), dict(id="first", num_dependents=2)], order=["first", "second"]),
    dict(name="TaskList/Priority", ref=f"{P}:460-465", tasks=[dict(id="second"), dict(id="first", priority=100)], order=["first", "second"]),
    dict(name="TaskList/ExpectedDuration", ref=f"{P}:466-480",
         tasks=[dict(id="second", prediction=dict(value=MIN, ttl=24 * H, collected_ago=0)),
                dict(id="first", prediction=dict(value=H, ttl=24 * H, collected_ago=0))], order=["first", "second"]),
]

# ---- GetDistroQueueInfo (task_queue_persister_test.go:33-122,203-204) ----
T = "scheduler/task_queue_persister_test.go"
queue_infos = [
    dict(name="persister/distroQueueInfo1", ref=f"{T}:33-119",
         tasks=[dict(id=f"t{i + 1}", build_variant=f"bv{i + 1}", requester=f"r{i + 1}", project=f"p{i + 1}",
                     activated_by=f"u{i + 1}", prediction=dict(value=(i + 1) * MIN)) for i in range(3)],
         distro_id="", threshold=30 * MIN, length=3, length_with_dependencies_met=3,
         expected_durations=[MIN, 2 * MIN, 3 * MIN]),
    dict(name="persister/distroQueueInfo2", ref=f"{T}:33-122,203-204",
         tasks=[dict(id="t4", build_variant="bv4", requester="r4", project="p4", activated_by="u4", prediction=dict(value=4 * MIN)),
                dict(id="t5", build_variant="bv5", requester="r5", project="p5", activated_by="u5",
                     depends_on=[dict(task_id="someTask", status="success")])],
         distro_id="", threshold=30 * MIN, length=2, length_with_dependencies_met=1,
         expected_durations=[4 * MIN, 10 * MIN]),
]

# ---- groupByTaskGroup (utilization_based_host_allocator_test.go:20-124) ----
A = "scheduler/utilization_based_host_allocator_test.go"
g2 = "g2___"
group_by = [
    dict(name="NoTaskGroups", ref=f"{A}:25-53", hosts=[dict(id="host1"), dict(id="host2")],
         infos=[dict(name="", count=2, max_hosts=1, expected_duration=2 * MIN)],
         buckets={"": dict(hosts=["host1", "host2"], count=2)}),
    dict(name="SomeRunningTaskGroups", ref=f"{A}:57-96",
         hosts=[dict(id="h1", running_task_group="g1", running_task="foo"), dict(id="h2", running_task_group="g1", running_task="bar")],
         infos=[dict(name=g2, count=1), dict(name="", count=1)],
         buckets={"g1___": dict(hosts=["h1", "h2"], count=0), g2: dict(hosts=[], count=1), "": dict(hosts=[], count=1)}),
    dict(name="SomeFinishedTaskGroups", ref=f"{A}:100-123",
         hosts=[dict(id="h1", running_task_group="g1"), dict(id="h2", running_task_group="g1")],
         infos=[dict(name=g2, count=1), dict(name="", count=1)],
         buckets={g2: dict(hosts=[], count=1), "": dict(hosts=["h1", "h2"], count=1)}),
]

json.dump(dict(now=1_800_000_000 * 10 ** 9, elapsed_ns=1000, unit_values=unit_values, plans=plans, task_lists=task_lists,
               queue_infos=queue_infos, group_by=group_by), open(OUT, "w"), indent=1)
print("wrote", OUT)
