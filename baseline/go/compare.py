#!/usr/bin/env python3
"""Check the reference's own results (go_results.jsonl from harness_test.go) against the oracle -- and, with --gpu,
against libevgsched.so -- modulo ties: see README.md."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from evergreen_b200 import model as M  # noqa: E402


def load(path):
    return [json.loads(line) for line in open(path)]


def tasks_of(fx):
    now = fx["NowNs"]
    out = []
    for f in fx["Tasks"]:
        out.append(M.Task(
            id=f["Id"], version=f["Version"], project=f["Project"], build_variant=f["BuildVariant"], task_group=f["TaskGroup"],
            task_group_order=f["TaskGroupOrder"], task_group_max_hosts=f["TaskGroupMaxHosts"], priority=f["Priority"],
            requester=f["Requester"], generate_task=f["GenerateTask"], activated_by=f["ActivatedBy"], num_dependents=f["NumDependents"],
            distro_id=f["DistroId"], activated_time=now - f["ActivatedAgoNs"], scheduled_time=now - f["ScheduledAgoNs"],
            override_dependencies=f["OverrideDependencies"], expected_duration=f["ExpectedNs"],
            duration_prediction=M.CachedDurationValue(value=f["ExpectedNs"], ttl=24 * M.HOUR, collected_at=now),
            depends_on=[M.Dependency(d, status=M.TASK_SUCCEEDED) for d in f["DependsOn"]]))
    return out


def distro_of(fx):
    p = fx["Planner"]
    ps = M.PlannerSettings(target_time=int(p["target_time_ns"]), group_versions=bool(p["group_versions"]),
                           patch_factor=int(p["patch_factor"]), patch_time_in_queue_factor=int(p["patch_time_in_queue_factor"]),
                           commit_queue_factor=int(p["commit_queue_factor"]), mainline_time_in_queue_factor=int(p["mainline_time_in_queue_factor"]),
                           expected_runtime_factor=int(p["expected_runtime_factor"]), generate_task_factor=int(p["generate_task_factor"]),
                           num_dependents_factor=float(p["num_dependents_factor"]), stepback_task_factor=int(p["stepback_task_factor"]))
    ver = M.DISPATCHER_VERSION_REVISED_WITH_DEPENDENCIES if p["includes_dependencies"] else ""
    return M.Distro(id=fx["Distro"], provider=M.PROVIDER_EC2_FLEET, planner_settings=ps, dispatcher_settings=M.DispatcherSettings(ver))


def tie_runs(values):
    v = np.asarray(values)
    cuts = np.nonzero(np.diff(v))[0] + 1
    return np.split(np.arange(len(v)), cuts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("fixture")
    ap.add_argument("results")
    ap.add_argument("--gpu", action="store_true", help="also run the fixture through libevgsched.so")
    a = ap.parse_args()
    from oracle import oracle as O
    fixtures, results = load(a.fixture), load(a.results)
    assert len(fixtures) == len(results)
    bad = 0
    for fx, go in zip(fixtures, results):
        d, now = distro_of(fx), fx["NowNs"]
        tasks = tasks_of(fx)
        plan, breakdowns = O.plan(d, tasks, now)[:2]
        mine_ids = [t.id for t in plan]
        mine_v = [b.total_value for b in breakdowns]
        if mine_v != go["TotalValue"]:
            bad += 1
            print(f"{fx['Distro']}: TotalValue sequences differ")
            continue
        for run in tie_runs(mine_v):
            if {mine_ids[k] for k in run} != {go["Order"][k] for k in run}:
                bad += 1
                print(f"{fx['Distro']}: different task sets inside the tie run at rank {run[0]}")
                break
        qi = O.queue_info(d.id, tasks, d.get_target_time(), bool(fx["Planner"]["includes_dependencies"]), now)
        g = go["Info"]
        for mine, theirs in ((qi.length, g["Length"]), (qi.length_with_dependencies_met, g["LengthWithDependenciesMet"]),
                             (qi.expected_duration, g["ExpectedDuration"]), (qi.count_duration_over_threshold, g["CountDurationOverThreshold"]),
                             (qi.duration_over_threshold, g["DurationOverThreshold"]), (qi.count_wait_over_threshold, g["CountWaitOverThreshold"]),
                             (qi.count_dep_filled_merge_queue_tasks, g["CountDepFilledMergeQueueTasks"])):
            if mine != theirs:
                bad += 1
                print(f"{fx['Distro']}: queue info differs ({mine} vs {theirs})")
        if a.gpu:
            from evergreen_b200 import scheduler as S
            gp, _ = S.PrioritizeTasks(d, tasks_of(fx), now=now)
            if [t.id for t in gp] != mine_ids:
                bad += 1
                print(f"{fx['Distro']}: GPU order differs from the oracle's")
    print("distros:", len(fixtures), "mismatches:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
