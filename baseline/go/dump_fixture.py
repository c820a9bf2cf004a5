#!/usr/bin/env python3
"""Write a synthetic tick as reference-shaped JSON lines for baseline/go/harness_test.go (see README.md)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from evergreen_b200 import _lib as L  # noqa: E402
from evergreen_b200 import synth  # noqa: E402

REQ = {0: "gitter_request", 1: "patch_request", 2: "github_merge_request"}
MIN = 60 * 10 ** 9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2, help="BASELINE config number 1..5 (synth.config)")
    ap.add_argument("--scale", type=float, default=0.01)
    ap.add_argument("--each", action="store_true")
    ap.add_argument("--snap", action="store_true", help="keep time-in-queue / wait 30 s away from minute boundaries")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    w = synth.config(a.config, a.scale, each=a.each)
    t, dt, h = w.tasks, w.distros, w.hosts
    if a.snap:
        for col in (t.queue_basis_ns, t.wait_basis_ns):
            col[:] = w.now - ((w.now - col) // MIN) * MIN - 30 * 10 ** 9
        t.expected_ns[:] = (t.expected_ns // MIN) * MIN + 30 * 10 ** 9
    with open(a.out, "w") as f:
        for d in range(dt.n_distros):
            a0, a1 = int(dt.task_off[d]), int(dt.task_off[d + 1])
            cfg = dt.cfg[d]
            tasks = []
            for i in range(a0, a1):
                fl = int(t.flags[i])
                deps = []
                if t.dep_idx is not None:
                    deps = [f"d{d}t{int(x)}" for x in t.dep_idx[int(t.dep_off[i]):int(t.dep_off[i + 1])]]
                met = bool(fl & L.EVG_TF_DEPS_MET)
                if not met and not deps:
                    deps = [f"d{d}t{(i - a0 + 1) % max(a1 - a0, 1)}"]  # an unmet dependency must be in the queue (no Mongo)
                g = int(t.group_id[i])
                tasks.append({
                    "Id": f"d{d}t{i - a0}", "Version": f"d{d}v{int(t.version_id[i])}", "Project": "p", "BuildVariant": "bv",
                    "TaskGroup": "" if g < 0 else f"tg{g}", "TaskGroupOrder": int(t.task_group_order[i]),
                    "TaskGroupMaxHosts": 0 if g < 0 else int(dt.group_max_hosts[int(dt.group_off[d]) + g]),
                    "Priority": int(t.priority[i]), "Requester": REQ[fl & 3], "GenerateTask": bool(fl & L.EVG_TF_GENERATE),
                    "ActivatedBy": "stepback" if fl & L.EVG_TF_STEPBACK else "", "NumDependents": int(t.num_dependents[i]),
                    "DistroId": "elsewhere" if fl & L.EVG_TF_OTHER_DISTRO else f"d{d}",
                    "ActivatedAgoNs": int(w.now - t.queue_basis_ns[i]), "ScheduledAgoNs": int(w.now - t.wait_basis_ns[i]),
                    "ExpectedNs": int(t.expected_ns[i]), "DependsOn": deps, "OverrideDependencies": met and bool(deps)})
            hosts = []
            if h is not None:
                for k in range(int(h.host_off[d]), int(h.host_off[d + 1])):
                    hosts.append({"Id": f"d{d}h{k}", "Free": not (int(h.flags[k]) & L.EVG_HF_RUNNING)})
            row = {"Distro": f"d{d}", "NowNs": int(w.now),
                   "Planner": {k: (float(cfg[k]) if k == "num_dependents_factor" else int(cfg[k])) for k in cfg.dtype.names if k != "_reserved"},
                   "Tasks": tasks, "Hosts": hosts}
            if h is not None:
                ac = h.cfg[d]
                row["Allocator"] = {k: (float(ac[k]) if k == "future_host_fraction" else int(ac[k])) for k in ac.dtype.names}
            f.write(json.dumps(row) + "\n")
    print(f"{a.out}: {dt.n_distros} distros, {w.n_tasks} tasks")


if __name__ == "__main__":
    main()
