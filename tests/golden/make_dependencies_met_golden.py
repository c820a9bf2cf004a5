#!/usr/bin/env python3
"""Write tests/golden/dependencies_met.json: the assertions of TestDependenciesMet (model/task/task_test.go:249-434,
fixture :37-55) and TestBlocked (:1195-1223), transcribed by hand (the Go test reads and writes MongoDB).

Each case: the task's DependsOn, what the tasks collection holds for the dependency tasks at that point of the test,
optional overrides, and the expected Task.DependenciesMet / Task.AllDependenciesSatisfied where the test asserts them."""
import json
import os

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dependencies_met.json")
F = "model/task/task_test.go"
UND, OK, FAIL = "undispatched", "success", "failed"
DEP_IDS = [dict(task_id="td1", status=OK), dict(task_id="td2", status=OK), dict(task_id="td3", status=""),
           dict(task_id="td4", status=FAIL), dict(task_id="td5", status="*")]  # :37-43


def db(**status):
    st = dict(td1=UND, td2=UND, td3=UND, td4=UND, td5=UND)  # :262-268
    st.update(status)
    return [dict(id=k, status=v) for k, v in st.items()]


UPDATED = dict(td1=OK, td2=OK, td3=OK, td4=FAIL, td5=FAIL)  # updateTestDepTasks :46-55
cases = [
    dict(name="nil case", ref=f"{F}:276-282", depends_on=[], db=db(), all_satisfied=True),
    dict(name="no dependencies", ref=f"{F}:284-292", depends_on=[], db=db(), met=True),
    dict(name="overridden dependencies", ref=f"{F}:294-302", depends_on=DEP_IDS, db=db(), override_dependencies=True, met=True),
    dict(name="only some finished", ref=f"{F}:304-323", depends_on=DEP_IDS, db=db(td1=OK), met=False),
    dict(name="all finished properly", ref=f"{F}:325-333", depends_on=DEP_IDS, db=db(**UPDATED), met=True),
    dict(name="pulled into the cache", ref=f"{F}:335-349", depends_on=DEP_IDS, db=db(**UPDATED), met=True),
    dict(name="cached copy altered to failed", ref=f"{F}:351-373", depends_on=DEP_IDS, db=db(**dict(UPDATED, td1=FAIL)), met=False),
    dict(name="extraneous/three dependencies, one failed", ref=f"{F}:375-417", depends_on=DEP_IDS[:3],
         db=db(td1=OK, td2=OK, td3=FAIL), met=False, all_satisfied=False),
    dict(name="extraneous/failed one removed from DependsOn", ref=f"{F}:419-431", depends_on=DEP_IDS[:2],
         db=db(td1=OK, td2=OK, td3=FAIL), met=True, all_satisfied=True),
]
blocked = [
    dict(name="Blocked", ref=f"{F}:1197-1207", unattainable=[False, False, True], expect=True),
    dict(name="NotBlocked", ref=f"{F}:1208-1218", unattainable=[False, False, False], expect=False),
]
json.dump(dict(source=F, cases=cases, blocked=blocked), open(OUT, "w"), indent=1)
print("wrote", OUT, len(cases), "+", len(blocked), "cases")
