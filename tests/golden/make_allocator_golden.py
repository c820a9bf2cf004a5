#!/usr/bin/env python3
"""Generate tests/golden/allocator_scenarios.json from the reference's own tests.

Reads /root/reference/scheduler/utilization_based_host_allocator_test.go
(only here, in the build container; the fixture it writes is what travels) and
replays every `func (s *UtilizationAllocatorSuite) Test...` body with a tiny
Go-literal -> Python transliteration, recording the inputs each test hands to
UtilizationBasedHostAllocator / calcNewHostsNeeded / calcExistingFreeHosts and
the values its assertions expect.

Clock: the reference calls time.Now() when building fixtures and time.Since()
later, so a strictly positive amount of time elapses in between.  The fixture
freezes `now` and places every `time.Now().Add(x)` at now + x - 1us, i.e. 1 us
of elapsed test time (some expectations, e.g. planner_test.go:250, depend on
elapsed > 0).
"""
import copy
import json
import os
import re
import sys

SRC = "/root/reference/scheduler/utilization_based_host_allocator_test.go"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "allocator_scenarios.json")

NOW = 1_800_000_000 * 10 ** 9
ELAPSED = 1000  # ns of test time between time.Now() and time.Since()
ZERO = -(2 ** 63)


class Obj:
    def __init__(self, typ, **kw):
        self.__dict__["_type"] = typ
        self.__dict__.update(kw)

    def __getattr__(self, k):  # unset Go fields read as zero values
        if k.startswith("__"):
            raise AttributeError(k)
        return DEFAULTS.get(self._type, {}).get(k, 0)

    def Insert(self, *a):
        DB[self._type].append(copy.deepcopy(self))
        return None

    def plain(self):
        out = {}
        for k, v in self.__dict__.items():
            if k == "_type":
                continue
            out[k] = plain(v)
        return out


def plain(v):
    if isinstance(v, Obj):
        return v.plain()
    if isinstance(v, list):
        return [plain(x) for x in v]
    if isinstance(v, Time):
        return v.ns
    return v


DEFAULTS = {
    "host.Host": {"Id": "", "RunningTask": "", "RunningTaskGroup": "", "RunningTaskProject": "",
                  "RunningTaskVersion": "", "RunningTaskBuildVariant": ""},
    "model.TaskGroupInfo": {"Name": ""},
    "distro.Distro": {"Id": "", "Provider": "", "Disabled": False},
    "distro.HostAllocatorSettings": {"RoundingRule": "", "FeedbackRule": "", "FutureHostFraction": 0.0},
}
DB = {"task.Task": [], "distro.Distro": []}


def mk(typ, *items, **kw):
    if typ.startswith("[]"):
        return list(items)
    return Obj(typ, **kw)


class Time:
    def __init__(self, ns):
        self.ns = ns

    def Add(self, d):
        return Time(self.ns + d)


class _TimePkg:
    Nanosecond, Microsecond, Millisecond = 1, 1000, 10 ** 6
    Second, Minute, Hour = 10 ** 9, 60 * 10 ** 9, 3600 * 10 ** 9

    @staticmethod
    def Now():
        return Time(NOW - ELAPSED)

    @staticmethod
    def Duration(x):
        return x


class _Evergreen:
    ProviderNameEc2Fleet = "ec2-fleet"
    ProviderNameStatic = "static"
    ProviderNameDocker = "docker"
    ProviderNameMock = "mock"
    ProviderNameEc2OnDemand = "ec2-ondemand"
    HostAllocatorRoundDown = "round-down"
    HostAllocatorRoundUp = "round-up"
    HostAllocatorRoundDefault = ""
    HostAllocatorNoFeedback = "no-feedback"
    HostAllocatorWaitsOverThreshFeedback = "waits-over-thresh-feedback"
    MaxDurationPerDistroHost = 30 * 60 * 10 ** 9
    MaxDurationPerDistroHostWithContainers = 2 * 60 * 10 ** 9


class _Fmt:
    @staticmethod
    def Sprintf(f, *a):
        return f % a


class Sentinel:
    def __init__(self, name):
        self.name = name

    def __radd__(self, o):
        return None

    def __add__(self, o):
        return None


class Suite:
    def __init__(self):
        self.distroName = "testDistro"
        self.projectName = "testProject"
        self.ctx = None
        self.calls = []
        self.vectors = []
        self.setup()

    def setup(self):  # SetupTest, reference lines 143-157
        self.distro = mk("distro.Distro", Id=self.distroName, Provider=_Evergreen.ProviderNameEc2Fleet,
                         HostAllocatorSettings=mk("distro.HostAllocatorSettings", MinimumHosts=0, MaximumHosts=50,
                                                  RoundingRule=_Evergreen.HostAllocatorRoundDown,
                                                  FeedbackRule=_Evergreen.HostAllocatorNoFeedback,
                                                  FutureHostFraction=.5))
        DB["task.Task"].clear()
        DB["distro.Distro"].clear()

    def NoError(self, *a):
        pass

    def T(self):
        return self

    def Context(self):
        return None

    def Equal(self, want, got, *a):
        if isinstance(got, Sentinel):
            self.calls[-1]["expect_" + got.name] = want
        elif isinstance(got, dict) and "fn" in got:
            got["expect"] = want
            self.vectors.append(got)


def transliterate(body: str) -> str:
    out = []
    for line in body.split("\n"):
        line = re.sub(r"//.*$", "", line).rstrip()
        if not line.strip():
            continue
        s = line.strip()
        if s.startswith(("ctx, cancel", "defer ")):
            continue
        s = s.replace(":=", "=").replace("&", "")
        s = re.sub(r"\btrue\b", "True", s)
        s = re.sub(r"\bfalse\b", "False", s)
        s = re.sub(r"\bdistro\s*=\s*distro\.Distro\{", "distro_ = distro.Distro{", s)
        s = re.sub(r"(\[\][\w\.]+|\b[A-Za-z_][\w]*\.[A-Z]\w*|\bHostAllocatorData)\{", lambda m: 'mk("%s",' % m.group(1), s)
        s = s.replace("}", ")")
        s = re.sub(r"^(\w+):\s*", r"\1=", s)
        s = s.replace("Distro=distro,", "Distro=distro_,")
        out.append(s)
    # join continuation lines: a statement ends when parentheses balance
    stmts, cur, depth = [], "", 0
    for s in out:
        cur += (" " if cur else "") + s
        depth += s.count("(") - s.count(")")
        if depth == 0:
            stmts.append(cur)
            cur = ""
    assert depth == 0, cur
    return "\n".join(stmts)


def run_test(name, body):
    s = Suite()
    env = {"s": s, "ctx": None, "mk": mk, "time": _TimePkg, "evergreen": _Evergreen, "fmt": _Fmt, "len": lambda x: len(x) if x is not None else 0}

    def allocator(ctx, data):
        snap = copy.deepcopy(data)
        s.calls.append({"data": snap, "tasks": copy.deepcopy(DB["task.Task"]), "distros": copy.deepcopy(DB["distro.Distro"])})
        return Sentinel("hosts"), Sentinel("free"), None

    def calc_new(*a):
        return {"fn": "calcNewHostsNeeded", "args": list(a)}

    def calc_free(ctx, hosts, frac, thr):
        rec = {"fn": "calcExistingFreeHosts", "hosts": plain(hosts), "fraction": frac, "threshold": thr,
               "tasks": plain(copy.deepcopy(DB["task.Task"]))}
        return rec, None

    env.update(UtilizationBasedHostAllocator=allocator, calcNewHostsNeeded=calc_new, calcExistingFreeHosts=calc_free)
    code = transliterate(body)
    exec(code, env)
    return s


def main():
    src = open(SRC).read()
    lines = src.split("\n")
    # locate suite methods
    starts = [(i, re.match(r"func \(s \*UtilizationAllocatorSuite\) (Test\w+)\(\)", l)) for i, l in enumerate(lines)]
    starts = [(i, m.group(1)) for i, m in starts if m]
    scenarios, vectors = [], []
    for k, (i, name) in enumerate(starts):
        j = i + 1
        while not lines[j].startswith("}"):
            j += 1
        body = "\n".join(lines[i + 1:j])
        s = run_test(name, body)
        for v in s.vectors:
            v["test"] = name
            v["ref"] = f"scheduler/utilization_based_host_allocator_test.go:{i + 1}"
            vectors.append(v)
        for c in s.calls:
            if "expect_hosts" not in c:
                continue
            data = c["data"]
            rec = {
                "test": name,
                "ref": f"scheduler/utilization_based_host_allocator_test.go:{i + 1}-{j + 1}",
                "now": NOW,
                "distro": plain(data.Distro),
                "hosts": plain(data.ExistingHosts or []),
                "queue_info": plain(data.DistroQueueInfo) if data.DistroQueueInfo else {},
                "container_pool": plain(data.ContainerPool) if data.ContainerPool else None,
                "running_tasks": plain(c["tasks"]),
                "db_distros": plain(c["distros"]),
                "expect_new_hosts": c["expect_hosts"],
                "expect_free_hosts": c["expect_free"],
            }
            scenarios.append(rec)
    json.dump({"generated_from": SRC, "now": NOW, "elapsed_ns": ELAPSED, "scenarios": scenarios, "vectors": vectors},
              open(OUT, "w"), indent=1, sort_keys=True)
    print(f"wrote {OUT}: {len(scenarios)} scenarios, {len(vectors)} vectors")


if __name__ == "__main__":
    sys.exit(main())
