"""The drop-in boundary without a GPU: libevgsched.so loads, exports every symbol
include/evg_sched.h declares, its structs have the layout the numpy mirrors
assume, and the product never reaches into oracle/."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from evergreen_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "evg_sched.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(evg_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in evg_sched.h but not exported by libevgsched.so"
        assert n in L.SYMBOLS, f"{n} has no ctypes prototype in evergreen_b200/_lib.py"
    assert set(L.SYMBOLS) == set(names)
    assert lib.evg_abi_version() == 1


def test_struct_layout_matches_numpy_mirrors(tmp_path):
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "evg_sched.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(evg_task_soa), sizeof(evg_distro_cfg), sizeof(evg_distro_table),
         sizeof(evg_group_info), sizeof(evg_queue_info), sizeof(evg_host_soa), sizeof(evg_alloc_cfg),
         sizeof(evg_alloc_result), sizeof(evg_plan_out));
  printf("%zu %zu %zu %zu %zu\n", offsetof(evg_distro_cfg, num_dependents_factor), offsetof(evg_distro_cfg, n_versions),
         offsetof(evg_queue_info, ungrouped), offsetof(evg_alloc_cfg, provider), offsetof(evg_alloc_result, deficit_ns));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(evg_deps_in), offsetof(evg_deps_in, n_ext), sizeof(evg_runnable_in),
         offsetof(evg_runnable_in, task_off), offsetof(evg_runnable_in, deps), sizeof(evg_alloc_out));
  printf("%zu %zu %zu %zu\n", sizeof(evg_duration_rows), offsetof(evg_duration_rows, window_start_ns), sizeof(evg_duration_stat),
         offsetof(evg_duration_stat, stddev_ns));
  return 0;
}'''
    c = tmp_path / "t.c"
    c.write_text(prog)
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    sizes = [int(x) for x in out[0].split()]
    assert sizes == [ctypes.sizeof(L.TaskSoAStruct), L.DISTRO_CFG_DTYPE.itemsize, ctypes.sizeof(L.DistroTableStruct),
                     L.GROUP_INFO_DTYPE.itemsize, L.QUEUE_INFO_DTYPE.itemsize, ctypes.sizeof(L.HostSoAStruct),
                     L.ALLOC_CFG_DTYPE.itemsize, L.ALLOC_RESULT_DTYPE.itemsize, ctypes.sizeof(L.PlanOutStruct)]
    offs = [int(x) for x in out[1].split()]
    assert offs == [L.DISTRO_CFG_DTYPE.fields["num_dependents_factor"][1], L.DISTRO_CFG_DTYPE.fields["n_versions"][1],
                    L.QUEUE_INFO_DTYPE.fields["ungrouped"][1], L.ALLOC_CFG_DTYPE.fields["provider"][1],
                    L.ALLOC_RESULT_DTYPE.fields["deficit_ns"][1]]
    more = [int(x) for x in out[2].split()]
    assert more == [ctypes.sizeof(L.DepsInStruct), L.DepsInStruct.n_ext.offset, ctypes.sizeof(L.RunnableInStruct),
                    L.RunnableInStruct.task_off.offset, L.RunnableInStruct.deps.offset, ctypes.sizeof(L.AllocOutStruct)]
    dur = [int(x) for x in out[3].split()]
    assert dur == [ctypes.sizeof(L.DurationRowsStruct), L.DurationRowsStruct.window_start_ns.offset,
                   L.DURATION_STAT_DTYPE.itemsize, L.DURATION_STAT_DTYPE.fields["stddev_ns"][1]]


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = L.load()
    h = ctypes.c_void_p()
    rc = lib.evg_init(0, None, ctypes.byref(h))
    assert rc == L.EVG_ERR_CUDA and not h.value
    assert "no CPU fallback" in L.last_error()
    from evergreen_b200 import scheduler
    with pytest.raises(L.EvgError):
        scheduler.Engine(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "evergreen_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "evg_oracle" not in text, f
    code = "import sys; import evergreen_b200, evergreen_b200.scheduler, evergreen_b200.synth, evergreen_b200.dist; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
