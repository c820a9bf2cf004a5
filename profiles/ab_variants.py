#!/usr/bin/env python3
"""A/B several builds of libevgsched.so on one box: resident configs[1] tick and its dominant kernel.
  python profiles/ab_variants.py build NAME -DEVG_CTA_TPT=1 ...   (here: nvcc -> evergreen_b200/variants/NAME.so)
  python profiles/ab_variants.py run [ticks]                      (on the GPU box: every variant + the in-tree build)
Each variant runs in its own process; medians over `ticks` resident ticks after 20 warm-up ticks."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, "evergreen_b200", "variants")


def build(name, flags):
    os.makedirs(VAR, exist_ok=True)
    out = os.path.join(VAR, name + ".so")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           *flags, "-o", out, os.path.join(ROOT, "evergreen_b200", "csrc", "evg_sched.cu")]
    subprocess.check_call(cmd)
    print(out)


CHILD = r'''
import sys, os, ctypes as C
sys.path.insert(0, %r)
import numpy as np
from evergreen_b200 import _lib as L
lib = C.CDLL(sys.argv[1])
for name, (res, args) in L.SYMBOLS.items():
    if hasattr(lib, name):
        fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
L._lib = lib
from evergreen_b200 import scheduler, synth
eng = scheduler.Engine(0)
w = synth.config(2)
eng.upload(w.tasks, w.distros, w.hosts)
n = int(sys.argv[2])
for _ in range(20): eng.run(w.now)
ms = []
for _ in range(n):
    eng.run(w.now); ms.append(eng.last_timing_ms()[0])
k = eng.kernel_timing_ms(min(n, 100))
po, ao = eng.download()
print("%%-28s tick median %%.4f ms  min %%.4f | kernel median %%.4f ms | checksum %%d" %% (os.path.basename(sys.argv[1]), float(np.median(ms)), float(min(ms)), float(np.median(k)), int(ao.result["new_hosts"].sum()) + int(po.order[::997].sum())))
'''


def run(ticks):
    libs = sorted(glob.glob(os.path.join(VAR, "*.so"))) + [os.path.join(ROOT, "evergreen_b200", "libevgsched.so")]
    for lib in libs:
        r = subprocess.run([sys.executable, "-c", CHILD % ROOT, lib, str(ticks)], capture_output=True, text=True, timeout=600)
        print((r.stdout.strip() or r.stderr.strip()[-400:]))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2], sys.argv[3:])
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 200)
