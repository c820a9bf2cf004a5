"""A/B of the one-shot call (host buffers in and out): python profiles/e2e_ab.py path/to/libevgsched.so [iters]"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from evergreen_b200 import _lib as L
path = sys.argv[1] if len(sys.argv) > 1 else L.LIB_PATH
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
lib = C.CDLL(path)
for name, (res, args) in list(L.SYMBOLS.items()):
    if hasattr(lib, name):
        fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
L._lib = lib
from evergreen_b200 import scheduler, synth
import bench
torch.cuda.set_device(0)
bound = bench.bind_to_gpu_node(torch, 0) is not None
eng = scheduler.Engine(0)
w = synth.config(2)
def pinned_like(a):
    v = a.view(np.int32) if a.dtype == np.uint32 else a
    return torch.from_numpy(v).pin_memory().numpy().view(a.dtype)
for name, _ in w.tasks.COLUMNS:
    setattr(w.tasks, name, pinned_like(getattr(w.tasks, name)))
for name, _ in w.hosts.COLUMNS:
    setattr(w.hosts, name, pinned_like(getattr(w.hosts, name)))
eng.plan_and_alloc_batch(w.tasks, w.distros, w.hosts, w.now)
ms = []
for _ in range(iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    po, ao = eng.plan_and_alloc_batch(w.tasks, w.distros, w.hosts, w.now)
    ms.append((time.perf_counter() - t0) * 1e3)
print(os.path.basename(path), "e2e ms median %.3f min %.3f max %.3f" % (float(np.median(ms)), min(ms), max(ms)), int(ao.result["new_hosts"].astype(np.int64).sum()), "bound" if bound else "unbound")
