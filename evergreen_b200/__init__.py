"""evergreen_b200: B200-native (sm_100a) implementation of Evergreen's scheduler
hot path -- scheduler.PlanDistro's tunable planner, DistroQueueInfo and the
utilization-based host allocator -- behind the reference's plug points.

The compute lives in libevgsched.so (evergreen_b200/csrc, C-ABI in
include/evg_sched.h).  Importing this package does not load CUDA; the first
call into `scheduler` does, and fails loudly without the library or a B200.
"""
from . import model  # noqa: F401

__all__ = ["model", "soa", "scheduler", "synth"]
