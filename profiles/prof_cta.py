"""configs[1] (1000 distros x 10k tasks) through the resident tick, for ncu: `python profiles/prof_cta.py [ticks]`."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evergreen_b200 import scheduler, synth
eng = scheduler.Engine(0)
w = synth.config(2)
eng.upload(w.tasks, w.distros, w.hosts)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    eng.run(w.now)
po, ao = eng.download()
print("ok", eng.last_timing_ms(), eng.kernel_timing_ms(1), int(ao.result["new_hosts"].sum()))
