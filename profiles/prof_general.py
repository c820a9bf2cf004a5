"""General-path shapes for the ncu launch list: `python profiles/prof_general.py 1m|100k`."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from evergreen_b200 import scheduler, synth
which = sys.argv[1] if len(sys.argv) > 1 else "1m"
sizes = np.full(8, 1000000) if which == "1m" else np.full(48, 100000)
eng = scheduler.Engine(0)
w = synth.make(sizes, synth.SEED_BASE + 3, n_hosts=5 * len(sizes))
eng.upload(w.tasks, w.distros, w.hosts)
for _ in range(3):
    eng.run(w.now)
po, ao = eng.download()
print("ok", which, eng.last_timing_ms(), int(ao.result["new_hosts"].sum()))
