import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from evergreen_b200 import scheduler, synth
eng = scheduler.Engine(0)
w = synth.make(np.full(296, 10000), synth.SEED_BASE + 2, n_hosts=1500)
eng.upload(w.tasks, w.distros, w.hosts)
for _ in range(3):
    eng.run(w.now)
po, ao = eng.download()
print("ok", eng.last_timing_ms(), int(ao.result["new_hosts"].sum()))
