#!/usr/bin/env python3
"""Write tests/golden/expected_duration.json: the known answer of model/task/expected_duration_test.go:14-69
(TestExpectedDuration), transcribed by hand (the Go test inserts into MongoDB and cannot run here)."""
import json
import os

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "expected_duration.json")
MIN = 60 * 10 ** 9
NOW = 1_800_000_000 * 10 ** 9
tasks = [dict(id=f"t{i + 1}", build_variant="bv", project="proj", status="success", finish_time=NOW,
              start_time=NOW - m * MIN, time_taken=m * MIN) for i, m in enumerate([10, 30, 35, 25])]
case = dict(name="TestExpectedDuration", ref="model/task/expected_duration_test.go:14-69", tasks=tasks,
            window_start=NOW - 60 * MIN, window_end=NOW,
            # assert.EqualValues(25*time.Minute, results[0].ExpectedDuration); assert.InDelta(9.35*min, StdDev, 0.01*min)
            expect=dict(key=["proj", "bv", ""], mean_ns=25 * MIN, mean_exact=True, stddev_ns=9.35 * MIN, stddev_delta_ns=0.01 * MIN))
json.dump(dict(source="model/task/expected_duration_test.go", cases=[case]), open(OUT, "w"), indent=1)
print("wrote", OUT)
