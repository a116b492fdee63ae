"""Target for ncu: configs[3] (10,000 servers x 8 accelerators, N <= 512), two fresh solves."""
import sys
import time
sys.path.insert(0, ".")
from workload_variant_autoscaler_b200 import Engine, synth_fleet
e = Engine(0)
f = synth_fleet(10000, 8, seed=44, max_batch_choices=(4, 8, 16, 32, 64, 128, 256, 512))
for _ in range(2):
    t0 = time.perf_counter()
    l0 = e.launch_count
    e.solve(f)
    print("solve ms", (time.perf_counter() - t0) * 1e3, "device ms", e.last_device_ms, "launches", e.launch_count - l0)
