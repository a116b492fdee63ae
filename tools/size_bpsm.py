"""Size path: blocks of sz2_solve per SM (WVA_SIZE_BLOCKS_PER_SM) on configs[3] (80k candidates, N <= 512, fresh solve)
and configs[4] (100k resident candidates, N <= 256, re-solve)."""
import os
import sys
import time
sys.path.insert(0, ".")
import numpy as np
from workload_variant_autoscaler_b200 import Engine, synth_fleet
e = Engine(0)
f3 = synth_fleet(10000, 8, seed=44, max_batch_choices=(4, 8, 16, 32, 64, 128, 256, 512))
f4 = synth_fleet(12500, 8, seed=45, max_batch_choices=(4, 8, 16, 32, 64, 128, 256))
for bpsm in ("4", "3", "2", "1"):
    os.environ["WVA_SIZE_BLOCKS_PER_SM"] = bpsm
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); e.solve(f3); ts.append((time.perf_counter() - t0) * 1e3)
    e.upload(f4); e.resolve()
    t4 = []
    for _ in range(5):
        f4.srv_arrival_rpm[:] = (f4.srv_arrival_rpm * 1.01).astype(np.float32)
        e.update_load(arrival_rpm=f4.srv_arrival_rpm)
        t0 = time.perf_counter(); e.resolve(); t4.append((time.perf_counter() - t0) * 1e3)
    print(f"blocks/SM {bpsm}: configs[3] solve {sorted(ts)[2]:.2f} ms   configs[4] resolve {sorted(t4)[2]:.2f} ms", flush=True)
