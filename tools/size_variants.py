"""Size-path variants on the streaming workload (100k candidates): WVA_SIZE_REV x WVA_SIZE_DEPTH, median resolve time."""
import os
import sys
import time
sys.path.insert(0, ".")
import numpy as np
from workload_variant_autoscaler_b200 import Engine, synth_fleet
e = Engine(0)
f = synth_fleet(12500, 8, seed=45, max_batch_choices=(4, 8, 16, 32, 64, 128, 256))
e.upload(f)
e.resolve()
ref = None
for rev in ("0", "1"):
    for depth in ("2", "3", "4"):
        os.environ["WVA_SIZE_REV"] = rev
        os.environ["WVA_SIZE_DEPTH"] = depth
        ts = []
        for k in range(5):
            f.srv_arrival_rpm[:] = (f.srv_arrival_rpm * 1.01).astype(np.float32)
            e.update_load(arrival_rpm=f.srv_arrival_rpm)
            l0 = e.launch_count
            t0 = time.perf_counter()
            w = e.resolve()
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"rev {rev} depth {depth}: median {sorted(ts)[2]:.2f} ms  min {min(ts):.2f}  launches {e.launch_count - l0}", flush=True)
