import sys
sys.path.insert(0, ".")
import numpy as np
from workload_variant_autoscaler_b200 import Engine, synth_fleet
e = Engine(0)
f = synth_fleet(12500, 8, seed=45, max_batch_choices=(4, 8, 16, 32, 64, 128, 256))
e.upload(f)
e.resolve()
f.srv_arrival_rpm[:] = (f.srv_arrival_rpm * 1.01).astype(np.float32)
e.update_load(arrival_rpm=f.srv_arrival_rpm)
e.resolve()
print(e.last_kernel_ms, e.launch_count)
