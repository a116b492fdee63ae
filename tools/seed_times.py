"""Device time of one config-2 grid solve per synthetic fleet seed (the N-GPU bench gives rank r seed 42 + r)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from workload_variant_autoscaler_b200 import Engine, config2_grid, synth_fleet
e = Engine(0)
g = config2_grid()
for seed in range(42, 50):
    f = synth_fleet(100, 4, seed=seed)
    d, k = [], []
    for _ in range(8):
        e.grid_solve(f, g); d.append(e.last_device_ms); k.append(e.last_kernel_ms)
    import ctypes as C, struct
    plan = (C.c_uint32 * 31)()
    e._L.wva_dbg_read_plan(e._h, plan)
    fl = lambda u: struct.unpack("f", struct.pack("I", u))[0]
    print("   plan: n_long", plan[0], "L", plan[3], " model t_short/t_long per L:", [(round(fl(plan[7 + 2 * l]), 1), round(fl(plan[8 + 2 * l]), 1)) for l in range(4)])
    print("seed", seed, "device ms %.3f" % np.median(d[2:]), "grid_kernel ms %.3f" % np.median(k[2:]))
