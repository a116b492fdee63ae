import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
from workload_variant_autoscaler_b200 import Engine, Grid, synth_fleet
e = Engine(0); L = e._L
f = synth_fleet(1, 1, seed=42)
f.srv_slo_tps[:] = 0
# choose arrival so that rate/rmax(b=256) ~ 0.998 at r = 1
import oracle
qa = oracle.Analyzer(256, 2560, f.perf_alpha[0,0], f.perf_beta[0,0], f.perf_gamma[0,0], f.perf_delta[0,0], int(f.srv_in_tokens[0]), int(f.srv_out_tokens[0]))
rmin, rmax = qa.rate_range()
f.srv_arrival_rpm[:] = np.float32(rmax * 0.998 * 60)
g = Grid([256], [1])
L.wva_dbg_enable_cycles(e._h, 1)
for _ in range(3): cells, win = e.grid_solve(f, g, want_cells=True)
cyc = np.zeros(64, np.uint32); cl = np.zeros(64, np.uint32)
L.wva_dbg_read_cycles.restype = C.c_longlong
L.wva_dbg_read_cycles(e._h, cyc.ctypes.data_as(C.c_void_p), cl.ctypes.data_as(C.c_void_p), C.c_longlong(1))
print("flags", cells["flags"], "rho", cells["rho"], "cycles", cyc[:2], "per step (2816 x 2 passes):", cyc[0] / 5632.0, "kernel ms", e.last_kernel_ms)
