"""Where does the end-to-end time of wva_grid_solve go (BASELINE config 2)?"""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
from workload_variant_autoscaler_b200 import Engine, config2_grid, synth_fleet, _abi
from workload_variant_autoscaler_b200._abi import Allocs
e = Engine(0); L = e._L
f = synth_fleet(100, 4, seed=42); g = config2_grid()
for _ in range(5): e.grid_solve(f, g)
n = 200
t0 = time.perf_counter()
for _ in range(n): e.grid_solve(f, g)
t_wrap = (time.perf_counter() - t0) / n
win = Allocs(f.n_servers)
fc, gc, wc = f.as_c(), g.as_c(), win.as_c()
t0 = time.perf_counter()
for _ in range(n): L.wva_grid_solve(e._h, C.byref(fc), C.byref(gc), None, C.byref(wc))
t_c = (time.perf_counter() - t0) / n
dev = []; ker = []
for _ in range(50):
    L.wva_grid_solve(e._h, C.byref(fc), C.byref(gc), None, C.byref(wc)); dev.append(e.last_device_ms); ker.append(e.last_kernel_ms)
print("python wrapper call  %.1f us" % (t_wrap * 1e6))
print("bare C ABI call      %.1f us" % (t_c * 1e6))
print("device span (events) %.1f us   grid_kernel %.1f us" % (np.median(dev) * 1e3, np.median(ker) * 1e3))
