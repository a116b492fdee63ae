"""Diagnostics: per-lane SM cycles of grid_kernel on BASELINE config 2 (run on the GPU box)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from workload_variant_autoscaler_b200 import Engine, config2_grid, synth_fleet  # noqa: E402

e = Engine(0)
L = e._L
f = synth_fleet(100, 4, seed=42)
g = config2_grid()
L.wva_dbg_enable_cycles(e._h, 1)
e.grid_solve(f, g)
e.grid_solve(f, g)
n = f.n_servers * f.n_acc * 256 * 64
cyc = np.zeros(2 * n, np.uint32)
cells = np.zeros(2 * n, np.uint32)
L.wva_dbg_read_cycles.restype = C.c_longlong
got = L.wva_dbg_read_cycles(e._h, cyc.ctypes.data_as(C.c_void_p), cells.ctypes.data_as(C.c_void_p), C.c_longlong(n))
print("kernel ms", e.last_kernel_ms)
cyc = cyc[:got]
w = cyc[: got // 32 * 32].reshape(-1, 32).max(axis=1)
w = w[w > 0]
print("warps with work", w.size, "sum warp cycles %.3e" % w.sum(dtype=np.float64), "mean", w.mean(), "max", w.max())
for q in (10, 50, 90, 99, 99.9):
    print("pct", q, np.percentile(w, q))
print("first 20 warps (launch order):", w[:20])
order = np.sort(w)[::-1]
cs = np.cumsum(order, dtype=np.float64)
for frac in (0.01, 0.05, 0.1, 0.25, 0.5):
    k = int(frac * w.size)
    print("top %4.0f%% of warps hold %5.1f%% of warp-cycles" % (100 * frac, 100 * cs[k] / cs[-1]))
print("ideal ms if perfectly packed on 148 SMs x 32 warps:", w.sum(dtype=np.float64) / (148 * 32) / 1.965e9 * 1e3)
