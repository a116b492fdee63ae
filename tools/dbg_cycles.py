"""Diagnostics: per-cell SM cycles of grid_kernel on BASELINE config 2 (run on the GPU box)."""
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
cyc = np.zeros(n, np.uint32)
cells = np.zeros(n, np.uint32)
L.wva_dbg_read_cycles.restype = C.c_longlong
got = L.wva_dbg_read_cycles(e._h, cyc.ctypes.data_as(C.c_void_p), cells.ctypes.data_as(C.c_void_p), C.c_longlong(n))
print("cells", got, "kernel ms", e.last_kernel_ms)
cyc = cyc[:got]
print("sum cycles %.3e  mean %.1f  max %d" % (cyc.sum(dtype=np.float64), cyc.mean(), cyc.max()))
w = cyc.reshape(-1, 32) if got % 32 == 0 else cyc[: got // 32 * 32].reshape(-1, 32)
wmax = w.max(axis=1)
print("warp-max sum %.3e (x32 = %.3e lane-cycles)" % (wmax.sum(dtype=np.float64), 32.0 * wmax.sum(dtype=np.float64)))
for q in (50, 90, 99, 99.9, 100):
    print("pct", q, np.percentile(cyc, q))
top = np.argsort(-cyc.astype(np.int64))[:10]
for i in top:
    c = int(cells[i]); ri = c % 64; bi = (c // 64) % 256; a = (c // (64 * 256)) % 4; s = c // (64 * 256 * 4)
    print("idx", i, "cycles", cyc[i], "cell s,a,b,r", s, a, bi + 1, ri + 1)
blk = cyc[: got // 256 * 256].reshape(-1, 256).max(axis=1)
print("block max: first 10", blk[:10], " mean", blk.mean(), "max", blk.max(), "argmax", blk.argmax(), "of", blk.size)
print("blocks with max > 1e6 cycles:", (blk > 1e6).sum(), " > 1e5:", (blk > 1e5).sum())
