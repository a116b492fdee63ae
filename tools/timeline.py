"""Diagnostics: per-warp timeline of grid_kernel on BASELINE config 2 (start/end ns, SM) -> per-SM busy profile."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from workload_variant_autoscaler_b200 import Engine, config2_grid, synth_fleet  # noqa: E402

e = Engine(0)
L = e._L
SEED = int(sys.argv[1]) if len(sys.argv) > 1 else 42
f = synth_fleet(100, 4, seed=SEED)
g = config2_grid()
L.wva_dbg_enable_cycles(e._h, 1)
for _ in range(6):
    e.grid_solve(f, g)
n = f.n_servers * f.n_acc * 256 * 64
cyc = np.zeros(n, np.uint32)
tl = np.zeros(n, np.uint32)
L.wva_dbg_read_cycles.restype = C.c_longlong
got = L.wva_dbg_read_cycles(e._h, cyc.ctypes.data_as(C.c_void_p), tl.ctypes.data_as(C.c_void_p), C.c_longlong(n))
print("seed", SEED, "kernel ms", e.last_kernel_ms)
nw = got // 32
cw = cyc[: nw * 32].reshape(nw, 32).max(axis=1)
t = tl[: nw * 32].reshape(nw, 32)
ok = cw > 0
start, sm, end = t[ok, 0].astype(np.int64), t[ok, 1], t[ok, 2].astype(np.int64)
cw = cw[ok]
t0 = start.min()
start -= t0
end -= t0
print("warps", cw.size, "span us", end.max() / 1e3, "clock GHz est", np.median(cw / np.maximum(end - start, 1)))
dur = end - start
# per-SM finish time and busy warp-time
fin = np.zeros(148)
busy = np.zeros(148)
for k in range(148):
    m = sm == k
    if m.any():
        fin[k] = end[m].max()
        busy[k] = dur[m].sum()
print("per-SM finish us: min %.1f median %.1f max %.1f" % (fin.min() / 1e3, np.median(fin) / 1e3, fin.max() / 1e3))
print("per-SM warp-busy us (sum of durations): min %.0f median %.0f max %.0f" % (busy.min() / 1e3, np.median(busy) / 1e3, busy.max() / 1e3))
# resident warps over time (whole GPU)
edges = np.linspace(0, end.max(), 21)
for a, b in zip(edges[:-1], edges[1:]):
    mid = 0.5 * (a + b)
    res = ((start <= mid) & (end > mid)).sum()
    print("t=%6.1f us resident warps %6d (%.1f / SM)" % (mid / 1e3, res, res / 148))
# launch-order view: duration of warps by launch index
idx = np.nonzero(ok)[0]
for lo in (0, 50, 100, 150, 200, 300, 400, 500, 600, 800, 1000, 1500, 2000, 3000, 3552, 5000, 8000, 12000, 16000, 20000):
    sel = (idx >= lo) & (idx < lo + 50)
    if sel.any():
        print("warps %6d..: start %.1f us dur %.1f us (max %.1f) cycles/lane mean %.0f" % (lo, start[sel].mean() / 1e3, dur[sel].mean() / 1e3, dur[sel].max() / 1e3, cw[sel].mean()))

for thr in (400e3, 300e3, 200e3, 150e3, 100e3, 50e3):
    print("items with lane cycles > %dk: %d" % (thr / 1e3, int((cw > thr).sum())))
print("sum of item durations (warp-us): %.0f   / (148 SMs x 13 warps) = %.1f us" % (dur.sum() / 1e3, dur.sum() / 1e3 / (148 * 13)))
