"""Diagnostics: per-request SM cycles of one sz2_solve round on the streaming workload (run on the GPU box)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from workload_variant_autoscaler_b200 import Engine, synth_fleet  # noqa: E402

e = Engine(0)
L = e._L
if os.environ.get("SIZE_DBG_FLEET") == "config3":  # 80 k candidates, N <= 512
    f = synth_fleet(10000, 8, seed=44, max_batch_choices=(4, 8, 16, 32, 64, 128, 256, 512))
else:
    f = synth_fleet(12500, 8, seed=45, max_batch_choices=(4, 8, 16, 32, 64, 128, 256))
e.upload(f)
e.resolve()
L.wva_dbg_read_size.restype = C.c_longlong
for rnd in [int(x) for x in (sys.argv[1:] or ["0", "5", "14"])]:
    os.environ["WVA_SIZE_DBG_ROUND"] = str(rnd)
    e.resolve()
    cap = 2_000_000
    buf = np.zeros((cap, 4), np.uint32)
    got = L.wva_dbg_read_size(e._h, buf.ctypes.data_as(C.c_void_p), C.c_longlong(cap))
    b = buf[:got]
    b = b[b[:, 0] > 0]
    cyc, N, kind, lam, w = b[:, 0], b[:, 1] & 0xffff, b[:, 1] >> 16, b[:, 2].view(np.float32), b[:, 3]
    print(f"== round {rnd}: requests {b.shape[0]}  kinds {np.bincount(kind, minlength=4)}  N hist "
          f"{dict(zip(*np.unique(N, return_counts=True)))}")
    order = np.argsort(w, kind="stable")
    ws, first = np.unique(w[order], return_index=True)
    wmax = np.maximum.reduceat(cyc[order], first)
    print("warps", ws.size, "warp cycles: sum %.3e mean %.0f p50 %.0f p99 %.0f max %d" %
          (wmax.sum(dtype=np.float64), wmax.mean(), np.percentile(wmax, 50), np.percentile(wmax, 99), wmax.max()))
    print("sum of warp cycles / (148 SMs x 32 warps): %.1f us;  longest warp: %.1f us;  warps above 80 %% of it: %d" %
          (wmax.sum(dtype=np.float64) / (148 * 32) / 1.965e3, wmax.max() / 1.965e3, int((wmax > 0.8 * wmax.max()).sum())))
    for k in np.argsort(wmax)[::-1][:4]:
        sel = order[first[k]: first[k] + 32]
        sel = sel[w[sel] == ws[k]]
        print("  warp", ws[k], "cycles", wmax[k], "lanes", sel.size, "N", np.unique(N[sel]), "kinds", np.bincount(kind[sel], minlength=4),
              "lane cycles min/med/max", cyc[sel].min(), int(np.median(cyc[sel])), cyc[sel].max(),
              "lambda min/max %.5g %.5g" % (lam[sel].min(), lam[sel].max()))
