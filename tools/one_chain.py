"""Diagnostics: SM cycles of ONE chain (one warp, one lane) through the real grid kernel, by batch size and
queue ratio, to separate the head (table loads) from the tail (pure recurrence) cost per step."""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
import oracle
from workload_variant_autoscaler_b200 import Engine, Grid, synth_fleet
e = Engine(0); L = e._L
L.wva_dbg_read_cycles.restype = C.c_longlong
res = {}
for ratio in (10, 30):
    for b in (32, 128, 256):
        f = synth_fleet(1, 1, seed=42)
        f.srv_slo_tps[:] = 0
        f.max_queue_to_batch_ratio = ratio
        qa = oracle.Analyzer(b, b * ratio, f.perf_alpha[0,0], f.perf_beta[0,0], f.perf_gamma[0,0], f.perf_delta[0,0], int(f.srv_in_tokens[0]), int(f.srv_out_tokens[0]))
        rmin, rmax = qa.rate_range()
        f.srv_arrival_rpm[:] = np.float32(rmax * 0.9995 * 60)
        g = Grid([b], [1])
        L.wva_dbg_enable_cycles(e._h, 1)
        for _ in range(3): cells, win = e.grid_solve(f, g, want_cells=True)
        cyc = np.zeros(64, np.uint32); cl = np.zeros(64, np.uint32)
        L.wva_dbg_read_cycles(e._h, cyc.ctypes.data_as(C.c_void_p), cl.ctypes.data_as(C.c_void_p), C.c_longlong(1))
        K = b * (1 + ratio)
        res[(ratio, b)] = int(cyc[0])
        print("ratio", ratio, "b", b, "K", K, "flags", cells["flags"], "cycles", cyc[0], "per state-step (2 passes):", cyc[0] / (2.0 * K))
for b in (32, 128, 256):
    tail = (res[(30, b)] - res[(10, b)]) / (2.0 * 20 * b)
    head = (res[(10, b)] - tail * 2 * 10 * b) / (2.0 * b)
    print("b", b, "tail cycles/step %.1f  head cycles/step (incl. fixed) %.1f" % (tail, head))
