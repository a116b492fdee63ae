"""Diagnostics (needs a -DWVA_PROF build in tools/var/libwva_prof.so): phase clocks and counters of one chain."""
import ctypes as C, sys, shutil
import numpy as np
sys.path.insert(0, ".")
shutil.copy("tools/var/libwva_prof.so", "workload_variant_autoscaler_b200/libwva_b200.so")
import oracle
from workload_variant_autoscaler_b200 import Engine, Grid, synth_fleet
e = Engine(0); L = e._L
for ratio, b in [(int(a.split(":")[0]), int(a.split(":")[1])) for a in (sys.argv[1:] or ["10:256", "10:64", "1:256"])]:
    f = synth_fleet(1, 1, seed=42)
    f.srv_slo_tps[:] = 0
    f.max_queue_to_batch_ratio = ratio
    qa = oracle.Analyzer(b, b * ratio, f.perf_alpha[0,0], f.perf_beta[0,0], f.perf_gamma[0,0], f.perf_delta[0,0], int(f.srv_in_tokens[0]), int(f.srv_out_tokens[0]))
    rmin, rmax = qa.rate_range()
    f.srv_arrival_rpm[:] = np.float32(rmax * 0.9995 * 60)
    g = Grid([b], [1])
    for _ in range(2): e.grid_solve(f, g)
    out = (C.c_longlong * 16)()
    L.wva_dbg_prof(out, 1)
    e.grid_solve(f, g)
    L.wva_dbg_prof(out, 0)
    o = list(out)
    print("ratio", ratio, "b", b, "K", b * (1 + ratio))
    print("  pass1 total", o[1] - o[0], " staged head", o[12] - o[0], "(chunks %d, left at n=%d)" % (o[4], o[5]), " rest", o[1] - o[12])
    print("  item: fetch+decode", o[10] - o[9], " solve entry->pass1", o[0] - o[10], " pass2 end->solve exit", o[11] - o[3], " epilogue", o[14] - o[11], " total", o[14] - o[9])
    print("  between passes + stash", o[2] - o[1], " pass2 staged head", o[13] - o[2], "(chunks %d, left at i=%d)" % (o[7], o[8]), " rest", o[3] - o[13])
