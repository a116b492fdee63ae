"""Experiment: one full wave (and fractions of it) of identical longest chains (b = 256, u = 0.998)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import oracle
from workload_variant_autoscaler_b200 import Engine, Grid, synth_fleet
e = Engine(0)
f = synth_fleet(1, 1, seed=42); f.srv_slo_tps[:] = 0
qa = oracle.Analyzer(256, 2560, f.perf_alpha[0,0], f.perf_beta[0,0], f.perf_gamma[0,0], f.perf_delta[0,0], int(f.srv_in_tokens[0]), int(f.srv_out_tokens[0]))
rmin, rmax = qa.rate_range()
f.srv_arrival_rpm[:] = np.float32(rmax * 0.998 * 60)
for warps_per_sm in (1, 2, 4, 8, 12, 16, 24):
    n = 148 * warps_per_sm * 32
    g = Grid([256], np.ones(n, np.int32))
    ks = []
    for _ in range(4):
        e.grid_solve(f, g); ks.append(e.last_kernel_ms)
    print("warps/SM", warps_per_sm, "cells", n, "grid_kernel ms", min(ks[1:]), "-> cycles/step", min(ks[1:]) * 1e-3 * 1.965e9 / 5632)
