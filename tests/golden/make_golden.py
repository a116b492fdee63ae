"""Generate the committed golden fixtures (run from the repo root: python tests/golden/make_golden.py).

The reference is Go and cannot be run in this image, so these vectors are NOT outputs of the
Go binary: they are outputs of the C oracle (oracle/wva_oracle.c) on inputs built from the
reference's own fixtures, frozen so that oracle or kernel regressions are caught, and
cross-checked at generation time against the independent numpy restatement.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import restate_np as R  # noqa: E402
from workload_variant_autoscaler_b200 import Grid, synth_fleet  # noqa: E402
from workload_variant_autoscaler_b200.fleet import CONFIG1_LOADS, CONFIG1_TOKENS, config1_fleet  # noqa: E402,F401

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    # --- config 1 -------------------------------------------------------------------------
    rows = []
    for it, ot in CONFIG1_TOKENS:
        for rpm in CONFIG1_LOADS:
            f = config1_fleet(rpm, it, ot)
            cand, win = oracle.solve(f)
            r = R.create_allocation(f, 0, 0)
            w = win[0]
            assert int(w["feasible"]) == r["feasible"]
            if r["feasible"]:
                assert int(w["replicas"]) == r["replicas"]
                for k in ("cost", "itl", "ttft", "rho", "max_rate"):
                    assert np.float32(w[k]).view(np.uint32) == np.float32(r[k]).view(np.uint32), (rpm, k)
            rows.append(w)
    out["config1_winners"] = np.array(rows, dtype=oracle.ALLOC_DTYPE)
    # --- small synthetic fleet: size candidates, winners, grid, sweep -------------------------
    f = synth_fleet(12, 3, seed=2024, max_batch_choices=(2, 4, 8, 16, 32), zero_load_frac=0.15)
    f.srv_min_replicas[::5] = 0
    cand, win = oracle.solve(f)
    for s in range(f.n_servers):  # cross-check a subset against the numpy restatement
        for a in range(0, f.n_acc, 2):
            r = R.create_allocation(f, s, a)
            o = oracle.create_allocation(f, s, a)
            assert o["feasible"] == r["feasible"]
            if r["feasible"]:
                assert o["replicas"] == r["replicas"]
                assert np.float32(o["ttft"]).view(np.uint32) == np.float32(r["ttft"]).view(np.uint32)
    out["synth_cand"], out["synth_win"] = cand, win
    grid = Grid([1, 2, 4, 8, 16, 31], [1, 2, 3, 4, 6, 8, 12, 16, 24, 33])
    cells, gwin = oracle.grid_solve(f, grid)
    out["grid_cells"], out["grid_win"] = cells, gwin
    sw = oracle.sweep(f, 16)
    for k, v in sw.items():
        out["sweep_" + k] = v
    np.savez_compressed(os.path.join(HERE, "golden_r01.npz"), **out)
    print("wrote golden_r01.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
