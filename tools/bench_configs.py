"""Secondary configurations of BASELINE.json (3: latency sweep, 4: 10k-server min-cost solve,
5: streaming re-solve). Prints one JSON line per configuration. Run on the GPU box."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from workload_variant_autoscaler_b200 import Engine, synth_fleet  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), float(np.min(ts))


def main():
    which = sys.argv[1:] or ["3", "4", "5"]
    e = Engine(0)
    if "3" in which:  # 1000 models x 8 accelerator types, 256 rates each
        f = synth_fleet(1000, 8, seed=43, max_batch_choices=(4, 8, 16, 32, 64, 128, 256, 512))
        med, best = timed(lambda: e.sweep(f, 256), reps=3, warm=1)
        n = f.n_servers * f.n_acc * 256
        print(json.dumps({"config": 3, "solves": n, "e2e_ms": med * 1e3, "kernel_ms": e.last_kernel_ms,
                          "solves_per_s_e2e": n / med, "solves_per_s_kernel": n / (e.last_kernel_ms * 1e-3)}))
    if "4" in which:  # 10,000 servers x 8 accelerators, unlimited min-cost assignment
        f = synth_fleet(10000, 8, seed=44, max_batch_choices=(4, 8, 16, 32, 64, 128, 256, 512))
        med, best = timed(lambda: e.solve(f, want_candidates=False), reps=3, warm=1)
        n = f.n_servers * f.n_acc
        print(json.dumps({"config": 4, "size_candidates": n, "e2e_ms": med * 1e3, "kernel_ms": e.last_kernel_ms,
                          "device_ms": e.last_device_ms, "candidates_per_s_e2e": n / med}))
    if "5" in which:  # 100k candidates resident, arrival churn each tick
        f = synth_fleet(12500, 8, seed=45, max_batch_choices=(4, 8, 16, 32, 64, 128, 256))
        e.upload(f)
        rng = np.random.default_rng(5)
        lat = []
        e.resolve()
        for _ in range(30):
            f.srv_arrival_rpm[:] = (f.srv_arrival_rpm * np.exp(rng.normal(0, 0.1, f.n_servers))).astype(np.float32)
            t0 = time.perf_counter()
            e.update_load(arrival_rpm=f.srv_arrival_rpm)
            e.resolve()
            lat.append(time.perf_counter() - t0)
        lat = np.array(lat[2:]) * 1e3
        print(json.dumps({"config": 5, "size_candidates": f.n_servers * f.n_acc, "tick_ms_p50": float(np.percentile(lat, 50)),
                          "tick_ms_p99": float(np.percentile(lat, 99)), "holds_10hz": bool(np.percentile(lat, 99) < 100.0),
                          "kernel_ms": e.last_kernel_ms}))


if __name__ == "__main__":
    main()
