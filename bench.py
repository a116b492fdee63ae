#!/usr/bin/env python
"""bench.py — headline benchmark of the WVA optimizer hot path on B200.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
100 models x 4 accelerator types x 256 batch sizes x 64 replica levels = 6,553,600
candidate allocations per step per GPU (synthetic fleet, SURVEY.md §8d, PCG64 seed 42).
One step = one pass of the hot path over that grid: state-dependent M/M/1/K evaluation
of every cell + SLO feasibility + cost + transition penalty + per-model argmin.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun, one rank per GPU: the fleet is sharded by model (weak
scaling: 100 models per GPU), the only exchange is ONE NCCL all-gather of the per-shard
winner records (40 B per model) after the local solve.

The JSON line printed by rank 0 follows the driver's contract; `value` is device-resident
throughput (inputs already in HBM), `e2e` goes through the C ABI with host buffers
(H2D of the fleet + D2H of the winners inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate allocations/sec"
UNIT = "candidates/s"
N_MODELS, N_ACC, N_BATCH, N_REPLICAS = 100, 4, 256, 64
ALGO_BYTES_PER_CELL = 100.0   # compulsory I/O per candidate (HBM view)
ALGO_F64_PER_STATE = 7.0      # reference fp64 ops per state (K mul, 2K div, 4K add/mul)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--cpu-sample-pairs", type=int, default=32,
                    help="(model, accelerator) pairs (16384 cells each) timed for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(rank: int):
    from workload_variant_autoscaler_b200 import config2_grid, synth_fleet
    # weak scaling = fixed work per GPU: every rank draws its 100 models from the same seeded stream (the
    # grid kernel's time varies by +-30 % with the draw, which would otherwise be measured as 'scaling')
    fleet = synth_fleet(N_MODELS, N_ACC, seed=42)
    grid = config2_grid(N_BATCH, N_REPLICAS)
    return fleet, grid


def config_dict(n_gpus: int) -> dict:
    return {
        "workload": "BASELINE configs[1]: 100 models x 4 accelerator types x 256 batch sizes x 64 replica levels "
                    "per GPU, state-dependent M/M/1/K Analyze per cell + per-model min-value SLO-feasible argmin",
        "cells_per_gpu": N_MODELS * N_ACC * N_BATCH * N_REPLICAS,
        "models_per_gpu": N_MODELS, "accelerators": N_ACC, "batch_sizes": N_BATCH, "replica_levels": N_REPLICAS,
        "mean_states_per_cell": 11 * (N_BATCH + 1) / 2,
        "seed": 42, "l2_flush_between_steps": True,
        "parallelism": f"dp{n_gpus} over models, one all-gather of the winner blocks per step" if n_gpus > 1 else "single GPU",
        "cell_table_materialised": False,
    }


# ----------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ----------------------------------------------------------------------------
def _cpu_worker(args):
    rank_seed, pairs = args
    import oracle
    from workload_variant_autoscaler_b200 import config2_grid, synth_fleet
    fleet = synth_fleet(N_MODELS, N_ACC, seed=rank_seed)
    grid = config2_grid(N_BATCH, N_REPLICAS)
    per = N_BATCH * N_REPLICAS
    n = 0
    for p in pairs:
        oracle.grid_cells(fleet, grid, p * per, (p + 1) * per)
        n += per
    return n


def sample_pairs(n_pairs: int):
    total = N_MODELS * N_ACC
    n_pairs = max(1, min(n_pairs, total))
    return [int(i * total / n_pairs) for i in range(n_pairs)]


def cpu_baseline_port(n_pairs: int) -> dict:
    """Oracle (C port of the Go reference) on ONE core, bounded sample of the same workload."""
    import oracle
    oracle.build()
    pairs = sample_pairs(n_pairs)
    _cpu_worker((42, pairs[:1]))  # warm caches / page in
    t0 = time.perf_counter()
    n = _cpu_worker((42, pairs))
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{len(pairs)} of {N_MODELS * N_ACC} (model, accelerator) pairs x all 256x64 cells "
                      f"= {n} cells in {dt:.2f} s; C restatement of the Go reference (single goroutine), "
                      f"not the Go binary (no Go toolchain on the box)"}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; Go cannot be built here) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    cores = min(cores, 64)
    pairs_per_step = sample_pairs(max(cores, 8))  # one (model, accelerator) pair = 16384 cells per worker per step
    chunks = [pairs_per_step[i::cores] for i in range(cores)]
    chunks = [c for c in chunks if c]
    per_step_cells = len(pairs_per_step) * N_BATCH * N_REPLICAS
    ctx = mp.get_context("fork")
    with ctx.Pool(len(chunks)) as pool:
        def step():
            return sum(pool.map(_cpu_worker, [(42, c) for c in chunks]))
        for _ in range(max(args.warmup, 1)):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    value = per_step_cells * args.steps / dt
    sample = (f"each step = {len(pairs_per_step)} of {N_MODELS * N_ACC} (model, accelerator) pairs x all 256x64 cells "
              f"= {per_step_cells} cells, spread over {len(chunks)} processes; C restatement of the Go reference "
              f"(the Go path itself is single-goroutine and cannot be built here: no Go toolchain)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": len(chunks), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------
# GPU leg
# ----------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from workload_variant_autoscaler_b200 import Engine, _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(local_rank)
    fleet, grid = workload(rank)
    S = fleet.n_servers
    n_cells = S * fleet.n_acc * grid.batch.size * grid.replicas.size

    # device-resident winner block: 10 columns x S x 4 B (the feasible column uses S bytes of its slot)
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    win_local = torch.zeros(10 * S, dtype=torch.int32, device=dev)
    win_all = torch.zeros(10 * S * world, dtype=torch.int32, device=dev) if world > 1 else None
    base = win_local.data_ptr()
    import ctypes as C
    cols = _abi.AllocsC()
    cols.feasible = C.cast(base, _abi.u8p)
    for k, name in enumerate(("acc", "replicas", "batch"), start=1):
        setattr(cols, name, C.cast(base + 4 * S * k, _abi.i32p))
    for k, name in enumerate(("cost", "value", "itl", "ttft", "rho", "max_rate"), start=4):
        setattr(cols, name, C.cast(base + 4 * S * k, _abi.f32p))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    eng.upload(fleet)  # inputs resident in HBM before the timed region

    # N > 1: one all-gather of the 4 KB winner blocks per step.  Default: NCCL.  WVA_BENCH_PEER=1 selects the
    # library's own exchange kernel (wva_xchg_publish: NVLink stores into every peer's buffer + epoch flags, one
    # launch); measured at N = 2 it is as fast as NCCL's low-latency path (0.405 vs 0.403 ms/step), so the
    # proven collective stays the default.
    xchg, exchange = None, "none (single GPU)"
    if world > 1:
        exchange = "nccl all_gather_into_tensor"
        try:
            if not os.environ.get("WVA_BENCH_PEER"):
                raise RuntimeError("not requested")
            from workload_variant_autoscaler_b200.parallel import PeerExchange
            xchg = PeerExchange(eng, win_local.numel() * 4)
            exchange = "peer-memory kernel (wva_xchg_publish): NVLink stores + epoch flags, one launch"
        except Exception as exc:  # noqa: BLE001 - report and use the library collective
            xchg = None
            if os.environ.get("WVA_BENCH_PEER"):
                exchange += f" (peer exchange unavailable: {exc})"
        ok = torch.tensor([1 if xchg is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            xchg = None

    def step_device():
        eng.grid_solve_device(grid, cols)
        if world > 1:
            if xchg is not None:
                xchg.publish(base)
            else:
                dist.all_gather_into_tensor(win_all, win_local)

    if xchg is not None:
        # one checked step: the peer exchange must deliver exactly what the library collective delivers
        class _Raw:  # zero-copy view of the gathered device buffer
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}
        with torch.cuda.stream(ext):
            eng.grid_solve_device(grid, cols)
            gptr, stride = xchg.publish(base)
            dist.all_gather_into_tensor(win_all, win_local)
        torch.cuda.synchronize()
        got = torch.as_tensor(_Raw(gptr, world * stride // 4), device=dev).view(world, stride // 4)[:, : win_local.numel()]
        if xchg.error() or not torch.equal(got.reshape(-1), win_all):
            raise SystemExit("peer exchange delivered a different gathered block than NCCL")

    def timed_steps(k):
        """K steps, each timed with CUDA events on the launching stream; L2 flushed in between."""
        evs = []
        with torch.cuda.stream(ext):
            for _ in range(k):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step_device()
                e1.record()
                evs.append((e0, e1))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    timed_steps(max(args.warmup, 3))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # a step is < 1 ms: keep the same load running for ~1.2 s before the timed region so that the
    # 100 ms nvidia-smi samples are taken under this workload (they continue through the timed region)
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 1.2:
        timed_steps(20)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = eng.launch_count
    wall0 = time.perf_counter()
    per_step = timed_steps(args.steps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launch_count - launches0
    total_ms = torch.tensor([sum(per_step)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    clocks = sampler.stop() if rank == 0 else None

    # dominant kernel: per-launch device time measured live (CUDA events on the engine's stream)
    k_ms = []
    for _ in range(5):
        with torch.cuda.stream(ext):
            flush.fill_(1)
        torch.cuda.synchronize()
        fleet_c, grid_c = fleet, grid
        eng.grid_solve(fleet_c, grid_c)  # records the grid kernel's own event pair
        k_ms.append(eng.last_kernel_ms)
    kernel_ms = statistics.median(k_ms)

    # e2e through the C ABI with host buffers (H2D fleet + D2H winners inside the timed region)
    from workload_variant_autoscaler_b200._abi import Allocs
    win = Allocs(fleet.n_servers)  # a reconcile loop keeps one winner block and hands it to every call
    for _ in range(2):
        eng.grid_solve(fleet, grid, out=win)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.grid_solve(fleet, grid, out=win)
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    h2d = sum(getattr(fleet, n).nbytes for n in fleet._F32 + fleet._I32 + fleet._U8) + grid.batch.nbytes * 2 + \
        grid.replicas.nbytes
    d2h = sum(v.nbytes for v in win.columns().values()) + 32

    if rank == 0:
        peaks, prof = {}, {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        try:  # ncu-derived figures of the dominant kernel (profiles/, committed per round)
            prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_latest.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        achieved = ALGO_BYTES_PER_CELL * n_cells / (kernel_ms * 1e-3) / 1e9
        states = float(config_dict(1)["mean_states_per_cell"]) * n_cells
        value = n_cells * world * args.steps / (total_ms * 1e-3)
        fp64_peak = float(prof.get("fp64_peak_tdfma_s", 17.07))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_dict(world), exchange=exchange),
            "e2e": {"value": n_cells * world * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h),
                    "note": "wva_grid_solve through the C ABI with host buffers: H2D of the fleet (staged through the "
                            "library's pinned arena) and D2H of the winner block inside the timed region"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": prof.get("grid_kernel_dram_bytes"), "peak_source": peak_src, "kernel": "grid_kernel",
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CELL * n_cells,
                "note": "achieved = 100 B/cell (SURVEY.md 8d) x cells / live CUDA-event time of grid_kernel alone. "
                        "The kernel is FP64-issue bound, not HBM bound; its actual DRAM traffic (traffic) is far "
                        "below the algorithmic bytes because inputs are factored tables, see fp64",
                "fp64": {
                    "peak_tdfma_per_s": fp64_peak, "peak_source": "tools/fp64_peak.cu on this pool (profiles/r01_fp64_peak.txt)",
                    "pipe_active_pct_ncu": prof.get("grid_kernel_fp64_pipe_pct"),
                    "issue_active_pct_ncu": prof.get("grid_kernel_issue_pct"),
                    "reference_ops_per_launch": ALGO_F64_PER_STATE * states,
                    "reference_equivalent_tflops": ALGO_F64_PER_STATE * states / (kernel_ms * 1e-3) / 1e12,
                    "note": "reference_equivalent counts the reference's 7*K float64 ops per solve; it exceeds the "
                            "pipe peak because exact early termination executes ~2% of the reference's state steps "
                            "(DESIGN.md 3.2); pipe_active_pct is what the hardware actually issued",
                },
            },
            "wall_s_timed_region": wall,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_port(args.cpu_sample_pairs)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
