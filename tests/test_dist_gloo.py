"""N > 1 path on CPU: two gloo ranks, servers sharded round-robin, one all-gather of winners.
The local solver here is the CPU oracle (test stand-in for the per-rank GPU engine); the
sharding / packing / collective / reassembly code is the product's (parallel.py)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import oracle
    from workload_variant_autoscaler_b200 import Allocs, parallel, synth_fleet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fleet = synth_fleet(13, 3, seed=77, max_batch_choices=(2, 4, 8))

    def solve_local(shard):
        _, win = oracle.solve(shard)
        a = Allocs(shard.n_servers)
        for name in ("feasible", "acc", "replicas", "batch", "cost", "value", "itl", "ttft", "rho", "max_rate"):
            getattr(a, name)[:] = win[name]
        return a

    full = parallel.solve_sharded(solve_local, fleet, rank=rank, world=world, all_gather=parallel.torch_all_gather())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **full.columns())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_solve_matches_single(tmp_path, oracle_mod):
    from workload_variant_autoscaler_b200 import synth_fleet
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    fleet = synth_fleet(13, 3, seed=77, max_batch_choices=(2, 4, 8))
    _, win = oracle_mod.solve(fleet)
    for r in range(world):
        got = np.load(os.path.join(tmp_path, f"rank{r}.npz"))
        for name in win.dtype.names:
            want = np.ascontiguousarray(win[name])
            g = got[name]
            if want.dtype == np.float32:
                assert np.array_equal(g.view(np.uint32), want.view(np.uint32)), name
            else:
                assert np.array_equal(g.astype(np.int64), want.astype(np.int64)), name
