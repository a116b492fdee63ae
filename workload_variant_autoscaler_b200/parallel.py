"""Multi-GPU driver: one process per GPU, servers sharded round-robin, ONE all-gather.

Unlimited mode is separable per server (pkg/solver/solver.go:63-79): every rank analyses and
solves its own servers with no data-path exchange, then the fixed-size winner records
(10 x 4 B per server) are all-gathered (NCCL over NVLink on GPUs; gloo in the CPU tests) and
every rank ends up with the full ``AllocationSolution``.  The accelerator / model tables are
tiny and simply replicated.
"""
from __future__ import annotations

import numpy as np

from ._abi import ALLOC_COLUMNS, Allocs
from .fleet import Fleet


def shard_indices(n_servers: int, rank: int, world: int) -> np.ndarray:
    return np.arange(n_servers)[rank::world]


def pack_winners(win: Allocs, pad_to: int) -> np.ndarray:
    """Allocs -> int32 [10, pad_to] (float columns bit-cast), padding = infeasible records."""
    out = np.zeros((len(ALLOC_COLUMNS), pad_to), np.int32)
    for k, (name, dt) in enumerate(ALLOC_COLUMNS):
        col = getattr(win, name)
        out[k, : win.n] = col.astype(np.int32) if dt is np.uint8 else col.view(np.int32)
    return out


def unpack_winners(blocks: np.ndarray, n_servers: int, world: int) -> Allocs:
    """int32 [world, 10, pad] gathered blocks -> Allocs in the original server order."""
    out = Allocs(n_servers)
    for r in range(world):
        idx = shard_indices(n_servers, r, world)
        for k, (name, dt) in enumerate(ALLOC_COLUMNS):
            col = blocks[r, k, : idx.size]
            getattr(out, name)[idx] = col.astype(np.uint8) if dt is np.uint8 else col.view(dt)
    return out


def solve_sharded(solve_local, fleet: Fleet, *, rank: int, world: int, all_gather) -> Allocs:
    """Solve ``fleet`` across ``world`` ranks.

    solve_local(shard: Fleet) -> Allocs of the shard's winners (the engine on this rank's GPU);
    all_gather(block: np.ndarray[int32]) -> np.ndarray [world, ...] (one collective).
    """
    if not fleet.unlimited:
        # the greedy pass shares ONE capacity map (pkg/solver/greedy.go:107-166): per-shard greedy solves would
        # each spend the full capacity.  Limited mode shards the candidate generation only (solve_sharded_limited).
        raise ValueError("solve_sharded needs an unlimited fleet; use solve_sharded_limited for greedy mode")
    shard = fleet.shard(rank, world)
    win = solve_local(shard)
    pad = (fleet.n_servers + world - 1) // world
    gathered = all_gather(pack_winners(win, pad))
    return unpack_winners(np.asarray(gathered), fleet.n_servers, world)


def pack_candidates(cand: Allocs, pad_to: int) -> np.ndarray:
    """Candidate table Allocs [n] -> int32 [10, pad_to] (float columns bit-cast), padding = nil candidates."""
    return pack_winners(cand, pad_to)


def solve_sharded_limited(analyze_local, fleet: Fleet, *, rank: int, world: int, all_gather, greedy=None):
    """Limited (greedy) mode across ``world`` ranks: candidate generation — the expensive part —
    shards over servers exactly like the unlimited solve; the greedy pass is ONE sequential walk over a shared
    capacity map (pkg/solver/greedy.go:107-166) and does not shard, so the per-shard candidate tables are
    all-gathered (one collective, S x A x 40 B) and every rank runs the same deterministic ``SolveGreedy`` over the
    full table: every rank ends with the identical solution and no second collective is needed.

    analyze_local(shard: Fleet) -> Allocs [S_shard * A] (``Server.Calculate``: the engine's ``analyze`` on this
    rank's GPU); greedy(fleet, candidates) -> (candidates, winners), default = the library's ``wva_solve_greedy``.
    Returns (candidates [S * A] after best-effort scaling, winners [S]) in the original server order.
    """
    from .engine import greedy_solve
    A = fleet.n_acc
    shard = fleet.shard(rank, world)
    cand_local = analyze_local(shard)
    pad = (fleet.n_servers + world - 1) // world
    gathered = np.asarray(all_gather(pack_candidates(cand_local, pad * A)))  # [world, 10, pad * A]
    full = Allocs(fleet.n_servers * A)
    for r in range(world):
        idx = shard_indices(fleet.n_servers, r, world)
        for k, (name, dt) in enumerate(ALLOC_COLUMNS):
            col = gathered[r, k, : idx.size * A].reshape(idx.size, A)
            dst = getattr(full, name).reshape(fleet.n_servers, A)
            dst[idx] = col.astype(np.uint8) if dt is np.uint8 else col.view(dt)
    return (greedy or greedy_solve)(fleet, full)


def torch_all_gather(group=None, device=None):
    """all_gather callable over torch.distributed (nccl on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    def _ag(block: np.ndarray) -> np.ndarray:
        world = dist.get_world_size(group)
        t = torch.from_numpy(np.ascontiguousarray(block)).reshape(-1)
        if device is not None:
            t = t.to(device)
        out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=group)
        return out.cpu().numpy().reshape((world,) + tuple(block.shape))

    return _ag


class PeerExchange:
    """All-gather of a fixed-size device block over peer memory (``wva_xchg_*``, include/wva_b200.h).

    One process per GPU on one node.  The 64-byte CUDA IPC handles of the ranks' gathered buffers are
    exchanged once over ``torch.distributed`` (host side); after that every step is ONE kernel on the
    engine's stream that writes this rank's block into all peers over NVLink and waits for theirs —
    no library collective, no extra stream hand-over.
    """

    def __init__(self, engine, block_bytes: int, group=None):
        import ctypes as C

        import torch.distributed as dist

        self._e, self._L, self._C = engine, engine._L, C
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.block_bytes = int(block_bytes)
        mine = C.create_string_buffer(64)
        engine._check(self._L.wva_xchg_create(engine._h, self.world, self.rank, self.block_bytes, mine))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine.raw), group=group)
        for r, hb in enumerate(handles):
            engine._check(self._L.wva_xchg_open(engine._h, r, C.create_string_buffer(hb, 64)))
        dist.barrier(group)  # every rank has mapped every buffer before the first publish

    def publish(self, src_device_ptr: int):
        """Enqueue the exchange of the block at ``src_device_ptr``; returns (device address of the gathered
        [world][slot_stride] buffer of this step, slot_stride)."""
        C = self._C
        out, stride = C.c_void_p(), C.c_size_t()
        self._e._check(self._L.wva_xchg_publish(self._e._h, C.c_void_p(src_device_ptr), C.byref(out), C.byref(stride)))
        return int(out.value), int(stride.value)

    def error(self) -> int:
        return int(self._L.wva_xchg_error(self._e._h))

    def close(self):
        self._L.wva_xchg_destroy(self._e._h)
