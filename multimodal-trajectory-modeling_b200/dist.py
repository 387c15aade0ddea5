"""Data-parallel sharding of the trajectory axis: one process per GPU, torch.distributed for the
rendezvous only.  The per-iteration all-reduce of the sufficient statistics happens inside the
library (mtm_em_fit / mtm_mstep / mtm_score) on its own NCCL communicator."""

from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as tdist


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of the trajectory axis owned by `rank` (sizes differ by at most 1)."""
    return rank * n // world, (rank + 1) * n // world


def shard_arrays(states, observations, world: int, rank: int):
    """The rank's own (T, n_rank, d|l) slices, repacked contiguously."""
    lo, hi = shard_bounds(states.shape[1], world, rank)
    return np.ascontiguousarray(states[:, lo:hi]), np.ascontiguousarray(observations[:, lo:hi])


def global_random_init(n: int, n_clusters: int, seed: int, world: int, rank: int) -> np.ndarray:
    """The rank's slice of the reference's initial assignment rng.integers(0, n_c, n)
    (marginalizable_mixture_model.py:166-168): identical to a single-process run."""
    lo, hi = shard_bounds(n, world, rank)
    return np.random.default_rng(seed).integers(low=0, high=n_clusters, size=n)[lo:hi]


def attach(engine, group=None) -> None:
    """Create the library's NCCL communicator over the ranks of `group` (default: world)."""
    world, rank = tdist.get_world_size(group), tdist.get_rank(group)
    if world == 1:
        return
    payload = [engine.unique_id() if rank == 0 else None]
    tdist.broadcast_object_list(payload, src=0, group=group)
    engine.comm_init(payload[0], world, rank)
    # The per-iteration exchange (6.8 KB at n_c = 8) is latency-bound: use the one-shot all-reduce over NVLink
    # peer memory unless MTM_B200_ALLREDUCE=nccl.  The 64-byte IPC handles travel through torch.distributed.
    if os.environ.get("MTM_B200_ALLREDUCE", "p2p").lower() != "nccl" and world <= 16:
        handles = [None] * world
        tdist.all_gather_object(handles, engine.p2p_export(), group=group)
        engine.p2p_attach(handles, world, rank)


def active_group(group=None):
    """(world, rank) of the process group a model should shard over, or (1, 0): no initialised
    torch.distributed group, a single-rank group, or group=False (explicitly local)."""
    if group is False or not (tdist.is_available() and tdist.is_initialized()):
        return 1, 0
    return tdist.get_world_size(group), tdist.get_rank(group)


def _on_wire(t, group):
    """gloo moves host tensors, nccl device tensors."""
    return t.cpu() if tdist.get_backend(group) == "gloo" else t


def gather_last(t, n_full: int, world: int, rank: int, group=None):
    """All-gather the ragged shards of the LAST axis: t is (..., n_rank) on every rank, the result (..., n_full)
    on every rank (same device as t).  Shard r covers shard_bounds(n_full, world, r)."""
    if world == 1:
        return t
    sizes = [shard_bounds(n_full, world, r)[1] - shard_bounds(n_full, world, r)[0] for r in range(world)]
    assert int(t.shape[-1]) == sizes[rank], (tuple(t.shape), sizes, rank)
    pad = max(sizes)
    buf = t.new_zeros(tuple(t.shape[:-1]) + (pad,))
    buf[..., : sizes[rank]] = t
    buf = _on_wire(buf.contiguous(), group)
    out = [torch.empty_like(buf) for _ in range(world)]
    tdist.all_gather(out, buf, group=group)
    return torch.cat([o[..., :s] for o, s in zip(out, sizes)], dim=-1).to(t.device)


def sum_int(value: int, device, group=None) -> int:
    """Sum of a host integer over the ranks (switch counts of a sharded E_step)."""
    t = _on_wire(torch.tensor([int(value)], dtype=torch.int64, device=device), group)
    tdist.all_reduce(t, group=group)
    return int(t.item())


def detach(engine) -> None:
    if engine.world > 1:
        engine.comm_destroy()


def gather_assignments(assign_local, n: int, world: int, rank: int, group=None) -> np.ndarray:
    """All-gather the per-rank label shards into the full (n,) int64 assignment (ragged shards)."""
    if world == 1:
        return assign_local.cpu().numpy().astype(np.int64)
    sizes = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.full((pad,), -1, dtype=torch.int32, device=assign_local.device)
    buf[: assign_local.numel()] = assign_local
    out = [torch.empty_like(buf) for _ in range(world)]
    tdist.all_gather(out, buf, group=group)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)]).astype(np.int64)
