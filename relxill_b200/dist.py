"""Multi-GPU plumbing: parameter vectors are independent, so the batch is sharded across ranks
(one process per GPU) with no data-path collective; the only exchange is the final gather of the
result spectra (NCCL all-gather over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_indices(n: int, world: int, rank: int, interleave: bool = False) -> np.ndarray:
    """Row indices of a rank: contiguous blocks, or round-robin for structured batches (parameter-grid
    sweeps) so every rank sees the same mix of cheap and expensive vectors."""
    if interleave:
        return np.arange(rank, n, world)
    lo, hi = shard_bounds(n, world, rank)
    return np.arange(lo, hi)


def sharded_eval(evaluate, params: np.ndarray, n_flux: int, interleave: bool = False) -> torch.Tensor:
    """Evaluates this rank's shard with `evaluate(params_shard) -> tensor [n_local, n_flux]` and
    all-gathers the spectra; every rank returns the full [N, n_flux] result in the input row order."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = params.shape[0]
    idx = shard_indices(n, world, rank, interleave)
    local = evaluate(params[idx])
    if world == 1:
        return local
    n_max = (n + world - 1) // world
    pad = torch.zeros((n_max, n_flux), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    gathered = torch.empty((world * n_max, n_flux), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, pad)
    out = torch.empty((n, n_flux), dtype=local.dtype, device=local.device)
    for r in range(world):
        ridx = shard_indices(n, world, r, interleave)
        out[torch.as_tensor(ridx, device=local.device)] = gathered[r * n_max: r * n_max + len(ridx)]
    return out
