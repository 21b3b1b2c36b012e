"""Data-parallel plumbing: independent (image-sequence, prompt) examples are
sharded contiguously across ranks, every rank runs the full replica on its shard with no
data-path collective, and ONE all-gather of the generated ids happens at the end
(``int32 [B/W, max_new]`` per rank → rank order = dataset order).  NCCL over NVLink on the GPUs;
the same code runs on gloo for the CPU tests."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n examples for `rank` (sizes differ by at most one)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_tokens(outs: Sequence[Sequence[int]], rows: int, max_new: int, device) -> torch.Tensor:
    """Local results → int32 [rows, max_new] padded with -1 (rows ≥ len(outs): uniform across ranks)."""
    t = torch.full((rows, max_new), -1, dtype=torch.int32)
    for i, o in enumerate(outs):
        t[i, :len(o)] = torch.tensor(list(o)[:max_new], dtype=torch.int32)
    return t.to(device)


def gather_tokens(local: torch.Tensor, n_total: int) -> List[List[int]]:
    """The single end-of-run collective.  local: int32 [rows, max_new] (same shape on every rank)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        full = local[None]
    else:
        flat = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(flat, local.contiguous())       # concatenation along dim 0, rank order
        full = flat.view(world, local.shape[0], local.shape[1])
    full = full.cpu()
    outs: List[List[int]] = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        for i in range(hi - lo):
            row = full[r, i]
            outs.append([int(x) for x in row[row >= 0]])
    return outs
