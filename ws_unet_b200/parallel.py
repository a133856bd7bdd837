"""Multi-GPU driver: one process per GPU, images sharded by contiguous index range, no data-path collective;
only the per-image beta_hat (and l1) vectors are gathered once per run over NCCL.
The reference has no distributed code (its only parallelism is joblib processes, src/fabrika.py:92-100)."""
from __future__ import annotations

import os
import typing

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> typing.Tuple[int, int]:
    """Contiguous [lo, hi) slice of n images owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: typing.Optional[str] = None) -> typing.Tuple[int, int, int]:
    """torchrun-style init (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*). Returns (rank, world, local_rank)."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def gather_shards(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather the per-rank result vectors (shard_range layout) into the full length-n_total vector.
    Shards are padded to equal length so a single all_gather_into_tensor suffices (4 B per image)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert local.numel() == n_total
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_total + world - 1) // world
    lo, hi = shard_range(n_total, rank, world)
    assert local.numel() == hi - lo, (local.numel(), lo, hi)
    send = torch.zeros(per, dtype=local.dtype, device=local.device)
    send[:hi - lo] = local
    recv = torch.empty(per * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send)
    parts = []
    for r in range(world):
        l, h = shard_range(n_total, r, world)
        parts.append(recv[r * per:r * per + (h - l)])
    return torch.cat(parts)


def estimate_sharded(n_total: int, load_images: typing.Callable[[int, int], torch.Tensor],
                     estimate: typing.Callable[[torch.Tensor], torch.Tensor], chunk: int = 256) -> torch.Tensor:
    """Run `estimate` (images -> beta_hat vector) over this rank's shard in chunks and gather the full vector.
    `load_images(lo, hi)` returns the (hi-lo,1,H,W) images of global indices [lo,hi) on the rank's device."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n_total, rank, world)
    outs = []
    for s in range(lo, hi, chunk):
        outs.append(estimate(load_images(s, min(hi, s + chunk))))
    if outs:
        local = torch.cat(outs)
    else:   # world > n_total: this rank owns nothing, but it still takes part in the gather with the right device/dtype
        dev = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
        local = torch.empty(0, dtype=torch.float32, device=dev)
    return gather_shards(local, n_total)
