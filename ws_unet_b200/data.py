"""Synthetic covers and LSB-replacement embedding (SURVEY.md section 8d recipe, probe-verified against the
reference's own `attack`). The reference ships no generator; its stego images are LSBr with independent
random message bits at rate alpha (data/split_te.csv: simulator 'mi'), i.e. change rate alpha/2.

Host-side plumbing (torch CPU/GPU ops), not part of the measured hot path.
"""
from __future__ import annotations

import torch

COVER_SEED = 0xC0DE
STEGO_SEED = 0x57E60


def synthetic_cover(index: int, H: int = 512, W: int = 512) -> torch.Tensor:
    """One deterministic uint8 (H,W) cover: blurred uniform noise rescaled to [8,247] plus sigma=2 sensor noise."""
    g = torch.Generator().manual_seed(COVER_SEED + int(index))
    x = torch.rand(1, 1, H + 8, W + 8, generator=g) * 255
    k = torch.ones(1, 1, 5, 5) / 25
    x = torch.nn.functional.conv2d(torch.nn.functional.conv2d(x, k), k)
    x = (x - x.min()) / (x.max() - x.min()) * (247 - 8) + 8
    x = x + torch.randn(x.shape, generator=g) * 2
    return x.round().clamp(0, 255).to(torch.uint8)[0, 0]


def synthetic_covers(n: int, H: int = 512, W: int = 512, start: int = 0) -> torch.Tensor:
    return torch.stack([synthetic_cover(start + i, H, W) for i in range(n)])[:, None]


def embed_lsbr(cover: torch.Tensor, alpha: float, index: int = 0) -> torch.Tensor:
    """LSB replacement at payload alpha: a fraction alpha of pixels gets its LSB overwritten by a random bit."""
    g = torch.Generator().manual_seed(STEGO_SEED + int(index))
    mask = torch.rand(cover.shape, generator=g) < alpha
    bits = torch.randint(0, 2, cover.shape, generator=g, dtype=torch.uint8)
    return torch.where(mask, (cover & 0xFE) + bits, cover)


def synthetic_stego(n: int, alpha: float, H: int = 512, W: int = 512, start: int = 0) -> torch.Tensor:
    return torch.stack([embed_lsbr(synthetic_cover(start + i, H, W), alpha, start + i) for i in range(n)])[:, None]


def synthetic_stego_fast(n: int, alpha: float, H: int, W: int, device, seed: int = 0, unique: int = 64) -> torch.Tensor:
    """Throughput-test input: `unique` exact-recipe covers tiled to n images with fresh LSBr embedding generated on
    `device` (same distribution as synthetic_stego, different RNG stream). Returns (n,1,H,W) uint8 on `device`."""
    base = synthetic_covers(min(unique, n), H, W).to(device)
    reps = (n + base.shape[0] - 1) // base.shape[0]
    cover = base.repeat(reps, 1, 1, 1)[:n].contiguous()
    g = torch.Generator(device=device).manual_seed(STEGO_SEED + seed)
    mask = torch.rand(cover.shape, generator=g, device=device) < alpha
    bits = torch.randint(0, 2, cover.shape, generator=g, device=device, dtype=torch.uint8)
    return torch.where(mask, (cover & 0xFE) + bits, cover)


GEN_CHUNK = 64   # images per generator chunk of synthetic_stego_chunk (shards are whole chunks)


def synthetic_stego_chunk(chunk_index: int, alphas, H: int, W: int, device, chunk: int = GEN_CHUNK) -> torch.Tensor:
    """Images [chunk_index*chunk, (chunk_index+1)*chunk) of the sharded synthetic workload (BASELINE.json configs[3]),
    generated ON `device` from a generator seeded by the chunk index alone: whichever rank produces a chunk gets the same
    bytes, so the vector gathered from N GPUs can be checked bit for bit against a single-GPU run. Same recipe as
    synthetic_cover / embed_lsbr (different RNG stream); image g uses alpha = alphas[g % len(alphas)].
    Returns (chunk,1,H,W) uint8 on `device`."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(COVER_SEED * 1000003 + int(chunk_index))
    x = torch.rand(chunk, 1, H + 8, W + 8, generator=g, device=dev) * 255
    k = torch.ones(1, 1, 5, 5, device=dev) / 25
    x = torch.nn.functional.conv2d(torch.nn.functional.conv2d(x, k), k)
    lo, hi = x.amin(dim=(1, 2, 3), keepdim=True), x.amax(dim=(1, 2, 3), keepdim=True)
    x = (x - lo) / (hi - lo) * (247 - 8) + 8
    x = x + torch.randn(x.shape, generator=g, device=dev) * 2
    cover = x.round().clamp(0, 255).to(torch.uint8)
    idx = torch.arange(chunk_index * chunk, (chunk_index + 1) * chunk, device=dev) % len(alphas)
    a = torch.tensor(list(alphas), dtype=torch.float32, device=dev)[idx].view(chunk, 1, 1, 1)
    mask = torch.rand(cover.shape, generator=g, device=dev) < a
    bits = torch.randint(0, 2, cover.shape, generator=g, device=dev, dtype=torch.uint8)
    return torch.where(mask, (cover & 0xFE) + bits, cover)


def synthetic_stego_shard(first: int, n: int, alphas, H: int, W: int, device, chunk: int = GEN_CHUNK) -> torch.Tensor:
    """Images [first, first+n) of the sharded workload; `first` and `n` must be multiples of `chunk`."""
    if first % chunk or n % chunk:
        raise ValueError(f'shards are whole generator chunks of {chunk} images')
    return torch.cat([synthetic_stego_chunk(c, alphas, H, W, device, chunk) for c in range(first // chunk, (first + n) // chunk)])
