"""Batched dataset driver - the GPU-side counterpart of the reference's per-file iteration
(src/fabrika.py:28-205 `collect_files` / `precovers` / `stego_spatial`, src/ws/estimate.py:149-205 `run`).

The reference maps `attack` over files one at a time (optionally in 4 joblib processes). With the estimator
~1000x faster than the CPU path, decoding and staging dominate, so here files are decoded by a thread pool into
pinned uint8 batches and every batch is one fused GPU call. `run(...)` returns a DataFrame with the reference's
columns (files.csv row fields + beta_hat, channels, weighted, correct_bias, model_name).
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import pathlib
import typing

import numpy as np
import pandas as pd
import torch

from . import filters, ws
from .unet.model import UNet

NAMED_FILTERS = ws.NAMED_FILTERS


def imread_gray_u8(fname) -> np.ndarray:
    """Luma of src/_defs/imread.py:19-23 (`imread4_u8(...)[..., 3]`): cv2 BGR2GRAY; for 8-bit grayscale PNGs this is the
    stored plane (cv2 gray == PIL exactly on the shipped images)."""
    try:
        import cv2
        x = cv2.imread(str(fname))
        if x is None:
            raise IOError(fname)
        return cv2.cvtColor(x, cv2.COLOR_BGR2GRAY)
    except ImportError:
        from PIL import Image
        return np.array(Image.open(fname).convert('L'))


def _resolve_case(path: pathlib.Path) -> pathlib.Path:
    """files.csv of the shipped stego sets spell directories LSBR/HILLR while the disk has LSBr/HILLr."""
    if path.exists():
        return path
    cur = pathlib.Path(path.anchor) if path.is_absolute() else pathlib.Path('.')
    for part in path.parts[1:] if path.is_absolute() else path.parts:
        nxt = cur / part
        if not nxt.exists():
            match = [p for p in cur.iterdir() if p.name.lower() == part.lower()] if cur.is_dir() else []
            if not match:
                return path
            nxt = match[0]
        cur = nxt
    return cur


def list_files(dataset, stego_method: str = None, alpha: float = None, skip_num_images: int = None,
               take_num_images: int = None, shuffle_seed: int = None, ignore_missing: bool = True, split: str = None,
               **filt) -> pd.DataFrame:
    """Rows of the `files.csv` tables under `dataset`, selected like fabrika.precovers (stego_method None: 'images*',
    cover rows only) or fabrika.stego_spatial ('stego*', filtered by stego_method / alpha / ...). `split` names a CSV
    inside `dataset` that replaces the globbed tables (src/fabrika.py:51-52)."""
    dataset = pathlib.Path(dataset)
    if split is not None:
        df = pd.read_csv(dataset / split, dtype={'device': str})
    else:
        pattern = 'images*' if not stego_method else 'stego*'
        dfs = []
        for path in sorted(glob.glob(str(dataset / pattern))):
            try:
                dfs.append(pd.read_csv(pathlib.Path(path) / 'files.csv'))
            except Exception:
                if not ignore_missing:
                    raise
        if not dfs:
            raise FileNotFoundError(f'no files.csv under {dataset}/{pattern}')
        df = pd.concat(dfs)
    if not stego_method:
        if 'stego_method' in df:
            df = df[df['stego_method'].isna()]
    else:
        df = df[df['stego_method'] == stego_method]
        if alpha is not None:
            df = df[df['alpha'] == alpha]
    for key, val in filt.items():
        if val is not None and key in df:
            df = df[df[key] == val]
    if 'quality' in df:
        df = df[df['quality'].isna()]
    if df.empty:
        raise Exception('pre_fn() returned empty dataframe')
    df = df.sort_values('name').reset_index(drop=True)
    if shuffle_seed:
        df = df.sample(frac=1., random_state=shuffle_seed)
    if skip_num_images:
        df = df[skip_num_images:]
    if take_num_images:
        df = df[:take_num_images]
    return df


class _PinnedRing:
    """`slots` pinned uint8 batches of shape (batch, H, W), allocated once and reused (pinning memory is a blocking driver
    call that costs more than decoding the batch it would hold)."""
    _cache: typing.Dict[tuple, typing.List[torch.Tensor]] = {}

    @classmethod
    def get(cls, slots: int, batch: int, shape: tuple) -> typing.List[torch.Tensor]:
        key = (slots, batch) + tuple(shape)
        if key not in cls._cache:
            cls._cache.clear()        # one geometry at a time: the ring can hold hundreds of MB of pinned memory
            cls._cache[key] = [torch.empty((batch,) + tuple(shape), dtype=torch.uint8).pin_memory() for _ in range(slots)]
        return cls._cache[key]


def estimate_files(paths: typing.Sequence, predictor, weighted: int = 0, correct_bias: bool = False, clip: bool = True,
                   batch: int = 256, imread: typing.Callable = imread_gray_u8, device=None, workers: int = None,
                   return_l1: bool = False, slots: int = 3):
    """beta_hat (and l1) for a list of image files.

    Three stages overlap: a thread pool decodes files straight into a ring of `slots` pre-pinned uint8 batches (cv2 / PIL
    release the GIL while decoding), the host-buffer C entry points (`wsu_*_estimate_host`) copy a full batch to the GPU in
    chunks on one stream while the previous chunk computes on another, and while the calling thread sits in that call the
    pool is already decoding the next two batches. Images whose size differs from the first decoded image go through a
    per-shape fallback at the end; unreadable files yield NaN."""
    dev = filters._device(device)
    n = len(paths)
    beta = np.full(n, np.nan, dtype=np.float32)
    l1 = np.full(n, np.nan, dtype=np.float32)
    if n == 0:
        return (beta, l1) if return_l1 else beta
    workers = workers or min(32, (os.cpu_count() or 4))

    def load(p):
        try:
            x = imread(_resolve_case(pathlib.Path(p)))
            x = x[..., -1] if x.ndim == 3 else x
            return np.ascontiguousarray(x, dtype=np.uint8)
        except Exception:
            return None

    first = None
    for p0 in paths:                       # the ring's geometry comes from the first readable file
        first = load(p0)
        if first is not None:
            break
    if first is None:
        return (beta, l1) if return_l1 else beta
    shape = first.shape
    ring = _PinnedRing.get(slots, batch, shape)
    odd = {}                               # index -> image of another size

    def decode_into(slot_np, k, i):
        im = load(paths[i])
        if im is None:
            return 0
        if im.shape != shape:
            odd[i] = im
            return 0
        np.copyto(slot_np[k], im)
        return 1

    def estimate(host, idx):
        out = ws.ws_estimate_host(host, predictor, weighted=weighted, clip=clip, correct_bias=correct_bias, device=dev,
                                  return_l1=True)
        beta[idx] = out[0].numpy()
        l1[idx] = out[1].numpy()

    starts = list(range(0, n, batch))
    with concurrent.futures.ThreadPoolExecutor(workers) as pool:
        def submit(bi):
            s0 = starts[bi]
            slot_np = ring[bi % slots].numpy()
            return [pool.submit(decode_into, slot_np, k, i) for k, i in enumerate(range(s0, min(n, s0 + batch)))]

        inflight = {bi: submit(bi) for bi in range(min(slots - 1, len(starts)))}
        for bi, s0 in enumerate(starts):
            ok = np.array([f.result() for f in inflight.pop(bi)], dtype=bool)
            nxt = bi + slots - 1                      # its slot was consumed by batch bi - 1, which has completed
            if nxt < len(starts):
                inflight[nxt] = submit(nxt)
            host = ring[bi % slots][:len(ok)]
            idx = s0 + np.nonzero(ok)[0]
            if ok.all():
                estimate(host, idx)
            elif ok.any():                            # holes (unreadable / odd-sized files): compact the good rows
                estimate(host[torch.from_numpy(np.nonzero(ok)[0])].contiguous(), idx)
    by_shape = {}
    for i, im in odd.items():
        by_shape.setdefault(im.shape, []).append(i)
    for shp, idx in by_shape.items():
        host = torch.from_numpy(np.stack([odd[i] for i in idx]))
        try:
            estimate(host, np.array(idx))
        except ValueError:                            # e.g. not divisible by 2^nsteps: reported as missing, like attack()
            pass
    return (beta, l1) if return_l1 else beta


def run(input_dir, stego_method: str, alpha: float, model_name: str, model_path: str = None, channels=(3,),
        imread: typing.Callable = None, predictor=None, weighted: int = 1, correct_bias: bool = False, batch: int = 256,
        device=None, **kw) -> pd.DataFrame:
    """src/ws/estimate.py:149-205. `model_name` in NAMED_FILTERS selects a linear predictor, otherwise a UNet is loaded
    from model_path/model_name (src/unet/evaluate.py:162-188) unless `predictor` (a UNet module) is given."""
    if model_name in NAMED_FILTERS:
        pred = model_name
    else:
        if predictor is None:
            from .unet import get_pretrained
            predictor = get_pretrained(model_path, channels, model_name=model_name, device=device)
        if not isinstance(predictor, UNet):
            raise TypeError('predictor must be a ws_unet_b200 UNet')
        pred = predictor
        model_name = 'UNet'
    list_kw = {k: kw.pop(k) for k in ('skip_num_images', 'take_num_images', 'shuffle_seed', 'demosaic', 'simulator',
                                      'color_strategy', 'split') if k in kw}
    df = list_files(input_dir, stego_method=stego_method, alpha=alpha, **list_kw)
    paths = [pathlib.Path(input_dir) / nm for nm in df['name']]
    rd = imread_gray_u8 if imread is None else (lambda f: np.asarray(imread(f))[..., list(channels)][..., 0])
    beta = estimate_files(paths, pred, weighted=weighted, correct_bias=correct_bias, batch=batch, imread=rd, device=device)
    res = df.copy()
    res['name'] = [str(p) for p in paths]
    res['model_name'] = model_name
    for k, v in kw.items():
        res[k] = v
    res['beta_hat'] = beta
    res['channels'] = ''.join(map(str, channels))
    res['weighted'] = weighted
    res['correct_bias'] = correct_bias
    return res[~res.beta_hat.isna()]
