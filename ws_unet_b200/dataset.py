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
               take_num_images: int = None, shuffle_seed: int = None, ignore_missing: bool = True, **filt) -> pd.DataFrame:
    """Rows of the `files.csv` tables under `dataset`, selected like fabrika.precovers (stego_method None: 'images*',
    cover rows only) or fabrika.stego_spatial ('stego*', filtered by stego_method / alpha / ...)."""
    dataset = pathlib.Path(dataset)
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


def estimate_files(paths: typing.Sequence, predictor, weighted: int = 0, correct_bias: bool = False, clip: bool = True,
                   batch: int = 64, imread: typing.Callable = imread_gray_u8, device=None, workers: int = None,
                   return_l1: bool = False):
    """beta_hat (and l1) for a list of image files: threaded decode -> pinned uint8 batch -> one fused GPU call per batch.
    Images of a batch must share their size (the reference's sets are 512x512); unreadable files yield NaN."""
    dev = filters._device(device)
    n = len(paths)
    beta = np.full(n, np.nan, dtype=np.float32)
    l1 = np.full(n, np.nan, dtype=np.float32)
    workers = workers or min(32, (os.cpu_count() or 4))

    def load(p):
        try:
            x = imread(_resolve_case(pathlib.Path(p)))
            return x[..., -1] if x.ndim == 3 else x
        except Exception:
            return None

    with concurrent.futures.ThreadPoolExecutor(workers) as pool:
        pending = None
        for s in range(0, n + batch, batch):
            nxt = [pool.submit(load, p) for p in paths[s:s + batch]] if s < n else None
            if pending is not None:
                s0, futs = pending
                imgs = [f.result() for f in futs]
                by_shape = {}
                for i, im in enumerate(imgs):
                    if im is not None:
                        by_shape.setdefault(im.shape, []).append(i)
                for shape, idx in by_shape.items():
                    host = torch.empty((len(idx), 1) + tuple(shape), dtype=torch.uint8).pin_memory()
                    for k, i in enumerate(idx):
                        host[k, 0] = torch.from_numpy(np.ascontiguousarray(imgs[i], dtype=np.uint8))
                    b, l = ws.ws_estimate(host.to(dev, non_blocking=True), predictor, weighted=weighted, clip=clip, crop=1,
                                          correct_bias=correct_bias, return_l1=True)
                    beta[[s0 + i for i in idx]] = b.cpu().numpy()
                    l1[[s0 + i for i in idx]] = l.cpu().numpy()
            pending = (s, nxt) if nxt is not None else None
    return (beta, l1) if return_l1 else beta


def run(input_dir, stego_method: str, alpha: float, model_name: str, model_path: str = None, channels=(3,),
        imread: typing.Callable = None, predictor=None, weighted: int = 1, correct_bias: bool = False, batch: int = 64,
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
                                      'color_strategy') if k in kw}
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
