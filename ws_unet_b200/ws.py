"""Weighted-Stego (WS) LSB-replacement change-rate estimator - drop-in for src/ws/estimate.py:31-136.

Primary entry points are batched and stay on the GPU:
  * `ws_estimate(images, predictor, ...)`   - predictor is a `ws_unet_b200.unet.UNet` module or the name of a
    linear filter ('KB', 'AVG', 'AVG9', '1'); one fused pass, returns beta_hat (B,) [and l1 (B,)].
  * `ws_from_prediction(images, x_hat, ...)` - any externally computed prediction.
`attack(fname, ...)` keeps the reference's per-file signature and return dict.
"""
from __future__ import annotations

import ctypes
import typing

import numpy as np
import torch

from . import _native, filters
from .unet.model import UNet

NAMED_FILTERS = filters.NAMED_FILTERS_2D  # src/ws/estimate.py:31-52 (same tables)


def _prep_images(images: torch.Tensor):
    if not images.is_cuda:
        raise RuntimeError("images must live on a CUDA device (no CPU fallback)")
    if images.dim() == 3:
        images = images[:, None]
    if images.dim() != 4 or images.shape[1] != 1:
        raise ValueError(f"expected (B,1,H,W) or (B,H,W) images, got {tuple(images.shape)}")
    if images.dtype == torch.uint8:
        dtype = _native.WSU_U8
    else:
        dtype = _native.WSU_F32
        images = images.to(torch.float32)
    return images.contiguous(), dtype


def ws_estimate(images: torch.Tensor, predictor, weighted: int = 0, clip: bool = True, crop: int = 1,
                correct_bias: bool = False, return_l1: bool = False, return_prediction: bool = False):
    """Batched WS estimator.

    images: (B,1,H,W) uint8 pixels (or float32 in [0,1] as in WSLoss, src/_defs/losses.py:46-61).
    weighted: 1 -> 1/(5+var), -1 -> 5+var, 0 -> uniform (src/ws/estimate.py:93-110).
    clip: clip beta_hat at 0 (attack: True, predict_unet: False).  crop: 1 -> interior only, 0 -> whole image.
    return_prediction: also return the predictor's output as the predictor itself defines it - a linear filter gives
    pixel units on the 'valid' interior, (B,H-2,W-2) (filters/evaluate.py:136-140); a UNet gives its sigmoid output in
    (0,1) on the whole image, (B,1,H,W) (unet.py:189; infere_single crops and scales it, evaluate.py:49-50).
    """
    images, dtype = _prep_images(images)
    B, _, H, W = images.shape
    dev = images.device
    lib = _native.load()
    beta = torch.empty(B, dtype=torch.float32, device=dev)
    l1 = torch.empty(B, dtype=torch.float32, device=dev) if return_l1 else None
    l1_ptr = ctypes.c_void_p(l1.data_ptr()) if l1 is not None else None
    with torch.cuda.device(dev):
        st = _native.stream_ptr(dev)
        if isinstance(predictor, str):
            if predictor not in _native.PRED_KINDS:
                raise KeyError(predictor)
            if crop != 1:
                raise ValueError("linear filters are 'valid' convolutions: crop must be 1")
            _native.check(lib.wsu_filter_ws_estimate(
                dev.index, ctypes.c_void_p(images.data_ptr()), dtype, _native.PRED_KINDS[predictor], int(weighted),
                int(bool(clip)), int(bool(correct_bias)), ctypes.c_void_p(beta.data_ptr()), l1_ptr,
                B, H, W, st), 'wsu_filter_ws_estimate')
            pred = filters.filter_predict(images, predictor) if return_prediction else None
        elif isinstance(predictor, UNet):
            h = predictor.native_handle(dev)
            yhat = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev) if return_prediction else None
            if correct_bias and dtype != _native.WSU_U8:
                raise ValueError("correct_bias needs uint8 images")
            # correct_bias (estimate.py:126-128): the library runs the predictor a second time on the LSB-difference image
            # x_bar - x (formed inside the first-layer kernel, scaled by 1/255 like any input, src/unet/evaluate.py:45) and its
            # head accumulates sum w (x - x_bar) x_bias next to the first pass's sums - one C call, no prediction map in HBM
            _native.check(lib.wsu_unet_ws_estimate(
                h, ctypes.c_void_p(images.data_ptr()), dtype, B, H, W, int(weighted), int(bool(clip)), int(crop),
                int(bool(correct_bias)), ctypes.c_void_p(beta.data_ptr()), l1_ptr,
                ctypes.c_void_p(yhat.data_ptr()) if yhat is not None else None, st), 'wsu_unet_ws_estimate')
            pred = yhat
        else:
            raise TypeError("predictor must be a ws_unet_b200 UNet or one of " + str(list(_native.PRED_KINDS)))
    out = (beta,)
    if return_l1:
        out += (l1,)
    if return_prediction:
        out += (pred,)
    return out[0] if len(out) == 1 else out


def ws_estimate_host(images: torch.Tensor, predictor, weighted: int = 0, clip: bool = True, correct_bias: bool = False,
                     device=None, return_l1: bool = False):
    """End-to-end variant for HOST images: (B,1,H,W) or (B,H,W) uint8 CPU tensor (pinned memory makes the copies
    asynchronous). The library copies chunks H2D on a side stream while the previous chunk computes and copies
    beta_hat / l1 back; returns CPU tensors. crop=1 semantics (attack / predict_unet)."""
    if images.is_cuda or images.dtype != torch.uint8:
        raise ValueError("ws_estimate_host takes a uint8 CPU tensor")
    if images.dim() == 4:
        images = images[:, 0]
    images = images.contiguous()
    B, H, W = images.shape
    dev = filters._device(device)
    lib = _native.load()
    beta = torch.empty(B, dtype=torch.float32)
    l1 = torch.empty(B, dtype=torch.float32)
    with torch.cuda.device(dev):
        if isinstance(predictor, str):
            if predictor not in _native.PRED_KINDS:
                raise KeyError(predictor)
            _native.check(lib.wsu_filter_ws_estimate_host(
                dev.index, ctypes.c_void_p(images.data_ptr()), _native.PRED_KINDS[predictor], int(weighted), int(bool(clip)),
                int(bool(correct_bias)), ctypes.c_void_p(beta.data_ptr()), ctypes.c_void_p(l1.data_ptr()) if return_l1 else None,
                B, H, W), 'wsu_filter_ws_estimate_host')
        elif isinstance(predictor, UNet):
            _native.check(lib.wsu_unet_ws_estimate_host(
                predictor.native_handle(dev), ctypes.c_void_p(images.data_ptr()), B, H, W, int(weighted), int(bool(clip)), 1,
                int(bool(correct_bias)), ctypes.c_void_p(beta.data_ptr()), ctypes.c_void_p(l1.data_ptr()) if return_l1 else None),
                'wsu_unet_ws_estimate_host')
        else:
            raise TypeError("predictor must be a ws_unet_b200 UNet or one of " + str(list(_native.PRED_KINDS)))
    return (beta, l1) if return_l1 else beta


def ws_from_prediction(images: torch.Tensor, x_hat: torch.Tensor, weighted: int = 0, clip: bool = True, crop: int = 1,
                       x_bias: typing.Optional[torch.Tensor] = None, return_l1: bool = False):
    """WS reduction against caller-supplied predictions in pixel units: x_hat (B,H,W) or cropped (B,H-2,W-2)."""
    images, dtype = _prep_images(images)
    B, _, H, W = images.shape
    x_hat = x_hat.to(torch.float32).reshape(B, *x_hat.shape[-2:]).contiguous()
    cropped = int(tuple(x_hat.shape[-2:]) == (H - 2, W - 2))
    if not cropped and tuple(x_hat.shape[-2:]) != (H, W):
        raise ValueError("x_hat must be (B,H,W) or (B,H-2,W-2)")
    if x_bias is not None:
        x_bias = x_bias.to(torch.float32).reshape(x_hat.shape).contiguous()
    dev = images.device
    beta = torch.empty(B, dtype=torch.float32, device=dev)
    l1 = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _native.check(_native.load().wsu_ws_from_prediction(
            dev.index, ctypes.c_void_p(images.data_ptr()), dtype, ctypes.c_void_p(x_hat.data_ptr()), cropped,
            ctypes.c_void_p(x_bias.data_ptr()) if x_bias is not None else None, int(weighted), int(bool(clip)), int(crop),
            ctypes.c_void_p(beta.data_ptr()), ctypes.c_void_p(l1.data_ptr()), B, H, W, _native.stream_ptr(dev)),
            'wsu_ws_from_prediction')
    return (beta, l1) if return_l1 else beta


def attack(fname: str, channels: typing.List[int], pixel_estimator, mean_estimator: np.ndarray = None,
           correct_bias: bool = False, weighted: int = 1, imread: typing.Callable = None,
           process_image: typing.Callable = None, device=None, **kw) -> dict:
    """src/ws/estimate.py:55-136, same arguments and return dict.

    `pixel_estimator` may be a predictor understood by `ws_estimate` (UNet module or filter name: fused GPU path)
    or any reference-style callable (H,W,C) float32 -> (H-2,W-2,1) float32; the latter is evaluated as given and only
    the WS arithmetic runs on the GPU. `mean_estimator` must be AVG (the only value the reference passes).
    """
    if mean_estimator is not None and not np.array_equal(np.asarray(mean_estimator), NAMED_FILTERS['AVG']):
        raise NotImplementedError("only mean_estimator=AVG is implemented (the reference never passes another)")
    dev = filters._device(device)
    x = imread(fname)                      # uint8 (H,W,C)
    xp = process_image(x) if process_image is not None else x[..., channels].astype('float32')
    x0 = np.ascontiguousarray(xp[..., 0])
    if not np.array_equal(x0, np.round(x0)) or x0.min() < 0 or x0.max() > 255:
        raise ValueError("attack expects integer pixel values in 0..255")
    img = torch.from_numpy(x0.astype(np.uint8)).to(dev)[None, None]
    # estimate.py:93,109: any `weighted` other than +-1 means uniform weights (abs(int(weighted)) == 1 selects the weighted arm)
    w_mode = int(weighted) if abs(int(weighted)) == 1 else 0
    try:
        if isinstance(pixel_estimator, (str, UNet)):
            if isinstance(pixel_estimator, UNet) and tuple(img.shape[-2:]) != (512, 512):
                # the reference's UNet estimator goes through CenterCrop(512) (src/unet/evaluate.py:46, loader.py:43-44):
                # its 510x510 prediction does not broadcast against any other interior, the ValueError is caught below
                # and the file is reported with beta_hat=None (estimate.py:115-118). Same here; the batched
                # `ws_estimate` takes any H, W divisible by 2^nsteps.
                raise ValueError(f'UNet pixel estimator expects 512x512 images, got {tuple(img.shape[-2:])}')
            beta = ws_estimate(img, pixel_estimator, weighted=w_mode, clip=True, crop=1, correct_bias=correct_bias)
        else:
            x1_hat = torch.from_numpy(np.ascontiguousarray(pixel_estimator(xp)[..., 0], dtype=np.float32)).to(dev)[None]
            x_bias = None
            if correct_bias:
                x_bar = process_image(x ^ 1) if process_image is not None else (x ^ 1)[..., channels].astype('float32')
                x_bias = torch.from_numpy(np.ascontiguousarray(pixel_estimator(x_bar - xp)[..., 0], dtype=np.float32)).to(dev)[None]
            beta = ws_from_prediction(img, x1_hat, weighted=w_mode, clip=True, crop=1, x_bias=x_bias)
        beta_hat = np.float32(beta.item())
    except ValueError:
        beta_hat = None
    return kw | {
        'beta_hat': beta_hat,
        'channels': ''.join(map(str, channels)),
        'weighted': weighted,
        'correct_bias': correct_bias,
    }
