"""Linear pixel predictors (KB, AVG, AVG9, identity) - drop-in for src/filters/evaluate.py:22-50,118-146.

`get_filter_estimator(name, flatten=False)` returns the same kind of callable the reference returns:
(H,W,C>=1) float32 pixels -> (H-2,W-2,1) float32 pixels, computed by libwsunet's stencil kernel on the GPU
(exact fp32 integer arithmetic instead of the reference's FFT path).
"""
from __future__ import annotations

import ctypes
import typing

import numpy as np
import torch

from . import _native

# coefficient tables kept for callers that read them (src/filters/evaluate.py:22-50)
NAMED_FILTERS = {
    'KB': np.array([[-1], [+2], [-1], [+2], [-1], [+2], [-1], [+2]], dtype='float64') / 4.,
    'AVG': np.ones((8, 1)) / 8.,
}
NAMED_FILTERS_2D = {
    'KB': np.array([[[-1, +2, -1], [+2, 0, +2], [-1, +2, -1]]], dtype='float32').T / 4.,
    'AVG': np.array([[[1, 1, 1], [1, 0, 1], [1, 1, 1]]], dtype='float32').T / 8.,
    'AVG9': np.array([[[1, 1, 1], [1, 1, 1], [1, 1, 1]]], dtype='float32').T / 9.,
    '1': np.array([[[0, 0, 0], [0, 1, 0], [0, 0, 0]]], dtype='float32').T / 1.,
}


def get_coefficients(filter_name: str, flatten: bool = True) -> np.ndarray:
    return NAMED_FILTERS[filter_name] if flatten else NAMED_FILTERS_2D[filter_name]


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("ws_unet_b200 needs a CUDA device (no CPU fallback)")
    return torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)


def filter_predict(images: torch.Tensor, filter_name: str) -> torch.Tensor:
    """Batched predictor: images (B,H,W) or (B,1,H,W), uint8 pixels or float32 in [0,1], on a CUDA device
    -> (B,H-2,W-2) float32 predictions in pixel units."""
    if filter_name not in _native.PRED_KINDS:
        raise KeyError(filter_name)
    if not images.is_cuda:
        raise RuntimeError("images must live on a CUDA device")
    if images.dim() == 4:
        images = images[:, 0]
    dtype = _native.WSU_U8 if images.dtype == torch.uint8 else _native.WSU_F32
    if dtype == _native.WSU_F32:
        images = images.to(torch.float32)
    images = images.contiguous()
    B, H, W = images.shape
    out = torch.empty((B, H - 2, W - 2), dtype=torch.float32, device=images.device)
    with torch.cuda.device(images.device):
        _native.check(_native.load().wsu_filter_predict(
            images.device.index, ctypes.c_void_p(images.data_ptr()), dtype, _native.PRED_KINDS[filter_name],
            ctypes.c_void_p(out.data_ptr()), B, H, W, _native.stream_ptr(images.device)), 'wsu_filter_predict')
    return out


def filter_residuals(images: torch.Tensor, filter_name: str) -> torch.Tensor:
    """Batched residual map y - y_hat in float64, (B,H-2,W-2): what get_filter_residuals returns per image, reshaped.
    For uint8 pixels the KB/AVG prediction is an integer multiple of 1/4 (1/8) and exact in float32, so the float64
    residual is exact."""
    if images.dim() == 4:
        images = images[:, 0]
    pred = filter_predict(images, filter_name).to(torch.float64)
    centre = images[:, 1:-1, 1:-1]
    centre = centre.to(torch.float64) if images.dtype == torch.uint8 else centre.to(torch.float64) * 255.
    return centre - pred


def get_filter_residuals(fname: str, filter: np.ndarray, process_image: typing.Callable, imread: typing.Callable = None,
                         device=None, **kw) -> np.ndarray:
    """src/filters/evaluate.py:53-76: residual of the linear predictor `filter` (8 x 1, float64) over the neighbour
    matrix `process_image(imread(fname))` (N x 9, last column = target) -> (N, 1) float64. Any coefficient vector is
    allowed (this is the OLS-fit path, not the per-image hot path): `wsu_filter_residual_rows` accumulates each row in
    float64 in the reference's column order. For the named KB / AVG vectors on 8-bit pixels every term is a multiple of
    1/8 below 2^11, so the result is exact and equals `filter_residuals(image, name)` (the stencil kernel)."""
    from . import defs
    dev = _device(device)
    mat = np.ascontiguousarray(np.asarray(process_image((imread or defs.imread4_u8)(fname))))
    if mat.ndim != 2 or mat.shape[1] != 9:
        raise ValueError(f'expected an N x 9 neighbour matrix, got {mat.shape}')
    code = {np.dtype('uint8'): 0, np.dtype('float32'): 1, np.dtype('float64'): 2}.get(mat.dtype)
    if code is None:
        mat, code = mat.astype(np.float64), 2
    mt = torch.from_numpy(mat).to(dev)
    coef = torch.from_numpy(np.ascontiguousarray(np.asarray(filter, dtype=np.float64).reshape(-1))).to(dev)
    if coef.numel() != 8:
        raise ValueError('filter must hold 8 coefficients (src/filters/evaluate.py:22-27)')
    out = torch.empty(mat.shape[0], dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _native.check(_native.load().wsu_filter_residual_rows(
            dev.index, ctypes.c_void_p(mt.data_ptr()), code, ctypes.c_void_p(coef.data_ptr()), ctypes.c_void_p(out.data_ptr()),
            mat.shape[0], _native.stream_ptr(dev)), 'wsu_filter_residual_rows')
    return out.cpu().numpy()[:, None]


def infere_single(x: np.ndarray, filter_name: str, device=None) -> np.ndarray:
    """src/filters/evaluate.py:136-141: (H,W,C) float32 pixel units -> (H-2,W-2,1). Non-integer inputs (e.g. the
    +-1 difference image of estimate.py:127) go through the float path: the kernel sees x/255 like the reference."""
    dev = _device(device)
    x0 = np.ascontiguousarray(np.asarray(x)[..., 0], dtype=np.float32)
    t = torch.from_numpy(x0 / np.float32(255.)).to(dev)[None]
    y = filter_predict(t, filter_name)[0]
    return y.cpu().numpy()[..., None]


def get_filter_estimator(filter_name: str, flatten: bool = False, device=None):
    """src/filters/evaluate.py:144-146."""
    if flatten:
        raise NotImplementedError("flatten=True hands the 8x1 OLS vector to a 2-D convolution in the reference "
                                  "(filters/evaluate.py:136-146), which is not a pixel predictor; the matrix form is "
                                  "get_filter_residuals")
    if filter_name not in NAMED_FILTERS_2D:
        raise KeyError(filter_name)
    return lambda x: infere_single(x, filter_name, device)
