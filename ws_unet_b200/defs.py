"""Host-side helpers either side of the hot path, mirroring the reference's `_defs` package for this path:
  * imread_u8 / imread_f32 / imread4_u8 / imread4_f32   - src/_defs/imread.py:8-27
  * get_processor (N x 9 neighbour matrix) / get_processor_2d (channel select) - src/_defs/filters.py:39-83
These are file decoding and index bookkeeping (numpy); the arithmetic on their outputs runs in libwsunet.
"""
from __future__ import annotations

import typing

import numpy as np

INBAYERS = ['00', '01', '10', '11']
# (row offset, column offset) of the columns of the neighbour matrix: the eight neighbours clockwise from the
# top-left corner, then the centre pixel (the regression target) - src/_defs/filters.py:55-65
NEIGHBOUR_ORDER = [(0, 0), (0, 1), (0, 2), (1, 2), (2, 2), (2, 1), (2, 0), (1, 0), (1, 1)]


def imread_u8(fname) -> np.ndarray:
    """PIL decode, always (H,W,C) - src/_defs/imread.py:8-12."""
    from PIL import Image
    x = np.asarray(Image.open(fname))
    return x[..., None] if x.ndim == 2 else x


def imread_f32(fname) -> np.ndarray:
    return imread_u8(fname).astype('float32')


def imread4_u8(fname) -> np.ndarray:
    """(H,W,4) uint8 = R, G, B and the cv2 luma Y - src/_defs/imread.py:19-23."""
    import cv2
    bgr = cv2.imread(str(fname))
    if bgr is None:
        raise IOError(f'cannot read {fname}')
    y = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    return np.dstack([bgr[..., 2], bgr[..., 1], bgr[..., 0], y])


def imread4_f32(fname) -> np.ndarray:
    return imread4_u8(fname).astype('float32')


def get_processor(channels: typing.Sequence[int], inbayer: str = None) -> typing.Callable:
    """src/_defs/filters.py:39-69: (H,W,C) image -> (N, 9) matrix, one row per interior pixel (8 neighbours of
    channels[0] clockwise from the top-left, then the pixel itself). `inbayer` 'ab' keeps one site of the 2x2 Bayer
    lattice: stride 2, after dropping the first/last row (a == '0') and column (b == '0')."""
    step = 2 if inbayer else 1
    trim_r = bool(inbayer) and inbayer[0] == '0'
    trim_c = bool(inbayer) and inbayer[1] == '0'
    ch = channels[0]

    def process_gray(x: np.ndarray) -> np.ndarray:
        p = np.asarray(x)[..., ch]
        if trim_r:
            p = p[1:-1]
        if trim_c:
            p = p[:, 1:-1]
        h, w = p.shape
        cols = [p[dr:h - 2 + dr:step, dc:w - 2 + dc:step].reshape(-1) for dr, dc in NEIGHBOUR_ORDER]
        return np.stack(cols, axis=-1)

    return process_gray


def get_processor_2d(channels: typing.Sequence[int]) -> typing.Callable:
    """src/_defs/filters.py:72-83: select `channels`, cast to float32, keep the 2-D layout."""
    idx = list(channels)

    def process_gray(x: np.ndarray) -> np.ndarray:
        return np.asarray(x)[:, :, idx].astype('float32')

    return process_gray
