"""ORACLE (test infrastructure): ctypes driver of oracle/unet_oracle.c - the reference's UNet.forward
(src/unet/model/unet.py:137-189) restated in plain C with double accumulation."""
import ctypes
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_lib = None
_fp = ctypes.POINTER(ctypes.c_float)


def build():
    subprocess.run(['make', '-C', str(_HERE)], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        so = _HERE / 'liboracle.so'
        if not so.exists():
            build()
        _lib = ctypes.CDLL(str(so))
    return _lib


def _p(a):
    return a.ctypes.data_as(_fp)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def conv3x3_reflect(x, w, b, relu=True):
    x, w, b = _f32(x), _f32(w), _f32(b)
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    y = np.empty((B, Cout, H, W), np.float32)
    lib().wso_conv3x3_reflect(_p(x), _p(w), _p(b), _p(y), B, Cin, Cout, H, W, int(relu))
    return y


def maxpool2(x):
    x = _f32(x)
    B, C, H, W = x.shape
    y = np.empty((B, C, H // 2, W // 2), np.float32)
    lib().wso_maxpool2(_p(x), _p(y), B, C, H, W)
    return y


def upconv2(x, w, b):
    x, w, b = _f32(x), _f32(w), _f32(b)
    B, Cin, H, W = x.shape
    Cout = w.shape[1]
    y = np.empty((B, Cout, 2 * H, 2 * W), np.float32)
    lib().wso_upconv2(_p(x), _p(w), _p(b), _p(y), B, Cin, Cout, H, W)
    return y


def head(x, w, b):
    x, w, b = _f32(x), _f32(w), _f32(b)
    B, C, H, W = x.shape
    y = np.empty((B, 1, H, W), np.float32)
    lib().wso_head(_p(x), _p(w.reshape(-1)), _p(b), _p(y), B, C, H, W)
    return y


def unet_forward(sd: dict, x: np.ndarray, nsteps: int, keep: bool = False):
    """sd: state_dict-like {name: ndarray} with the reference's keys; x (B,Cin,H,W) float32 in [0,1]."""
    acts, enc = {}, []
    h = _f32(x)
    for l in range(nsteps + 1):
        a = conv3x3_reflect(h, sd[f'e{l + 1}1.weight'], sd[f'e{l + 1}1.bias'])
        b = conv3x3_reflect(a, sd[f'e{l + 1}2.weight'], sd[f'e{l + 1}2.bias'])
        acts[f'e{l + 1}1'], acts[f'e{l + 1}2'] = a, b
        enc.append(b)
        if l < nsteps:
            h = maxpool2(b)
            acts[f'p{l + 1}'] = h
    h = enc[-1]
    for l in range(nsteps - 1, -1, -1):
        k = 4 - l
        u = upconv2(h, sd[f'upconv{k}.weight'], sd[f'upconv{k}.bias'])
        acts[f'u{k}'] = u
        a = conv3x3_reflect(np.concatenate([u, enc[l]], axis=1), sd[f'd{k}1.weight'], sd[f'd{k}1.bias'])
        h = conv3x3_reflect(a, sd[f'd{k}2.weight'], sd[f'd{k}2.bias'])
        acts[f'd{k}1'], acts[f'd{k}2'] = a, h
    y = head(h, sd['outconv.weight'], sd['outconv.bias'])
    return (y, acts) if keep else y


def ws_attack_c(img_u8, kind=0, xhat=None, xbias=None, weighted=0, clip=True, correct_bias=False):
    """C restatement of src/ws/estimate.py:83-128 (double precision, direct stencils). Returns (beta, l1, beta_raw)."""
    img = np.ascontiguousarray(img_u8, dtype=np.uint8)
    H, W = img.shape[:2]
    out = (ctypes.c_double * 3)()
    xh = _f32(xhat).reshape(-1) if xhat is not None else None
    xb = _f32(xbias).reshape(-1) if xbias is not None else None
    lib().wso_ws_attack(img.ctypes.data_as(ctypes.c_void_p), H, W, int(kind), _p(xh) if xh is not None else None,
                        _p(xb) if xb is not None else None, int(weighted), int(clip), int(correct_bias), out)
    return out[0], out[1], out[2]


def numpy_weights(nsteps: int, in_channels: int = 1, seed: int = 0, scale: float = 1.0) -> dict:
    """Deterministic, torch-version-independent weights with PyTorch-default-like magnitudes
    (uniform(+-1/sqrt(fan_in))). Keys/shapes = the reference's state_dict."""
    rng = np.random.default_rng(seed)
    sd = {}

    def conv(name, cout, cin, k):
        bound = scale / np.sqrt(cin * k * k)
        sd[f'{name}.weight'] = rng.uniform(-bound, bound, (cout, cin, k, k)).astype(np.float32)
        sd[f'{name}.bias'] = rng.uniform(-bound, bound, (cout,)).astype(np.float32)

    def convT(name, cin, cout):
        bound = scale / np.sqrt(cout * 4)
        sd[f'{name}.weight'] = rng.uniform(-bound, bound, (cin, cout, 2, 2)).astype(np.float32)
        sd[f'{name}.bias'] = rng.uniform(-bound, bound, (cout,)).astype(np.float32)

    ch = lambda l: 64 << l
    for l in range(nsteps + 1):
        conv(f'e{l + 1}1', ch(l), in_channels if l == 0 else ch(l - 1), 3)
        conv(f'e{l + 1}2', ch(l), ch(l), 3)
    for l in range(nsteps - 1, -1, -1):
        k = 4 - l
        convT(f'upconv{k}', ch(l + 1), ch(l))
        conv(f'd{k}1', ch(l), 2 * ch(l), 3)
        conv(f'd{k}2', ch(l), ch(l), 3)
    conv('outconv', 1, 64, 1)
    return sd
