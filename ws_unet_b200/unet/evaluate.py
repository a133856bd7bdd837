"""Per-image UNet inference helpers - drop-in for src/unet/evaluate.py:31-52,109-139,162-188."""
from __future__ import annotations

import json
import pathlib
import typing

import numpy as np
import torch

from .model import get_model
from .. import ws as _ws


def _cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("ws_unet_b200 needs a CUDA device (no CPU fallback)")
    if device is None or torch.device(device).type != 'cuda':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device(device)


def _center_crop_512(x: torch.Tensor) -> torch.Tensor:
    """transforms.CenterCrop(512) of get_timm_transform (src/unet/data/loader.py:43-44): crop if larger,
    zero-pad if smaller (torchvision semantics, same rounding)."""
    h, w = x.shape[-2:]
    if h < 512 or w < 512:
        pl = (512 - w) // 2 if w < 512 else 0
        pt = (512 - h) // 2 if h < 512 else 0
        pr = (512 - w + 1) // 2 if w < 512 else 0
        pb = (512 - h + 1) // 2 if h < 512 else 0
        x = torch.nn.functional.pad(x, (pl, pr, pt, pb))
        h, w = x.shape[-2:]
    top, left = int(round((h - 512) / 2.)), int(round((w - 512) / 2.))
    return x[..., top:top + 512, left:left + 512]


def infere_single(x: np.ndarray, model: typing.Callable, device=None) -> np.ndarray:
    """src/unet/evaluate.py:31-52: (H,W,1) float32 pixels -> /255 -> ToTensor, CenterCrop(512) -> model ->
    crop 1 px -> *255 -> (510,510,1) float32."""
    dev = _cuda(device)
    x_ = torch.from_numpy(np.ascontiguousarray((np.asarray(x, dtype=np.float32) / 255.).transpose(2, 0, 1)))
    x_ = _center_crop_512(x_)[None, :1].contiguous().to(dev)
    y_ = model(x_)
    y = y_.detach().cpu().numpy()[0, 0, 1:-1, 1:-1] * 255.
    return y[..., None]


def predict_unet(fname: str, model: torch.nn.Module, *, device=None, imread: typing.Callable = None, **kw):
    """src/unet/evaluate.py:109-139: beta_hat = mean((x - x_bar)(x - x_hat)) (unweighted, unclipped) and
    l1 = mean|x - x_hat| over the 510x510 interior, computed by the fused UNet->WS kernel chain."""
    dev = _cuda(device)
    if imread is None:                      # reference default: _defs.imread4_f32 -> (H,W,4) with luma in channel 3
        from ..dataset import imread_gray_u8 as imread
    x = np.asarray(imread(fname))
    if x.ndim == 2:
        x = x[..., None]
    x = x[..., 3:] if x.shape[-1] >= 4 else x[..., -1:]
    img = torch.from_numpy(np.ascontiguousarray(x[..., 0]).astype(np.uint8)).to(dev)[None, None]
    img = _center_crop_512(img).contiguous()
    beta, l1 = _ws.ws_estimate(img, model, weighted=0, clip=False, crop=1, return_l1=True)
    return {**kw, 'beta_hat': np.float32(beta.item()), 'l1': np.float32(l1.item())}


def get_pretrained(model_path, channels, *, model_name: str = None, device=None):
    """src/unet/evaluate.py:162-188."""
    model_path = pathlib.Path(model_path)
    with open(model_path / model_name / 'config.json') as f:
        config = json.load(f)
    dev = _cuda(device)
    model = get_model(config['network'], in_channels=1, out_channels=1, channel=[0], drop_rate=0.).to(dev)
    checkpoint = torch.load(model_path / model_name / 'model' / 'best_model.pt.tar', map_location=dev, weights_only=True)
    model.load_state_dict(checkpoint['state_dict'])
    return model
