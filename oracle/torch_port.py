"""ORACLE (test infrastructure): CPU port of the reference's per-image UNet -> WS path built from the SAME library
calls the reference makes (torch ATen CPU conv2d / conv_transpose2d / max_pool2d, numpy reductions):
  reference_forward  = UNet.forward                       src/unet/model/unet.py:137-189
  infere_single      = /255 -> model -> crop 1 px -> *255 src/unet/evaluate.py:31-52 (batch 1, autograd on)
  predict_unet       = beta_hat / l1                      src/unet/evaluate.py:125-132
It is what bench.py times as the CPU baseline (kind "port": /root/reference does not exist on the GPU box) and what
tests use as the fp32 reference of the floating-point kernels. The product path never imports it."""
import numpy as np
import torch
import torch.nn.functional as F


def _conv(m, x):
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), m.weight, m.bias)


def reference_forward(model, x, keep=False):
    """x: (B,C,H,W) float32 in [0,1] on any device. Returns sigmoid output and (optionally) every feature map
    keyed like wsu_debug_layer names."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n = model.nsteps
    acts = {}
    enc = []
    h = x
    for l in range(n + 1):
        a = F.relu(_conv(getattr(model, f'e{l + 1}1'), h))
        b = F.relu(_conv(getattr(model, f'e{l + 1}2'), a))
        acts[f'e{l + 1}1'], acts[f'e{l + 1}2'] = a, b
        enc.append(b)
        if l < n:
            h = F.max_pool2d(b, 2, 2)
            acts[f'p{l + 1}'] = h
    h = enc[-1]
    for l in range(n - 1, -1, -1):
        k = 4 - l
        up = getattr(model, f'upconv{k}')
        u = F.conv_transpose2d(h, up.weight, up.bias, stride=2)
        acts[f'u{k}'] = u
        a = F.relu(_conv(getattr(model, f'd{k}1'), torch.cat([u, enc[l]], dim=1)))
        h = F.relu(_conv(getattr(model, f'd{k}2'), a))
        acts[f'd{k}1'], acts[f'd{k}2'] = a, h
    y = torch.sigmoid(F.conv2d(h, model.outconv.weight, model.outconv.bias))
    return (y, acts) if keep else y


def infere_single(x_hw1: "np.ndarray", model, no_grad: bool = False) -> "np.ndarray":
    """src/unet/evaluate.py:31-52 for an already 512x512 (or any /4-divisible) image: batch 1, CPU, autograd graph built
    and discarded exactly like the reference unless no_grad is set."""
    x_ = torch.from_numpy(np.ascontiguousarray((x_hw1 / 255.).transpose(2, 0, 1)))[None]
    if no_grad:
        with torch.no_grad():
            y_ = reference_forward(model, x_)
    else:
        y_ = reference_forward(model, x_)
    y = y_.detach().numpy()[0, 0, 1:-1, 1:-1] * 255.
    return y[..., None]


def predict_unet(x_u8_hw: "np.ndarray", model, no_grad: bool = False):
    """src/unet/evaluate.py:118-132 on an in-memory uint8 image."""
    x = x_u8_hw.astype('float32')[..., None]
    x_hat = infere_single(x, model, no_grad=no_grad)
    x = x[1:-1, 1:-1]
    x_bar = (x.astype('uint8') ^ 1).astype('float32')
    beta_hat = np.mean((x - x_bar) * (x - x_hat))
    l1_hat = np.mean(np.abs(x - x_hat))
    return beta_hat, l1_hat
