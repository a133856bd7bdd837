"""First-light GPU check: per-layer parity of the tcgen05 chain against the PyTorch fp32 reference, filter/WS kernels,
and a rough timing. Run on a B200 via gpurun; prints a compact report."""
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import _native
from oracle.torch_port import reference_forward


def debug_layer(model, name, dev, halo=0):
    lib = _native.load()
    dims = (ctypes.c_int64 * 4)()
    cap = 1 << 28
    buf = torch.empty(cap, dtype=torch.float32, device=dev)
    _native.check(lib.wsu_debug_layer(model._handle, name.encode(), ctypes.c_void_p(buf.data_ptr()), cap, halo, dims,
                                      _native.stream_ptr(dev)))
    n = dims[0] * dims[1] * dims[2] * dims[3]
    return buf[:n].view(dims[0], dims[1], dims[2], dims[3]).clone()


def layer_report(nsteps, B, H, Wd, seed=0):
    dev = torch.device('cuda', 0)
    torch.manual_seed(seed)
    model = W.get_model(f'unet_{nsteps}', 1).to(dev)
    x = torch.rand(B, 1, H, Wd, device=dev)
    _native.load().wsu_set_option(model.native_handle(dev), b'fuse_e11', 0)   # keep e11 inspectable
    y = model(x)
    torch.cuda.synchronize()
    yref, acts = reference_forward(model, x, keep=True)
    print(f'--- unet_{nsteps} B={B} {H}x{Wd}: out max|d|={(y - yref).abs().max().item():.3e} (x255: {(y - yref).abs().max().item() * 255:.3e} px)')
    for name, ref in acts.items():
        try:
            got = debug_layer(model, name, dev)
        except Exception as e:  # head input is not materialised
            continue
        d = (got - ref).abs()
        print(f'   {name:5s} shape={tuple(ref.shape)} max|d|={d.max().item():.3e} ref max={ref.abs().max().item():.3e}'
              f' bad={(d > 1e-3 * (1 + ref.abs())).float().mean().item():.4f}')
        if name in ('e11', 'e12'):
            gh = debug_layer(model, name, dev, halo=1)
            rh = torch.nn.functional.pad(ref, (1, 1, 1, 1), mode='reflect')
            print(f'      halo max|d|={(gh - rh).abs().max().item():.3e}')
    return model


def filters_report():
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (3, 1, 64, 96), generator=g, dtype=torch.uint8).to(dev)
    xf = img.float()
    kb = torch.tensor([[-1, 2, -1], [2, 0, 2], [-1, 2, -1]], dtype=torch.float32, device=dev)[None, None] / 4
    ref = torch.nn.functional.conv2d(xf, kb)[:, 0]
    got = W.filters.filter_predict(img, 'KB')
    print('KB predict max|d| =', (got - ref).abs().max().item())
    x1 = xf[:, 0, 1:-1, 1:-1]
    xbar = (img ^ 1).float()[:, 0, 1:-1, 1:-1]
    beta_ref = ((x1 - xbar) * (x1 - ref)).mean(dim=(1, 2)).clamp_min(0)
    beta = W.ws_estimate(img, 'KB', weighted=0)
    print('KB beta (w=0):', beta.tolist(), 'ref', beta_ref.tolist())
    avg = torch.ones(1, 1, 3, 3, device=dev) / 8
    avg[0, 0, 1, 1] = 0
    mu = torch.nn.functional.conv2d(xf.double(), avg.double())[:, 0]
    mu2 = torch.nn.functional.conv2d(xf.double() ** 2, avg.double())[:, 0]
    w = 1 / (5 + (mu2 - mu ** 2))
    w = w / w.sum(dim=(1, 2), keepdim=True)
    beta_ref = (w * (x1 - xbar) * (x1 - ref)).sum(dim=(1, 2)).clamp_min(0)
    beta = W.ws_estimate(img, 'KB', weighted=1)
    print('KB beta (w=1):', beta.tolist(), 'ref', beta_ref.tolist())


def fused_report(model):
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(2)
    img = torch.randint(0, 256, (4, 1, 64, 64), generator=g, dtype=torch.uint8).to(dev)
    beta, l1, yhat = W.ws_estimate(img, model, weighted=0, clip=False, return_l1=True, return_prediction=True)
    yref = reference_forward(model, img.float() / 255.)
    xh = yref * 255
    x = img.float()
    xbar = (img ^ 1).float()
    c = (slice(None), 0, slice(1, -1), slice(1, -1))
    beta_ref = ((x[c] - xbar[c]) * (x[c] - xh[c])).mean(dim=(1, 2))
    l1_ref = (x[c] - xh[c]).abs().mean(dim=(1, 2))
    print('fused yhat max|d| px =', ((yhat - yref).abs().max() * 255).item())
    print('fused beta', beta.tolist(), 'ref', beta_ref.tolist())
    print('fused l1  ', l1.tolist(), 'ref', l1_ref.tolist())


def timing(B=32):
    dev = torch.device('cuda', 0)
    torch.manual_seed(0)
    model = W.get_model('unet_2', 1).to(dev)
    img = torch.randint(0, 256, (B, 1, 512, 512), dtype=torch.uint8, device=dev)
    for _ in range(2):
        W.ws_estimate(img, model)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for _ in range(n):
        W.ws_estimate(img, model)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f'timing: B={B} 512x512 {ms:.2f} ms/batch -> {B / ms * 1e3:.1f} img/s, {202.199e9 * B / ms * 1e3 / 1e12:.1f} TFLOP/s algorithmic')
    y = model(img[:4].float() / 255)
    yref = reference_forward(model, img[:4].float() / 255)
    print('512x512 out max|d| px =', ((y - yref).abs().max() * 255).item())


if __name__ == '__main__':
    which = sys.argv[1:] or ['filters', 'small', 'fused', 'timing']
    print(torch.cuda.get_device_name(0))
    if 'filters' in which:
        filters_report()
    m = None
    if 'small' in which:
        layer_report(0, 2, 32, 48)
        m = layer_report(2, 2, 64, 64)
    if 'fused' in which:
        fused_report(m or layer_report(2, 2, 64, 64))
    if 'mid' in which:
        layer_report(2, 2, 256, 256)
    if 'timing' in which:
        timing()
