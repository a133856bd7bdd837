"""Throughput of every depth the reference supports (unet_0..unet_4, src/unet/model/unet.py:99-132) and of the estimator
variants on unet_2, 128 images of 512x512 (CUDA events, 3 passes). Usage (on a B200): python tools/depth_variants_bench.py"""
import sys

import torch

sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import data as wdata

GF = {0: 19.66, 1: 110.93, 2: 202.199, 3: 293.47, 4: 384.74}   # algorithmic GFLOP per 512x512 image
dev = torch.device('cuda', 0)
imgs = wdata.synthetic_stego_fast(128, 0.4, 512, 512, dev, unique=32)


def rate(model, **kw):
    for _ in range(2):
        W.ws_estimate(imgs, model, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        W.ws_estimate(imgs, model, **kw)
    e1.record()
    torch.cuda.synchronize()
    return 128 / (e0.elapsed_time(e1) / 3) * 1e3


for n in range(5):
    torch.manual_seed(0)
    m = W.get_model(f'unet_{n}', 1).to(dev)
    r3 = rate(m, weighted=0)
    rep = m.calibrate_precision(imgs[:8])
    r = rate(m, weighted=0)
    print(f'unet_{n}: three-term {r3:8.1f} img/s ({GF[n] * r3 / 1e3:6.1f} TFLOP/s algorithmic) | calibrated plan {rep["chosen"]:7s} '
          f'{r:8.1f} img/s ({GF[n] * r / 1e3:6.1f} TFLOP/s)  errors {rep["max_abs_px"]}')
    del m
torch.manual_seed(0)
m = W.get_model('unet_2', 1).to(dev)
m.calibrate_precision(imgs[:8])
for kw in (dict(weighted=0), dict(weighted=1), dict(weighted=0, return_l1=True), dict(weighted=0, return_prediction=True),
           dict(weighted=1, correct_bias=True)):
    print('unet_2', kw, f'{rate(m, **kw):8.1f} img/s')
