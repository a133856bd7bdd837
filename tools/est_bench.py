"""Throughput of the linear-filter WS estimator kernels on resident uint8 images (BASELINE configs[1])."""
import sys
import torch
sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import data as wdata

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
imgs = wdata.synthetic_stego_fast(n, 0.4, 512, 512, torch.device('cuda', 0), unique=32)
ref = None
for name in ['KB', 'AVG']:
    for kw in [dict(weighted=0), dict(weighted=0, return_l1=True), dict(weighted=1)]:
        for _ in range(10):
            out = W.ws_estimate(imgs, name, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            W.ws_estimate(imgs, name, **kw)
        e1.record()
        torch.cuda.synchronize()
        s = e0.elapsed_time(e1) / 50 / 1e3
        b = out[0] if isinstance(out, tuple) else out
        print(name, kw, f'{n / s / 1e6:.2f} M img/s  {262148 * n / s / 1e9:.0f} GB/s  {262148 * n / s / 1e9 / 6544:.3f} of HBM peak  beta[:3]={b[:3].tolist()}')
