"""Per-call latency of the drop-in per-image entry points (how the reference's scripts call the path: one file at a
time) and throughput of small batches."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import data as wdata

dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = W.get_model('unet_2', 1).to(dev)
imgs = wdata.synthetic_stego_fast(64, 0.4, 512, 512, dev, unique=32)
for b in (1, 2, 4, 8, 16, 32):
    x = imgs[:b].contiguous()
    for _ in range(3):
        W.ws_estimate(x, model)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 20
    for _ in range(reps):
        W.ws_estimate(x, model)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'device batch {b:2d}: {ms:7.3f} ms/call -> {b / ms * 1e3:7.1f} img/s')
# reference-style per-file call: numpy image in, dict out (predict_unet), wall clock
x4 = np.repeat(imgs[0, 0].cpu().numpy()[..., None], 4, axis=2).astype('float32')
from ws_unet_b200.unet import predict_unet
for _ in range(3):
    predict_unet('mem', model, imread=lambda f: x4)
t0 = time.perf_counter()
n = 50
for _ in range(n):
    r = predict_unet('mem', model, imread=lambda f: x4)
dt = (time.perf_counter() - t0) / n
print(f'predict_unet(fname, model) per call: {dt * 1e3:.3f} ms -> {1 / dt:.1f} img/s  beta_hat={r["beta_hat"]:.6f}')
