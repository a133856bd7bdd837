"""Throughput of the fused UNet->WS chain vs micro-batch size (images per pass through the layer chain)."""
import sys
import torch
sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import data as wdata

dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = W.get_model('unet_2', 1).to(dev)
imgs = wdata.synthetic_stego_fast(256, 0.4, 512, 512, dev, unique=32)
for mb in [int(a) for a in sys.argv[1:]] or [2, 4, 8, 16, 32, 64]:
    model.set_micro_batch(mb, dev)
    for _ in range(2):
        W.ws_estimate(imgs, model)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        W.ws_estimate(imgs, model)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print(f'micro_batch={mb:3d}: {ms:8.2f} ms / 256 images -> {256 / ms * 1e3:7.1f} img/s')
