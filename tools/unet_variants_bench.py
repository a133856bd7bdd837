import sys, torch
sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import data as wdata
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = W.get_model('unet_2', 1).to(dev)
imgs = wdata.synthetic_stego_fast(128, 0.4, 512, 512, dev, unique=32)
for kw in (dict(weighted=0), dict(weighted=1), dict(weighted=0, return_l1=True), dict(weighted=0, return_prediction=True)):
    for _ in range(2): W.ws_estimate(imgs, model, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): W.ws_estimate(imgs, model, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(kw, f'{ms:.2f} ms / 128 images -> {128 / ms * 1e3:.1f} img/s')
