import sys, time, torch
sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import data as wdata
dev = torch.device('cuda', 0)
n = 10000
imgs = wdata.synthetic_stego_fast(n, 0.4, 512, 512, dev, unique=32)
def t(reps, label, x):
    for _ in range(3): W.ws_estimate(x, 'KB', weighted=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): W.ws_estimate(x, 'KB', weighted=0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'{label}: reps={reps} {ms:.3f} ms/call  {262148*n/ms/1e6:.0f} GB/s', flush=True)
for reps in (5, 20, 100, 500, 5, 20):
    t(reps, 'unique32', imgs)
# bench-like content: 256 distinct images repeated
base = torch.stack([wdata.embed_lsbr(wdata.synthetic_cover(i), 0.4, i) for i in range(64)])[:, None].to(dev)
rep = base.repeat(157, 1, 1, 1)[:n].contiguous()
for reps in (5, 20, 100):
    t(reps, 'covers64', rep)
rnd = torch.randint(0, 256, (n, 1, 512, 512), dtype=torch.uint8, device=dev)
for reps in (5, 20, 100):
    t(reps, 'random', rnd)
zero = torch.zeros((n, 1, 512, 512), dtype=torch.uint8, device=dev)
for reps in (5, 20, 100):
    t(reps, 'zeros', zero)
