"""End-to-end (host buffer) throughput of the fused chain vs micro-batch size."""
import sys, time, torch
sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import data as wdata
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = W.get_model('unet_2', 1).to(dev)
host = wdata.synthetic_stego_fast(256, 0.4, 512, 512, dev, unique=32).cpu().pin_memory()
for mb in (16, 32, 64, 32, 64):
    model.set_micro_batch(mb, dev)
    for _ in range(2): W.ws_estimate_host(host, model)
    ts = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        W.ws_estimate_host(host, model)
        ts.append(time.perf_counter() - t0)
    print(f'micro_batch={mb}: ' + ' '.join(f'{256 / t:.0f}' for t in ts) + ' img/s per call')
