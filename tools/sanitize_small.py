"""Small shapes through every kernel family, meant to be run under compute-sanitizer (memcheck / racecheck)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import data as wdata

dev = torch.device('cuda', 0)
torch.manual_seed(0)
rng = np.random.default_rng(0)
for nsteps, (h, w) in ((2, (40, 72)), (2, (64, 96)), (1, (34, 50)), (0, (24, 40))):
    h, w = (h // (1 << nsteps)) * (1 << nsteps), (w // (1 << nsteps)) * (1 << nsteps)
    m = W.get_model(f'unet_{nsteps}', 1).to(dev)
    img = torch.from_numpy(rng.integers(0, 256, (3, 1, h, w), dtype=np.uint8)).to(dev)
    b, l1, y = W.ws_estimate(img, m, weighted=1, clip=True, return_l1=True, return_prediction=True)
    b0 = W.ws_estimate(img, m, weighted=0, clip=False)
    print('unet', nsteps, h, w, b.tolist(), b0.tolist())
    if nsteps in (1, 2):      # the reduced precision plans: fp16 maps, fp16 + e4m3 maps, resident weights, TMA stores, staged pooling
        for mode in ('fp16x1', 'fp16x1_f8'):
            m.set_precision(mode)
            bq = W.ws_estimate(img, m, weighted=1, clip=True, correct_bias=True)
            print('  ', mode, m.active_precision(dev), bq.tolist())
for h, w in ((3, 16), (5, 32), (66, 528), (35, 516), (130, 1040)):
    img = torch.from_numpy(rng.integers(0, 256, (9, 1, h, w), dtype=np.uint8)).to(dev)
    for name in ('KB', 'AVG'):
        for wt in (0, 1, -1):
            b = W.ws_estimate(img, name, weighted=wt, clip=False)
            b2, l1 = W.ws_estimate(img, name, weighted=wt, clip=False, return_l1=True)
    print('filters', h, w, b[:2].tolist())
torch.cuda.synchronize()
print('done')
