"""A/B of one wsu_set_option key on the per-layer times (CUDA events, median of 2*reps passes, variants interleaved).
Usage (on a B200): python tools/option_ab.py <key> <value_a> <value_b> [images=32] [reps=3] [precision=fp16x1]"""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import _native

key, va, vb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 32
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
mode = sys.argv[6] if len(sys.argv) > 6 else 'fp16x1'
dev = torch.device('cuda', 0)
torch.manual_seed(1234)
model = W.get_model('unet_2', 1).to(dev).set_precision(mode)
imgs = torch.randint(0, 256, (n, 1, 512, 512), dtype=torch.uint8, device=dev)
lib = _native.load()
model.set_micro_batch(n, dev)
h = model.native_handle(dev)
res, outs = {va: [], vb: []}, {}
for v in (va, vb):
    _native.check(lib.wsu_set_option(h, key.encode(), v))
    outs[v] = W.ws_estimate(imgs[:4], model, weighted=0, clip=False)
    W.ws_estimate(imgs, model, weighted=0, clip=True, crop=1)          # warm-up of this variant's kernels
lib.wsu_set_option(h, b'profile', 1)
for rep in range(2 * reps):                 # A B B A A B B A ...: clock drift inside the process hits both variants alike
    for v in ((va, vb) if rep % 2 == 0 else (vb, va)):
        _native.check(lib.wsu_set_option(h, key.encode(), v))
        W.ws_estimate(imgs, model, weighted=0, clip=True, crop=1)
        torch.cuda.synchronize()
        buf = (ctypes.c_float * 64)()
        k = lib.wsu_profile_read(h, buf, 64)
        res[v].append([buf[i] for i in range(k)])
names = [lib.wsu_profile_name(h, i).decode() for i in range(len(res[va][0]))]
lib.wsu_set_option(h, b'profile', 0)
med = {v: [sorted(r[i] for r in res[v])[len(res[v]) // 2] for i in range(len(names))] for v in (va, vb)}
print(f'{key}: {va} vs {vb}; {n} images, {mode}, median of {2 * reps} interleaved passes; results equal: {torch.equal(outs[va], outs[vb])}')
for i, nm in enumerate(names):
    print(f'{nm:9s} {med[va][i]:8.3f} {med[vb][i]:8.3f}  {100 * (med[vb][i] / med[va][i] - 1):+6.1f} %')
print(f'total     {sum(med[va]):8.3f} {sum(med[vb]):8.3f}  {100 * (sum(med[vb]) / sum(med[va]) - 1):+6.1f} %')
