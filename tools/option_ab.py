"""A/B of one wsu_set_option key on the per-layer times of the fp16x1 plan (CUDA events, min of reps).
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
res, outs = {}, {}
for v in (va, vb, va, vb):
    _native.check(lib.wsu_set_option(h, key.encode(), v))
    outs[v] = W.ws_estimate(imgs[:4], model, weighted=0, clip=False)
    lib.wsu_set_option(h, b'profile', 1)
    acc = res.get(v)
    for _ in range(reps):
        W.ws_estimate(imgs, model, weighted=0, clip=True, crop=1)
        torch.cuda.synchronize()
        buf = (ctypes.c_float * 64)()
        k = lib.wsu_profile_read(h, buf, 64)
        cur = [buf[i] for i in range(k)]
        acc = cur if acc is None else [min(a, c) for a, c in zip(acc, cur)]
    res[v] = acc
    names = [lib.wsu_profile_name(h, i).decode() for i in range(len(acc))]
    lib.wsu_set_option(h, b'profile', 0)
print(f'{key}: {va} vs {vb}; {n} images, {mode}; results equal: {torch.equal(outs[va], outs[vb])}')
for i, nm in enumerate(names):
    print(f'{nm:9s} {res[va][i]:8.3f} {res[vb][i]:8.3f}  {100 * (res[vb][i] / res[va][i] - 1):+6.1f} %')
print(f'total     {sum(res[va]):8.3f} {sum(res[vb]):8.3f}  {100 * (sum(res[vb]) / sum(res[va]) - 1):+6.1f} %')
