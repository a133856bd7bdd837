"""Per-layer device times of the UNet chain under each precision plan (CUDA events around every launch, min of `reps`).
Usage (on a B200): python tools/precision_profile.py [images=32] [reps=3]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import _native

GF = {'e11': 0.302, 'e12': 19.327, 'e21': 9.664, 'e22': 19.327, 'e31': 9.664, 'e32': 19.327, 'upconv3': 4.295,
      'd31': 38.655, 'd32': 19.327, 'upconv4': 4.295, 'd41': 38.655, 'd42': 19.361}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device('cuda', 0)
    torch.manual_seed(1234)
    model = W.get_model('unet_2', 1).to(dev)
    imgs = torch.randint(0, 256, (n, 1, 512, 512), dtype=torch.uint8, device=dev)
    lib = _native.load()
    model.set_micro_batch(n, dev)
    res, ref = {}, None
    for mode in ('bf16x3', 'fp16x1', 'fp16x1+pair2', 'fp16x1_f8+pair2'):
        model.set_precision(mode.split('+')[0])
        h = model.native_handle(dev)
        lib.wsu_set_option(h, b'cta_pair', 2 if mode.endswith('pair2') else 1)
        for kv in os.environ.get('WSU_OPTS', '').split(','):       # e.g. WSU_OPTS=tma_store=1 for A/B runs
            if '=' in kv:
                k, v = kv.split('=')
                _native.check(lib.wsu_set_option(h, k.encode(), int(v)), 'wsu_set_option')
        y = model(imgs[:4])
        if ref is None:
            ref = y
        err = ((y - ref).abs().max() * 255).item()
        lib.wsu_set_option(h, b'profile', 1)
        acc = None
        for _ in range(reps):
            W.ws_estimate(imgs, model, weighted=0, clip=True, crop=1)
            torch.cuda.synchronize()
            buf = (ctypes.c_float * 64)()
            k = lib.wsu_profile_read(h, buf, 64)
            cur = [buf[i] for i in range(k)]
            acc = cur if acc is None else [min(a, c) for a, c in zip(acc, cur)]
        names = [lib.wsu_profile_name(h, i).decode() for i in range(len(acc))]
        lib.wsu_set_option(h, b'profile', 0)
        info = ctypes.c_int64()
        lib.wsu_get_info(h, b'bytes_per_image', ctypes.byref(info))
        res[mode] = (dict(zip(names, acc)), err, info.value)
    names = list(res['bf16x3'][0])
    print(f'{n} images 512x512, ms per layer (min of {reps}); algorithmic TFLOP/s in brackets')
    print('layer    ' + ''.join(f'{m:>22s}' for m in res))
    for nm in names:
        print(f'{nm:9s}' + ''.join(f'{res[m][0][nm]:12.3f} [{GF[nm] * n / res[m][0][nm]:6.0f}]  ' for m in res))
    for m in res:
        tot = sum(res[m][0].values())
        print(f'{m}: total {tot:.3f} ms = {n / tot * 1e3:.0f} images/s; max|x_hat - three-term| = {res[m][1]:.2e} px; activation bytes/image (incl. 1/8 slack) {res[m][2] / 1e6:.0f} MB')


if __name__ == '__main__':
    main()
