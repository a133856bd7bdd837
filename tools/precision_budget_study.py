"""CPU study (not product code): per-layer precision budget of the tensor-core layers.

The shipped kernels compute every layer as a_hi*w_hi + a_lo*w_hi + a_hi*w_lo on bf16 pairs (3 MMAs per algorithmic MAC,
4.6e-5 px against the 1e-3 px bar). This script measures, with explicit roundings in float64 torch ops, what happens when
SOME layers read their input activations as ONE fp16 value (11 significant bits) against fp16 (hi, lo) weights - two
products, and a single 2-byte activation plane in HBM instead of two:
  a16   : act = fp16(a);             w = fp16 hi + fp16 lo;   y = a*w_hi + a*w_lo
  w16   : act = bf16 hi + bf16 lo;   w = fp16(w);             y = a_hi*w + a_lo*w
  split : the shipped three-term scheme
Usage: python tools/precision_budget_study.py <size> <n_images> <seed> <layers,comma> [mode=a16]
       layers e.g. d31,d41   (upconv outputs feeding an a16 layer are rounded where they are consumed)"""
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from oracle import unet_oracle as uo
from ws_unet_b200 import data as wdata

torch.set_num_threads(8)
D = torch.float64


def bf16(x):
    return x.to(torch.float32).to(torch.bfloat16).to(D)


def fp16(x):
    return x.to(torch.float32).to(torch.float16).to(D)


def f32(x):
    return x.to(torch.float32).to(D)


def split_bf16(x):
    x32 = f32(x)
    hi = bf16(x32)
    return hi, bf16(x32 - hi)


def split_fp16(x):
    x32 = f32(x)
    hi = fp16(x32)
    return hi, fp16(x32 - hi)


def conv3(x, w, b, mode):
    xp = F.pad(x, (1, 1, 1, 1), mode='reflect')
    if mode == 'exact':
        return F.conv2d(xp, w, b)
    if mode == 'split':
        ah, al = split_bf16(xp)
        wh, wl = split_bf16(w)
        y = F.conv2d(ah, wh) + F.conv2d(al, wh) + F.conv2d(ah, wl)
    elif mode == 'a16':
        a = fp16(xp)
        wh, wl = split_fp16(w)
        y = F.conv2d(a, wh) + F.conv2d(a, wl)
    elif mode == 'w16':
        ah, al = split_bf16(xp)
        wq = fp16(w)
        y = F.conv2d(ah, wq) + F.conv2d(al, wq)
    elif mode == 'a16w16':
        y = F.conv2d(fp16(xp), fp16(w))
    else:
        raise ValueError(mode)
    return f32(f32(y) + b.view(1, -1, 1, 1))


def upconv(x, w, b, mode):
    if mode == 'exact':
        return F.conv_transpose2d(x, w, b, stride=2)
    if mode == 'a16':
        a = fp16(x)
        wh, wl = split_fp16(w)
        y = F.conv_transpose2d(a, wh, stride=2) + F.conv_transpose2d(a, wl, stride=2)
    elif mode == 'a16w16':
        y = F.conv_transpose2d(fp16(x), fp16(w), stride=2)
    else:
        ah, al = split_bf16(x)
        wh, wl = split_bf16(w)
        y = F.conv_transpose2d(ah, wh, stride=2) + F.conv_transpose2d(al, wh, stride=2) + F.conv_transpose2d(ah, wl, stride=2)
    return f32(f32(y) + b.view(1, -1, 1, 1))


def conv3_cat(up, skip, w, b, mode_up, mode_skip):
    """3x3 conv over cat([up, skip]) with a precision mode per concat source (bias added once)"""
    c = up.shape[1]
    z = torch.zeros_like(b)
    return f32(conv3(up, w[:, :c], z, mode_up) + conv3(skip, w[:, c:], z, mode_skip) + b.view(1, -1, 1, 1))


def forward(sd, x, modes):
    """modes: dict layer -> mode; missing layers use modes['*']"""
    m = lambda n: modes.get(n, modes['*'])
    g = lambda n: (sd[n + '.weight'], sd[n + '.bias'])
    c3 = lambda n, t: F.relu(conv3(t, *g(n), m(n)))
    e11 = F.relu(conv3(x, *g('e11'), 'exact' if modes['*'] == 'exact' else 'split'))
    e12 = c3('e12', e11)
    e21 = c3('e21', F.max_pool2d(e12, 2))
    e22 = c3('e22', e21)
    e31 = c3('e31', F.max_pool2d(e22, 2))
    e32 = c3('e32', e31)
    d31 = c3('d31', torch.cat([upconv(e32, *g('upconv3'), m('upconv3')), e22], 1))
    d32 = c3('d32', d31)
    if 'd41.up' in modes:
        d41 = F.relu(conv3_cat(upconv(d32, *g('upconv4'), m('upconv4')), e12, *g('d41'), modes['d41.up'], m('d41')))
    else:
        d41 = c3('d41', torch.cat([upconv(d32, *g('upconv4'), m('upconv4')), e12], 1))
    d42 = c3('d42', d41)
    z = F.conv2d(d42, *g('outconv'))
    return torch.sigmoid(z) * 255.


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    nimg = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 102
    layers = sys.argv[4].split(',') if len(sys.argv) > 4 and sys.argv[4] else []
    mode = sys.argv[5] if len(sys.argv) > 5 else 'a16'
    sd = {k: torch.from_numpy(v).to(D) for k, v in uo.numpy_weights(2, seed=seed).items()}
    if len(sys.argv) > 6 and sys.argv[6] == 'torchinit':
        import ws_unet_b200  # noqa: F401  (only for the layer list)
        torch.manual_seed(seed)
        from oracle import torch_port
        sd = {k: v.to(D) for k, v in torch_port.reference_like_state_dict(2).items()} if hasattr(torch_port, 'reference_like_state_dict') else sd
    imgs = torch.stack([wdata.embed_lsbr(wdata.synthetic_cover(1000 * seed + i, size, size), 0.4, i) for i in range(nimg)])[:, None]
    x = imgs.to(D) / 255.
    worst = {}
    for i in range(nimg):
        xi = x[i:i + 1]
        ref = forward(sd, xi, {'*': 'exact'})
        for name, md in (('split', {'*': 'split'}), (mode + ':' + ','.join(layers), dict({'*': 'split'}, **{l: mode for l in layers}))):
            err = (forward(sd, xi, md) - ref).abs().max().item()
            worst[name] = max(worst.get(name, 0.), err)
    for k, v in worst.items():
        print(f'{k:40s} max |x_hat - x_hat_fp64| = {v:.3e} px over {nimg} images {size}x{size}, seed {seed}', flush=True)


if __name__ == '__main__':
    main()
