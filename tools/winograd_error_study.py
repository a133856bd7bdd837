"""CPU study (not product code): would Winograd F(2x2, 3x3) with split-bf16 operands stay inside the 1e-3 px bar?

Emulates, in float64 torch ops with explicit roundings, the arithmetic a tensor-core implementation would perform:
  direct   : activations and weights as hi + lo bf16 pairs (what the shipped kernels do), products a_hi*w_hi + a_lo*w_hi +
             a_hi*w_lo, fp32 accumulation emulated as exact sums rounded once to fp32 per output
  winograd : per 3x3 layer V = B^T d B from the (hi + lo) activations in fp32, V and U = G g G^T each split into hi + lo bf16,
             the same three-term products summed over input channels, Y = A^T M A in fp32
and compares the sigmoid output (x255 = pixels) with the plain float64 forward on a random-init unet_2.
Usage: python tools/winograd_error_study.py [size=128] [seed=102]"""
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from oracle import unet_oracle as uo

torch.set_num_threads(8)
D = torch.float64
Bt = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=D)          # B^T (4x4)
G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=D)                      # G   (4x3)
At = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=D)                                          # A^T (2x4)
# F(4x4, 3x3) (Lavin & Gray): 36 multiplies per 16 outputs
Bt4 = torch.tensor([[4, 0, -5, 0, 1, 0], [0, -4, -4, 1, 1, 0], [0, 4, -4, -1, 1, 0], [0, -2, -1, 2, 1, 0],
                    [0, 2, -1, -2, 1, 0], [0, 4, 0, -5, 0, 1]], dtype=D)
G4 = torch.tensor([[1 / 4, 0, 0], [-1 / 6, -1 / 6, -1 / 6], [-1 / 6, 1 / 6, -1 / 6], [1 / 24, 1 / 12, 1 / 6],
                   [1 / 24, -1 / 12, 1 / 6], [0, 0, 1]], dtype=D)
At4 = torch.tensor([[1, 1, 1, 1, 1, 0], [0, 1, -1, 2, -2, 0], [0, 1, 1, 4, 4, 0], [0, 1, -1, 8, -8, 1]], dtype=D)


def bf16(x):
    return x.to(torch.float32).to(torch.bfloat16).to(D)


def split(x):
    x32 = x.to(torch.float32).to(D)
    hi = bf16(x32)
    lo = bf16(x32 - hi)
    return hi, lo


def f32(x):
    return x.to(torch.float32).to(D)


def conv_direct(x, w, b, mode):
    xp = F.pad(x, (1, 1, 1, 1), mode='reflect')
    if mode == 'exact':
        return F.conv2d(xp, w, b)
    ah, al = split(xp)
    wh, wl = split(w)
    y = F.conv2d(ah, wh) + F.conv2d(al, wh) + F.conv2d(ah, wl)
    return f32(f32(y) + b.view(1, -1, 1, 1))


def conv_winograd(x, w, b, m=2):
    Bn, C, H, W = x.shape
    bt, g, at = (Bt, G, At) if m == 2 else (Bt4, G4, At4)
    xp = F.pad(x, (1, 1, 1, 1), mode='reflect')
    xh, xl = split(xp)
    xq = xh + xl                                                       # what the kernel reads back from HBM
    tiles = xq.unfold(2, m + 2, m).unfold(3, m + 2, m)                 # (B, C, H/m, W/m, m+2, m+2)
    V = f32(torch.einsum('ij,bcyxjk,lk->bcyxil', bt, tiles, bt))       # input transform in fp32
    U = torch.einsum('ij,ocjk,lk->ocil', g, w, g)                      # weight transform on the host (float64), then split
    Vh, Vl = split(V)
    Uh, Ul = split(U)
    M = (torch.einsum('bcyxil,ocil->boyxil', Vh, Uh) + torch.einsum('bcyxil,ocil->boyxil', Vl, Uh)
         + torch.einsum('bcyxil,ocil->boyxil', Vh, Ul))
    M = f32(M)                                                         # fp32 accumulators
    Y = f32(torch.einsum('ij,boyxjk,lk->boyxil', at, M, at))           # (B, O, H/m, W/m, m, m)
    y = Y.permute(0, 1, 2, 4, 3, 5).reshape(Bn, -1, H, W)
    return f32(y + b.view(1, -1, 1, 1))


def upconv(x, w, b, mode):
    if mode == 'exact':
        return F.conv_transpose2d(x, w, b, stride=2)
    ah, al = split(x)
    wh, wl = split(w)
    y = F.conv_transpose2d(ah, wh, stride=2) + F.conv_transpose2d(al, wh, stride=2) + F.conv_transpose2d(ah, wl, stride=2)
    return f32(f32(y) + b.view(1, -1, 1, 1))


def forward(sd, x, mode):
    if mode == 'winograd':
        c3 = conv_winograd
    elif mode == 'winograd4':
        c3 = lambda x, w, b: conv_winograd(x, w, b, 4)
    else:
        c3 = lambda x, w, b: conv_direct(x, w, b, mode)
    up = lambda x, w, b: upconv(x, w, b, 'exact' if mode == 'exact' else 'split')
    g = lambda n: (sd[n + '.weight'], sd[n + '.bias'])
    e11 = F.relu(conv_direct(x, *g('e11'), 'exact' if mode == 'exact' else 'split'))   # Cin = 1: CUDA-core fp32 layer
    e12 = F.relu(c3(e11, *g('e12')))
    e21 = F.relu(c3(F.max_pool2d(e12, 2), *g('e21')))
    e22 = F.relu(c3(e21, *g('e22')))
    e31 = F.relu(c3(F.max_pool2d(e22, 2), *g('e31')))
    e32 = F.relu(c3(e31, *g('e32')))
    d31 = F.relu(c3(torch.cat([up(e32, *g('upconv3')), e22], 1), *g('d31')))
    d32 = F.relu(c3(d31, *g('d32')))
    d41 = F.relu(c3(torch.cat([up(d32, *g('upconv4')), e12], 1), *g('d41')))
    d42 = F.relu(c3(d41, *g('d42')))
    z = F.conv2d(d42, *g('outconv'))
    return torch.sigmoid(z) * 255.


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 102
    sd = {k: torch.from_numpy(v).to(D) for k, v in uo.numpy_weights(2, seed=seed).items()}
    x = torch.from_numpy(np.random.default_rng(seed).integers(0, 256, (1, 1, size, size)).astype(np.float64) / 255.)
    ref = forward(sd, x, 'exact')
    for mode in ('split', 'winograd', 'winograd4'):
        y = forward(sd, x, mode)
        err = (y - ref).abs()
        print(f'{mode:9s}: max |x_hat - x_hat_fp64| = {err.max().item():.3e} px, mean {err.mean().item():.3e} px ({size}x{size}, seed {seed})')


if __name__ == '__main__':
    main()
