"""CPU study (not product code): worst-case prediction error of the candidate precision PLANS over many images and weight
seeds (float64 emulation with explicit roundings, tools/precision_budget_study.py).
  split    : every layer three-term split-bf16 (shipped in round 1)
  deep1    : layers whose input lives at level >= 1 (e21 e22 e31 e32 upconv3 d31 d32 upconv4) read ONE fp16 activation
             against ONE fp16 weight (a single MMA per MAC); e12, d41, d42 stay three-term
  deep2    : same layers, fp16 activation against fp16 (hi, lo) weights (two MMAs per MAC)
  deep1+up : deep1, and d41 reads its upsampled half (u4) as one fp16 value against one fp16 weight
Usage: python tools/precision_plan_study.py <size> <n_images> <seeds,comma> [real]"""
import importlib.util
import pathlib
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
spec = importlib.util.spec_from_file_location('pbs', 'tools/precision_budget_study.py')
pbs = importlib.util.module_from_spec(spec)
spec.loader.exec_module(pbs)
from oracle import unet_oracle as uo
from ws_unet_b200 import data as wdata

D = torch.float64
DEEP = ['e21', 'e22', 'e31', 'e32', 'upconv3', 'd31', 'd32', 'upconv4']
PLANS = {
    'split': {'*': 'split'},
    'deep1': dict({'*': 'split'}, **{l: 'a16w16' for l in DEEP}),
    'deep2': dict({'*': 'split'}, **{l: 'a16' for l in DEEP}),
    'deep1+up': dict({'*': 'split', 'd41.up': 'a16w16'}, **{l: 'a16w16' for l in DEEP}),
}


def main():
    size, nimg = int(sys.argv[1]), int(sys.argv[2])
    seeds = [int(s) for s in sys.argv[3].split(',')]
    real = len(sys.argv) > 4 and sys.argv[4] == 'real'
    for seed in seeds:
        sd = {k: torch.from_numpy(v).to(D) for k, v in uo.numpy_weights(2, seed=seed).items()}
        if real:
            from PIL import Image
            imgs = [torch.from_numpy(np.array(Image.open(p))[:size, :size].copy()) for p in sorted(pathlib.Path('/root/reference/data/images').glob('*.png'))]
        else:
            imgs = [wdata.embed_lsbr(wdata.synthetic_cover(1000 * seed + i, size, size), [0.01, 0.05, 0.1, 0.2, 0.4, 1.0][i % 6], i) for i in range(nimg)]
        worst = {k: 0. for k in PLANS}
        for im in imgs:
            x = im[None, None].to(D) / 255.
            ref = pbs.forward(sd, x, {'*': 'exact'})
            for k, md in PLANS.items():
                worst[k] = max(worst[k], (pbs.forward(sd, x, md) - ref).abs().max().item())
        print(f'seed {seed:4d} {"real covers" if real else "synthetic"} {len(imgs)} x {size}x{size}: ' +
              '  '.join(f'{k} {v:.3e}' for k, v in worst.items()) + '  px max-abs vs float64', flush=True)


if __name__ == '__main__':
    main()
