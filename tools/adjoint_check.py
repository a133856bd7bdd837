"""A/B check of the adjoint (parity-plane) estimator kernel against the packed 16-bit-lane kernel and the exact
integer value, over ragged shapes. Run twice: default and WSU_EST_KERNEL=packed; beta_hat must be bit-identical."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import ws_unet_b200 as W

dev = torch.device('cuda', 0)
rng = np.random.default_rng(5)
shapes = [(3, 16), (4, 16), (5, 32), (64, 48), (65, 512), (66, 528), (130, 1024), (512, 512), (200, 1040), (67, 2064)]
bad = 0
for h, w in shapes:
    for nb in (1, 9):
        img = rng.integers(0, 256, (nb, 1, h, w), dtype=np.uint8)
        if nb == 9:
            img[1] = 255
            img[2] = (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)
            img[3] = 0
        d = torch.from_numpy(img).to(dev)
        for name, D in (('KB', 4), ('AVG', 8)):
            got = W.ws_estimate(d, name, weighted=0, clip=False).cpu().numpy()
            x = img[:, 0].astype(np.int64)
            c = x[:, 1:-1, 1:-1]
            cross = x[:, :-2, 1:-1] + x[:, 2:, 1:-1] + x[:, 1:-1, :-2] + x[:, 1:-1, 2:]
            diag = x[:, :-2, :-2] + x[:, :-2, 2:] + x[:, 2:, :-2] + x[:, 2:, 2:]
            R = 4 * c - 2 * cross + diag if name == 'KB' else 8 * c - cross - diag
            s = np.where(c & 1, 1, -1)
            tot = (s * R).sum(axis=(1, 2))
            exact = ((tot / D) / float((h - 2) * (w - 2))).astype(np.float32)
            ok = np.array_equal(got, exact)
            bad += not ok
            if not ok:
                print('MISMATCH', h, w, nb, name, got[:4], exact[:4])
print('kernel:', os.environ.get('WSU_EST_KERNEL', 'adjoint'), 'mismatches:', bad)
sys.exit(1 if bad else 0)
