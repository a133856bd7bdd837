"""Fixture for the precision-plan error budget test: 256x256 centre crops of the five cover images the reference ships
(/root/reference/data/images/*.png, 512x512 8-bit grayscale) -> tests/golden/real_covers_256.npz (5 x 64 KB).
Run: python tests/golden/make_golden_covers.py   (needs /root/reference; the GPU box never runs this)."""
import pathlib

import numpy as np
from PIL import Image

HERE = pathlib.Path(__file__).resolve().parent
files = sorted(pathlib.Path('/root/reference/data/images').glob('*.png'), key=lambda p: int(p.stem))
imgs = np.stack([np.array(Image.open(f))[128:384, 128:384] for f in files])
assert imgs.shape == (5, 256, 256) and imgs.dtype == np.uint8
np.savez_compressed(HERE / 'real_covers_256.npz', covers=imgs, names=np.array([f.name for f in files]))
print('wrote', HERE / 'real_covers_256.npz', (HERE / 'real_covers_256.npz').stat().st_size, 'bytes')
