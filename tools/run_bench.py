"""Dataset-level benchmark of ws_unet_b200.dataset.run (the drop-in for src/ws/estimate.py:149-205 `run`): N PNG files on
disk -> DataFrame of beta_hat, with the decode-only ceiling of the same thread pool next to it.
Usage (on a B200): python tools/run_bench.py [n_files=5000] [dir=/dev/shm/wsu_run_bench]"""
import concurrent.futures
import json
import os
import pathlib
import shutil
import sys
import time

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, '.')
import ws_unet_b200 as W
from ws_unet_b200 import data as wdata
from ws_unet_b200 import dataset as D


def make_dataset(root: pathlib.Path, n: int):
    from PIL import Image
    if root.exists():
        shutil.rmtree(root)
    (root / 'stego_LSBr_alpha_0.4').mkdir(parents=True)
    covers = [wdata.embed_lsbr(wdata.synthetic_cover(i), 0.4, i).numpy() for i in range(32)]
    names = [f'stego_LSBr_alpha_0.4/{i:06d}.png' for i in range(n)]

    def write(i):
        Image.fromarray(covers[i % 32]).save(root / names[i], compress_level=1)

    with concurrent.futures.ThreadPoolExecutor(os.cpu_count() or 8) as pool:
        list(pool.map(write, range(n)))
    pd.DataFrame({'name': names, 'stego_method': 'LSBR', 'alpha': 0.4}).to_csv(root / 'stego_LSBr_alpha_0.4' / 'files.csv', index=False)
    return sum((root / nm).stat().st_size for nm in names[:32]) / 32


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    root = pathlib.Path(sys.argv[2] if len(sys.argv) > 2 else '/dev/shm/wsu_run_bench')
    dev = torch.device('cuda', 0)
    t0 = time.perf_counter()
    png_bytes = make_dataset(root, n)
    t_make = time.perf_counter() - t0
    paths = [root / nm for nm in D.list_files(root, stego_method='LSBR', alpha=0.4)['name']]
    workers = min(32, os.cpu_count() or 4)
    out = {'files': n, 'png_bytes_avg': png_bytes, 'workers': workers, 'cores': os.cpu_count(), 'dataset_write_s': round(t_make, 2)}

    # decode-only ceiling: the same reader on the same thread pool, results thrown away
    with concurrent.futures.ThreadPoolExecutor(workers) as pool:
        list(pool.map(D.imread_gray_u8, paths[:256]))
        t0 = time.perf_counter()
        list(pool.map(D.imread_gray_u8, paths))
        out['decode_only_images_per_s'] = n / (time.perf_counter() - t0)

    torch.manual_seed(1234)
    model = W.get_model('unet_2', 1).to(dev)
    model.calibrate_precision(torch.stack([wdata.embed_lsbr(wdata.synthetic_cover(i), 0.4, i) for i in range(8)])[:, None].to(dev))
    for name, kw in (('KB_w0', dict(model_name='KB', weighted=0)), ('KB_w1', dict(model_name='KB', weighted=1)),
                     ('UNet_w0', dict(model_name='UNet', predictor=model, weighted=0))):
        D.run(root, 'LSBR', 0.4, take_num_images=512, **kw)          # warm-up: pinned ring, plans, kernels
        t0 = time.perf_counter()
        df = D.run(root, 'LSBR', 0.4, **kw)
        dt = time.perf_counter() - t0
        assert len(df) == n and df['beta_hat'].notna().all()
        out[f'run_{name}_images_per_s'] = n / dt
        out[f'run_{name}_beta_mean'] = float(df['beta_hat'].mean())
    print(json.dumps(out))
    shutil.rmtree(root, ignore_errors=True)


if __name__ == '__main__':
    main()
