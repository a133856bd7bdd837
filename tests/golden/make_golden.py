"""Generates the golden fixtures in tests/golden/ by running the UNMODIFIED reference (/root/reference/src) in
this container. Inert stub modules stand in for imports the hot path never uses (SURVEY.md section 8c recipe).
Run: python tests/golden/make_golden.py   (needs /root/reference; the GPU box never runs this)."""
import importlib.machinery
import pathlib
import sys
import types

import numpy as np
import torch

REPO = pathlib.Path(__file__).resolve().parents[2]
REF = pathlib.Path('/root/reference')
sys.path.insert(0, str(REPO))


def import_reference():
    for name in ['timm', 'conseal', 'jpeglib', 'seaborn', 'matplotlib', 'matplotlib.pyplot', 'torchinfo']:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            m.__path__ = []
            sys.modules[name] = m
    sys.path.insert(0, str(REF / 'src'))
    import os
    os.chdir(REF / 'src')  # the reference appends 'unet', 'detector', '.' to sys.path relative to cwd=src/
    import _defs, filters, unet, ws  # noqa
    import unet.model  # noqa
    import ws.estimate  # noqa
    return _defs, filters, unet, ws


def main():
    from oracle import unet_oracle as uo
    from ws_unet_b200 import data as wdata
    from PIL import Image
    _defs, rfilters, runet, rws = import_reference()
    torch.set_num_threads(8)
    out = REPO / 'tests' / 'golden'

    # ------------------------------------------------------------------ UNet goldens (reference torch module)
    g = {}
    for nsteps, hw in [(0, (24, 40)), (1, (32, 48)), (2, (64, 64)), (2, (40, 72)), (3, (64, 64)), (4, (64, 96))]:
        sd = uo.numpy_weights(nsteps, seed=100 + nsteps)
        model = runet.model.get_model(f'unet_{nsteps}', in_channels=1, out_channels=1, channel=[0], drop_rate=0.)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        x = np.random.default_rng(200 + nsteps).random((2, 1) + hw, dtype=np.float32)
        with torch.no_grad():
            y = model(torch.from_numpy(x.copy())).numpy()
        key = f'unet{nsteps}_{hw[0]}x{hw[1]}'
        g[key + '_x'] = x
        g[key + '_y'] = y
    # 512x512 through the reference's own infere_single / predict_unet / attack (config 1 shape)
    sd = uo.numpy_weights(2, seed=102)
    model = runet.model.get_model('unet_2', in_channels=1, out_channels=1, channel=[0], drop_rate=0.)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    stego = wdata.embed_lsbr(wdata.synthetic_cover(0), 0.4, 0).numpy()
    x4 = np.repeat(stego[..., None], 4, axis=2)  # imread4-style (H,W,4) with Y in channel 3
    xhat = runet.infere_single(x4[..., 3:].astype('float32'), model)
    res = runet.evaluate.predict_unet('mem', model, imread=lambda f: x4.astype('float32'))
    g['unet2_512_stego_sha'] = np.frombuffer(__import__('hashlib').sha256(stego.tobytes()).digest(), dtype=np.uint8)
    g['unet2_512_xhat_sub'] = xhat[::7, ::7, 0].astype(np.float32)
    g['unet2_512_beta_l1'] = np.array([res['beta_hat'], res['l1']], dtype=np.float64)
    proc = _defs.get_processor_2d(channels=(3,))
    est = lambda v: runet.infere_single(v, model)
    att = []
    for weighted in (0, 1, -1):
        r = rws.estimate.attack('mem', channels=(3,), pixel_estimator=est, correct_bias=False, weighted=weighted,
                                imread=lambda f: x4, process_image=proc)
        att.append(r['beta_hat'])
    r = rws.estimate.attack('mem', channels=(3,), pixel_estimator=est, correct_bias=True, weighted=1,
                            imread=lambda f: x4, process_image=proc)
    att.append(r['beta_hat'])
    g['unet2_512_attack_w0_w1_wm1_w1bias'] = np.array(att, dtype=np.float64)
    np.savez_compressed(out / 'unet_golden.npz', **g)

    # ------------------------------------------------------------------ WS / filter goldens on the shipped images
    w = {}
    crops = {}
    cover = np.array(Image.open(REF / 'data/images/6.png'))
    for tag, path in [('cover', 'data/images/6.png'),
                      ('lsbr04', 'data/stego_LSBr_alpha_0.4_independent_images/6.png'),
                      ('lsbr10', 'data/stego_LSBr_alpha_1.0_independent_images/6.png'),
                      ('hill04', 'data/stego_HILLr_alpha_0.4_independent_images/6.png')]:
        im = np.array(Image.open(REF / path))
        assert im.shape == (512, 512) and im.dtype == np.uint8
        crops[tag] = im[192:320, 160:320].copy()  # 128 x 160 crop
    for tag, im in crops.items():
        w[f'img_{tag}'] = im
        x4 = np.repeat(im[..., None], 4, axis=2)
        for name in ('KB', 'AVG', 'AVG9', '1'):
            est = rfilters.get_filter_estimator(filter_name=name, flatten=False)
            w[f'pred_{tag}_{name}'] = est(x4[..., 3:].astype('float32'))[..., 0].astype(np.float32)
            vals = []
            for weighted in (0, 1, -1):
                for bias in (False, True):
                    r = rws.estimate.attack('mem', channels=(3,), pixel_estimator=est, correct_bias=bias,
                                            weighted=weighted, imread=lambda f: x4, process_image=proc)
                    vals.append(r['beta_hat'])
            w[f'beta_{tag}_{name}'] = np.array(vals, dtype=np.float64)  # order: (w0,nb),(w0,b),(w1,nb),(w1,b),(w-1,nb),(w-1,b)
    # full-image KB/AVG values the survey recorded - recomputed here, used only for bookkeeping
    full = np.array(Image.open(REF / 'data/stego_LSBr_alpha_0.4_independent_images/6.png'))
    x4 = np.repeat(full[..., None], 4, axis=2)
    est = rfilters.get_filter_estimator(filter_name='KB', flatten=False)
    w['full6_lsbr04_KB_w0_w1'] = np.array([
        rws.estimate.attack('mem', channels=(3,), pixel_estimator=est, weighted=wt, imread=lambda f: x4, process_image=proc)['beta_hat']
        for wt in (0, 1)], dtype=np.float64)
    # batched training-side definitions (WSLoss / WSMeter)
    rng = np.random.default_rng(5)
    xin = (rng.integers(0, 256, (3, 1, 32, 48)) / 255.).astype(np.float32)
    xout = np.clip(xin + rng.normal(0, 0.01, xin.shape), 0, 1).astype(np.float32)
    loss = _defs.losses.WSLoss()
    betas = np.zeros(3, dtype=np.float32)
    err = loss._error(torch.from_numpy(xout), torch.from_numpy(xin), torch.from_numpy(betas)).numpy()
    w['wsloss_xin'], w['wsloss_xout'], w['wsloss_betas_hat'] = xin, xout, err.astype(np.float64)
    np.savez_compressed(out / 'ws_golden.npz', **w)

    # ------------------------------------------------------------------ synthetic generator fingerprint
    import hashlib
    fp = {}
    for i in (0, 1, 7):
        c = wdata.synthetic_cover(i).numpy()
        fp[f'cover{i}_sha'] = np.frombuffer(hashlib.sha256(c.tobytes()).digest(), dtype=np.uint8)
        fp[f'cover{i}_stats'] = np.array([c.min(), c.max(), c.mean(), c.std()], dtype=np.float64)
    s = wdata.embed_lsbr(wdata.synthetic_cover(0), 0.4, 0).numpy()
    fp['stego0_rate'] = np.array([(s != wdata.synthetic_cover(0).numpy()).mean()])
    np.savez_compressed(out / 'data_golden.npz', **fp)
    print('golden fixtures written to', out)
    for f in out.glob('*.npz'):
        print(f.name, f.stat().st_size)


if __name__ == '__main__':
    main()
