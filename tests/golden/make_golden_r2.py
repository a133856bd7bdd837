"""Round-2 golden fixtures at BASELINE.json's configuration sizes, produced by the UNMODIFIED reference
(/root/reference/src, imported with the inert stubs of make_golden.py). Writes tests/golden/config_golden.npz.
Run: python tests/golden/make_golden_r2.py   (needs /root/reference; takes ~3 min on 8 cores; the GPU box never runs this).

Everything a test needs to rebuild the inputs is deterministic and lives in the repo: the synthetic cover / LSBr
generator (ws_unet_b200/data.py) and the reference's own parameter initialisation under torch.manual_seed(1234)
(tests/test_host_logic.py pins that ws_unet_b200.get_model draws the same weights as the reference).
"""
import hashlib
import importlib.util
import json
import os
import pathlib
import sys
import tempfile

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
REPO = HERE.parents[1]
sys.path.insert(0, str(REPO))

MODEL_SEED = 1234            # bench.py::build_model
CFG1_N, CFG1_ALPHA = 64, 0.4   # BASELINE.json configs[0]
CFG3_ALPHAS, CFG3_PER = [0.01, 0.05, 0.1, 0.2, 0.4, 1.0], 8   # configs[2] (subsample)
CFG3_START = 100


def weights_sha(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def main():
    spec = importlib.util.spec_from_file_location('make_golden', HERE / 'make_golden.py')
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    cwd = os.getcwd()
    _defs, rfilters, runet, rws = mg.import_reference()
    os.chdir(cwd)
    import ws.roc as rroc  # noqa  (reference module)
    from ws_unet_b200 import data as wdata
    torch.set_num_threads(os.cpu_count() or 8)
    g = {}

    torch.manual_seed(MODEL_SEED)
    model = runet.model.get_model('unet_2', in_channels=1, out_channels=1, channel=[0], drop_rate=0.)
    g['weights_sha'] = weights_sha(model.state_dict())
    proc = _defs.get_processor_2d(channels=(3,))

    def x4_of(img_u8):
        return np.repeat(img_u8[..., None], 4, axis=2)

    # ---------------------------------------------------------------- config 1: 64 images, alpha = 0.4, predict_unet
    vals = []
    for i in range(CFG1_N):
        st = wdata.embed_lsbr(wdata.synthetic_cover(i), CFG1_ALPHA, i).numpy()
        r = runet.evaluate.predict_unet('mem', model, imread=lambda f, s=st: x4_of(s).astype('float32'))
        vals.append([r['beta_hat'], r['l1']])
        if i in (0, 37):
            xh = runet.infere_single(st[..., None].astype('float32'), model)
            g[f'cfg1_xhat_sub3_{i}'] = xh[::3, ::3, 0].astype(np.float32)
        print('cfg1', i, vals[-1], flush=True)
    g['cfg1_beta_l1'] = np.array(vals, dtype=np.float64)

    # ---------------------------------------------------------------- config 3: alpha sweep, predict_unet + attack(w=0)
    vals = []
    for ai, a in enumerate(CFG3_ALPHAS):
        for k in range(CFG3_PER):
            idx = CFG3_START + ai * CFG3_PER + k
            st = wdata.embed_lsbr(wdata.synthetic_cover(idx), a, idx).numpy()
            r = runet.evaluate.predict_unet('mem', model, imread=lambda f, s=st: x4_of(s).astype('float32'))
            xh = runet.infere_single(st[..., None].astype('float32'), model)
            mae_cover = float(np.mean(np.abs(wdata.synthetic_cover(idx).numpy()[1:-1, 1:-1].astype(np.float32) - xh[..., 0])))
            vals.append([a, r['beta_hat'], r['l1'], mae_cover])
        print('cfg3', a, vals[-1], flush=True)
    g['cfg3_alpha_beta_l1_maecover'] = np.array(vals, dtype=np.float64)
    # one attack() call per alpha (clipped, all weight modes) on the first image of each group
    att = []
    est = lambda v: runet.infere_single(v, model)
    for ai, a in enumerate(CFG3_ALPHAS):
        idx = CFG3_START + ai * CFG3_PER
        st = wdata.embed_lsbr(wdata.synthetic_cover(idx), a, idx).numpy()
        row = []
        for weighted, bias in [(0, False), (1, False), (-1, False), (1, True), (0, True)]:
            r = rws.estimate.attack('mem', channels=(3,), pixel_estimator=est, correct_bias=bias, weighted=weighted,
                                    imread=lambda f, s=st: x4_of(s), process_image=proc)
            row.append(r['beta_hat'])
        att.append(row)
        print('cfg3 attack', a, row, flush=True)
    g['cfg3_attack_w0_w1_wm1_w1b_w0b'] = np.array(att, dtype=np.float64)

    # ---------------------------------------------------------------- config 5: one full 1024x1024 image, model(x) directly
    st = wdata.embed_lsbr(wdata.synthetic_cover(5000, 1024, 1024), 0.4, 5000).numpy()
    x = torch.from_numpy((st.astype('float32') / 255.)[None, None])
    with torch.no_grad():
        y = model(x.clone()).numpy()[0, 0]
    g['cfg5_y_sub5'] = y[::5, ::5].astype(np.float32)
    xf = st.astype('float32')[1:-1, 1:-1]
    xh = y[1:-1, 1:-1] * 255.
    xb = (st[1:-1, 1:-1] ^ 1).astype('float32')
    g['cfg5_beta_l1'] = np.array([np.mean((xf - xb) * (xf - xh)), np.mean(np.abs(xf - xh))], dtype=np.float64)
    print('cfg5', g['cfg5_beta_l1'], flush=True)

    # ---------------------------------------------------------------- infere_single on non-512 inputs (CenterCrop(512) crops / zero-pads)
    for tag, (h, w) in {'crop': (600, 520), 'pad': (400, 512), 'mixed': (530, 300)}.items():
        img = wdata.embed_lsbr(wdata.synthetic_cover(7000 + h, h, w), 0.2, h).numpy()
        xh = runet.infere_single(img[..., None].astype('float32'), model)
        assert xh.shape == (510, 510, 1)
        g[f'infere_{tag}_sub5'] = xh[::5, ::5, 0].astype(np.float32)
        print('infere', tag, float(xh.mean()), flush=True)

    # ---------------------------------------------------------------- file-based entry points: predict_unet(fname), get_unet_estimator
    from PIL import Image
    with tempfile.TemporaryDirectory() as td:
        td = pathlib.Path(td)
        st = wdata.embed_lsbr(wdata.synthetic_cover(9000), 0.4, 9000).numpy()
        Image.fromarray(st).save(td / 'stego.png')
        r = runet.evaluate.predict_unet(td / 'stego.png', model, tag='x')     # default imread = _defs.imread4_f32
        assert r['tag'] == 'x'
        g['file_predict_beta_l1'] = np.array([r['beta_hat'], r['l1']], dtype=np.float64)
        mdir = td / 'models' / 'unet' / 'LSBR' / 'run0'
        (mdir / 'model').mkdir(parents=True)
        json.dump({'network': 'unet_2'}, open(mdir / 'config.json', 'w'))
        torch.save({'state_dict': model.state_dict()}, mdir / 'model' / 'best_model.pt.tar')
        est2 = runet.get_unet_estimator(td / 'models' / 'unet' / 'LSBR', (3,), model_name='run0')
        xh = est2(st[..., None].astype('float32'))
        g['file_estimator_sub5'] = xh[::5, ::5, 0].astype(np.float32)
        r = rws.estimate.attack(td / 'stego.png', channels=(3,), pixel_estimator=est2, correct_bias=False, weighted=1,
                                imread=_defs.imread4_u8, process_image=proc)
        g['file_attack_w1'] = np.array([r['beta_hat']], dtype=np.float64)
        print('file', g['file_predict_beta_l1'], g['file_attack_w1'], flush=True)

    # ---------------------------------------------------------------- produce_roc with beta_hat on both sides of 0.5 (tpr_50 uses a stale FN)
    import pandas as pd
    rng = np.random.default_rng(11)
    rows = []
    for m in ['KB', 'UNet']:
        for i in range(50):
            rows.append(dict(stego_method='Cover', model_name=m, alpha=0., beta_hat=rng.normal(0, 0.02)))
        for i in range(70):
            rows.append(dict(stego_method='LSBR', model_name=m, alpha=1.0, beta_hat=rng.normal(0.5, 0.03)))
    df = pd.DataFrame(rows)
    ref = rroc.produce_roc(df).reset_index(drop=True)
    g['roc_in_beta'] = df['beta_hat'].to_numpy()
    g['roc_in_alpha'] = df['alpha'].to_numpy()
    for col in ['tpr', 'fpr', 'p_e', 'tau0', 'auc', 'fpr_50', 'tpr_50']:
        g['roc_' + col] = ref[col].to_numpy(dtype=np.float64)
    print('roc tpr_50', sorted(set(ref['tpr_50'])), flush=True)

    np.savez_compressed(HERE / 'config_golden.npz', **g)
    print('wrote', HERE / 'config_golden.npz', (HERE / 'config_golden.npz').stat().st_size, 'bytes')


if __name__ == '__main__':
    main()
