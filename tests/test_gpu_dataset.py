"""GPU: the 'next' rows of SURVEY.md section 8f - run()-compatible batched dataset driver (N1), WSLoss / WSMeter
definitions (N2), ROC from beta_hat (N3)."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import unet_oracle as uo
from oracle import ws_oracle as wo

pytestmark = pytest.mark.gpu


def _write_dataset(root, n=6, alpha=0.4, hw=(64, 96)):
    from PIL import Image
    from ws_unet_b200 import data as wdata
    (root / 'images').mkdir(parents=True)
    sdir = root / f'stego_LSBr_alpha_{alpha}_independent_images'   # on-disk spelling of the shipped data (F11)
    sdir.mkdir()
    rows_c, rows_s, imgs = [], [], {}
    for i in range(n):
        c = wdata.synthetic_cover(i, *hw)
        s = wdata.embed_lsbr(c, alpha, i)
        Image.fromarray(c.numpy()).save(root / 'images' / f'{i}.png')
        Image.fromarray(s.numpy()).save(sdir / f'{i}.png')
        rows_c.append(dict(name=f'images/{i}.png', height=hw[0], width=hw[1]))
        rows_s.append(dict(name=f'stego_LSBR_alpha_{alpha}_independent_images/{i}.png', height=hw[0], width=hw[1],
                           stego_method='LSBR', alpha=alpha))
        imgs[f'images/{i}.png'], imgs[rows_s[-1]['name']] = c.numpy(), s.numpy()
    pd.DataFrame(rows_c).to_csv(root / 'images' / 'files.csv', index=False)
    pd.DataFrame(rows_s).to_csv(sdir / 'files.csv', index=False)
    return imgs


def test_run_matches_oracle_per_file(cuda_dev, tmp_path):
    from ws_unet_b200 import dataset as D
    import ws_unet_b200 as W
    imgs = _write_dataset(tmp_path)
    for stego, alpha in [(None, None), ('LSBR', 0.4)]:
        for model_name, weighted in [('KB', 0), ('KB', 1), ('AVG', 1)]:
            df = D.run(tmp_path, stego, alpha, model_name, channels=(3,), weighted=weighted, batch=4)
            assert {'name', 'beta_hat', 'channels', 'weighted', 'correct_bias', 'model_name'} <= set(df.columns)
            assert len(df) == 6 and (df['channels'] == '3').all()
            for _, r in df.iterrows():
                key = '/'.join(r['name'].split('/')[-2:])
                ref = uo.ws_attack_c(imgs[key], kind={'KB': 0, 'AVG': 1}[model_name], weighted=weighted)[0]
                assert abs(r['beta_hat'] - ref) < 1e-4
    sd = uo.numpy_weights(2, seed=31)
    m = W.get_model('unet_2', 1).to(cuda_dev)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    df = D.run(tmp_path, 'LSBR', 0.4, 'my_unet', predictor=m, weighted=0, batch=4)
    assert (df['model_name'] == 'UNet').all()
    for _, r in df.iterrows():
        im = imgs['/'.join(r['name'].split('/')[-2:])]
        xhat = uo.unet_forward(sd, (im.astype(np.float32) / np.float32(255.))[None, None], 2)[0, 0, 1:-1, 1:-1] * np.float32(255.)
        assert abs(r['beta_hat'] - uo.ws_attack_c(im, xhat=xhat, weighted=0)[0]) < 1e-4


def test_wsloss_wsmeter_match_reference_definitions(cuda_dev, ws_golden):
    from ws_unet_b200.metrics import WSLoss, WSMeter, ws_betas_hat
    xin = torch.from_numpy(ws_golden['wsloss_xin']).to(cuda_dev)
    xout = torch.from_numpy(ws_golden['wsloss_xout']).to(cuda_dev)
    got = ws_betas_hat(xout, xin, crop=0).cpu().numpy()
    assert np.abs(got - ws_golden['wsloss_betas_hat']).max() < 1e-4      # reference WSLoss._error with betas = 0
    alphas = torch.tensor([0.1, 0.2, 0.4])
    loss = WSLoss()(xout, (None, alphas.to(cuda_dev)), xin).item()
    assert abs(loss - np.mean(np.abs(ws_golden['wsloss_betas_hat'] - alphas.numpy() / 2))) < 1e-4
    meter = WSMeter()
    meter.update(ws_golden['wsloss_xin'], ws_golden['wsloss_xout'], alphas.numpy())
    ref = wo.wsloss_betas(ws_golden['wsloss_xout'], ws_golden['wsloss_xin'], crop=1)
    assert abs(meter.avg - np.mean(np.abs(ref - alphas.numpy() / 2))) < 1e-4


def test_roc_from_gpu_beta_hat(cuda_dev):
    """N3: the consumer of beta_hat. KB-WS on synthetic covers vs alpha=0.4 stego must separate almost perfectly."""
    import ws_unet_b200 as W
    from ws_unet_b200 import data as wdata
    from ws_unet_b200.metrics import produce_roc
    covers = wdata.synthetic_covers(24, 256, 256)
    stego = torch.stack([wdata.embed_lsbr(covers[i, 0], 0.4, i) for i in range(24)])[:, None]
    bc = W.ws_estimate(covers.to(cuda_dev), 'KB', weighted=1).cpu().numpy()
    bs = W.ws_estimate(stego.to(cuda_dev), 'KB', weighted=1).cpu().numpy()
    df = pd.DataFrame({'stego_method': ['Cover'] * 24 + ['LSBR'] * 24, 'model_name': 'KB', 'alpha': [0.] * 24 + [0.4] * 24,
                       'beta_hat': np.concatenate([bc, bs])})
    roc = produce_roc(df)
    assert roc['auc'].iloc[0] > 0.95 and roc['p_e'].iloc[0] < 0.1


def test_ws_losses_backward_matches_torch_autograd(cuda_dev):
    """WSLoss / L1WSLoss (src/_defs/losses.py:45-115) are differentiable in the predictor output: the gradient from
    wsu_ws_grad_prediction must equal torch autograd applied to the reference's own formula."""
    from ws_unet_b200.metrics import L1WSLoss, WSLoss
    g = torch.Generator().manual_seed(3)
    B, H, Wd = 5, 24, 40
    inputs = (torch.randint(0, 256, (B, 1, H, Wd), generator=g).float() / 255.).to(cuda_dev)
    covers = (torch.randint(0, 256, (B, 1, H, Wd), generator=g).float() / 255.).to(cuda_dev)
    alphas = torch.tensor([0.0, 0.1, 0.4, 1.0, 0.2], device=cuda_dev)
    base = (inputs + 0.02 * torch.randn(B, 1, H, Wd, generator=g).to(cuda_dev)).clamp(0, 1)
    sgn = 2. * (torch.round(inputs[1] * 255.).int() & 1).float() - 1.
    base[1] = inputs[1] + 0.3 * sgn / 255.      # beta_hat = -0.3 on this image: relu gate closed, WS gradient 0 there

    def ref_ws(outputs):                        # losses.py:46-89 verbatim in torch ops
        x, o = inputs * 255., outputs * 255.
        xbar = (torch.round(x).int() ^ 1).float()
        w = torch.ones_like(x) / (torch.numel(x) / float(x.size(0)))
        bh = torch.relu(torch.sum(w * (x - xbar) * (x - o), dim=(1, 2, 3)))
        return torch.mean(torch.abs(bh - alphas / 2.))

    for ours, ref in ((WSLoss(), ref_ws), (L1WSLoss(), lambda o: torch.mean(torch.abs(covers - o)) + ref_ws(o))):
        o1 = base.clone().requires_grad_(True)
        o2 = base.clone().requires_grad_(True)
        l1 = ours(o1, (covers, alphas), inputs)
        l2 = ref(o2)
        l1.backward()
        l2.backward()
        assert abs(l1.item() - l2.item()) < 1e-5
        assert torch.allclose(o1.grad, o2.grad, rtol=1e-5, atol=1e-9)
        assert o1.grad[2].abs().max() > 0 and o1.grad[1].abs().max() == (0 if isinstance(ours, WSLoss) else o1.grad[1].abs().max())


def test_filter_residuals_match_reference_golden(cuda_dev, ws_golden, defs_golden):
    """get_filter_residuals (matrix form, float64) with the reference's argument list, and the
    batched stencil-kernel residual map, both equal the reference's residuals exactly on 8-bit pixels; the MAE of
    results/prediction/filters.csv is their mean absolute value (also what ws_estimate(..., return_l1=True) returns)."""
    import ws_unet_b200 as W
    from ws_unet_b200 import defs, filters
    proc = defs.get_processor(channels=(3,))
    for tag in ('cover', 'lsbr04'):
        img = ws_golden[f'img_{tag}']
        x4 = np.repeat(img[..., None], 4, axis=2)
        d = torch.from_numpy(img)[None, None].to(cuda_dev)
        for name in ('KB', 'AVG'):
            ref = defs_golden[f'resid_{tag}_{name}']
            got = filters.get_filter_residuals('mem', filter=filters.NAMED_FILTERS[name], process_image=proc, imread=lambda f: x4)
            assert got.dtype == np.float64 and got.shape == ref.shape and np.array_equal(got, ref)
            stencil = filters.filter_residuals(d, name)[0].cpu().numpy()
            assert np.array_equal(stencil.reshape(-1, 1), ref)
            mae = defs_golden[f'mae_{tag}_{name}'][0]
            assert abs(np.abs(got).mean() - mae) < 1e-12
            _, l1 = W.ws_estimate(d, name, weighted=0, return_l1=True)
            assert abs(l1.item() - mae) < 1e-4
    x4 = np.repeat(ws_golden['img_cover'][..., None], 4, axis=2)
    got = filters.get_filter_residuals('mem', filter=defs_golden['coef_ols'], imread=lambda f: x4,
                                       process_image=defs.get_processor(channels=(3,), inbayer='01'))
    assert np.abs(got - defs_golden['resid_cover_ols']).max() < 1e-10     # fitted vector: float64 matvec, order of sums differs


def test_new_entry_points_reject_bad_arguments(cuda_dev):
    """wsu_uniform_dropout / wsu_filter_residual_rows / correct_bias: invalid arguments come back as WSU_ERR_INVALID with a
    message (ValueError in the shim), never as a launch."""
    import ctypes
    import ws_unet_b200 as W
    from ws_unet_b200 import _native
    lib = _native.load()
    x = torch.rand(1, 1, 8, 8, device=cuda_dev)
    m = torch.ones(1, 1, 8, 8, device=cuda_dev)
    o = torch.empty_like(x)
    st = _native.stream_ptr(cuda_dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.wsu_uniform_dropout(0, p(x), 1, p(m), p(o), 1, 1, 1, 8, 1, st) == _native.WSU_ERR_INVALID      # H < 2
    assert lib.wsu_uniform_dropout(0, None, 1, p(m), p(o), 1, 1, 8, 8, 1, st) == _native.WSU_ERR_INVALID
    assert lib.wsu_uniform_dropout(0, p(x), 7, p(m), p(o), 1, 1, 8, 8, 1, st) == _native.WSU_ERR_INVALID
    assert lib.wsu_uniform_dropout(0, p(x), 1, p(m), p(o), 1, 1, 8, 8, 0, st) == 0                             # no channel selected: a copy
    torch.cuda.synchronize()
    assert torch.equal(o, x)
    mat = torch.zeros(4, 9, dtype=torch.float64, device=cuda_dev)
    coef = torch.zeros(8, dtype=torch.float64, device=cuda_dev)
    out = torch.empty(4, dtype=torch.float64, device=cuda_dev)
    assert lib.wsu_filter_residual_rows(0, p(mat), 5, p(coef), p(out), 4, st) == _native.WSU_ERR_INVALID
    assert lib.wsu_filter_residual_rows(0, p(mat), 2, p(coef), p(out), 0, st) == _native.WSU_ERR_INVALID
    model = W.get_model('unet_1', 1).to(cuda_dev)
    with pytest.raises(ValueError):
        W.ws_estimate(torch.rand(1, 1, 16, 16, device=cuda_dev), model, correct_bias=True)     # bias correction needs uint8 pixels
    with pytest.raises(ValueError):
        _native.check(lib.wsu_set_option(model.native_handle(cuda_dev), b'precision', 4))


def test_filter_residual_rows_float_inputs(cuda_dev):
    """get_filter_residuals with an arbitrary (OLS-style) coefficient vector on float32 / float64 neighbour matrices."""
    import ws_unet_b200 as W
    rng = np.random.default_rng(4)
    coef = rng.normal(0, 0.3, (8, 1))
    for dt in (np.float32, np.float64, np.uint8, np.int16):
        mat = (rng.random((1000, 9)) * 255).astype(dt)
        got = W.filters.get_filter_residuals('mem', coef, process_image=lambda x: x, imread=lambda f: mat, device=cuda_dev)
        m64 = mat.astype(np.float64)
        ref = m64[:, -1:] - m64[:, :-1] @ coef
        assert got.shape == (1000, 1) and got.dtype == np.float64
        assert np.abs(got - ref).max() < 1e-10
