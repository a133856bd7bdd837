"""CPU, only where the reference mount exists: check the oracle against the reference's own published artefacts
(results/prediction/filters.csv, reproduced bit-for-bit by the survey) and shipped stego images."""
import csv
import pathlib

import numpy as np
import pytest

from oracle import ws_oracle as wo

REF = pathlib.Path('/root/reference')
pytestmark = pytest.mark.skipif(not (REF / 'results/prediction/filters.csv').exists(), reason='reference mount absent')


def _read(path):
    from PIL import Image
    return np.array(Image.open(path))


def _matrix_mae(img, name):
    """src/filters/evaluate.py:53-76,98 + src/_defs/filters.py:39-69: N x 9 neighbour matrix (float32) @ 8x1 float64
    filter, residual y - y_hat, nanmean |resid|."""
    x = img.astype('float32')
    cols = [x[:-2, :-2], x[:-2, 1:-1], x[:-2, 2:], x[1:-1, 2:], x[2:, 2:], x[2:, 1:-1], x[2:, :-2], x[1:-1, :-2], x[1:-1, 1:-1]]
    m = np.stack([c.flatten() for c in cols], axis=-1)
    filt = {'KB': np.array([[-1], [2], [-1], [2], [-1], [2], [-1], [2]], dtype='float64') / 4., 'AVG': np.ones((8, 1)) / 8.}[name]
    resid = m[..., -1:] - m[..., :-1] @ filt
    return np.nanmean(np.abs(resid))


def test_filters_csv_reproduced():
    rows = list(csv.DictReader(open(REF / 'results/prediction/filters.csv')))
    checked = 0
    for r in rows:
        img = _read(REF / 'data' / r['name'])
        for name in ('KB', 'AVG'):
            v = r[f'mae_3_{name}']
            if v:
                assert abs(_matrix_mae(img, name) - float(v)) < 1e-12
                # the oracle's 2-D stencil predictor gives the same MAE as the matrix form
                pred = wo.filter_predict_exact(img[..., None].astype(np.float64), name)[..., 0]
                assert abs(np.mean(np.abs(img[1:-1, 1:-1] - pred)) - float(v)) < 1e-9
                checked += 1
    assert checked == 10


@pytest.mark.parametrize('alpha,rate', [('0.01', .0050), ('0.1', .0499), ('0.4', .2003), ('1.0', .5006)])
def test_shipped_lsbr_semantics(alpha, rate):
    cover = _read(REF / 'data/images/6.png')
    diffs, n = 0, 0
    for i in range(6, 11):
        c = _read(REF / f'data/images/{i}.png')
        s = _read(REF / f'data/stego_LSBr_alpha_{alpha}_independent_images/{i}.png')
        assert np.all((c ^ s) <= 1)                 # only LSB flips
        diffs += (c != s).sum()
        n += c.size
    assert abs(diffs / n - rate) < 2e-4             # measured rates, SURVEY.md section 4
    assert cover.dtype == np.uint8


def _import_reference():
    import importlib.util
    import os
    import sys
    spec = importlib.util.spec_from_file_location('make_golden', pathlib.Path(__file__).parent / 'golden' / 'make_golden.py')
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    cwd = os.getcwd()
    try:
        return mg.import_reference()
    finally:
        os.chdir(cwd)
        sys.path[:] = [p for p in sys.path if not p.startswith(str(REF))]


def test_produce_roc_matches_reference(capsys):
    import pandas as pd
    from ws_unet_b200.metrics import produce_roc
    _defs, rfilters, runet, rws = _import_reference()
    rng = np.random.default_rng(3)
    rows = []
    for m in ['KB', 'UNet']:
        for i in range(40):
            rows.append(dict(stego_method='Cover', model_name=m, alpha=0., beta_hat=rng.normal(0, 0.02)))
        for i in range(60):
            rows.append(dict(stego_method='LSBR', model_name=m, alpha=0.05, beta_hat=rng.normal(0.025, 0.02)))
    df = pd.DataFrame(rows)
    ref = rws.roc.produce_roc(df).reset_index(drop=True)
    got = produce_roc(df).reset_index(drop=True)
    for col in ['tau', 'tpr', 'fpr', 'p_e', 'tau0', 'auc', 'fpr_50', 'tpr_50']:
        assert np.allclose(ref[col].to_numpy(dtype=float), got[col].to_numpy(dtype=float), atol=1e-12), col
    assert list(ref['label']) == list(got['label'])


def test_list_files_matches_fabrika_selection():
    from ws_unet_b200 import dataset as D
    covers = D.list_files(REF / 'data')
    assert covers['name'].tolist() == sorted(f'images/{i}.png' for i in range(6, 11))
    st = D.list_files(REF / 'data', stego_method='LSBR', alpha=0.4)
    assert len(st) == 5 and set(st['alpha']) == {0.4} and set(st['stego_method']) == {'LSBR'}
    # files.csv says LSBR, the directory on disk is LSBr: paths must still resolve
    assert all(D._resolve_case(REF / 'data' / n).exists() for n in st['name'])
    st2 = D.list_files(REF / 'data', stego_method='HILLR', alpha=0.05, take_num_images=2)
    assert len(st2) == 2
    # split=: a CSV inside the dataset replaces the globbed files.csv tables (src/fabrika.py:51-52)
    import pandas as pd
    import tempfile, pathlib
    with tempfile.TemporaryDirectory() as td:
        td = pathlib.Path(td)
        pd.DataFrame({'name': ['images/b.png', 'images/a.png', 'stego/x.png'], 'stego_method': [None, None, 'LSBR'],
                      'alpha': [None, None, 0.4], 'device': ['007', '007', '007']}).to_csv(td / 'split_te.csv', index=False)
        assert D.list_files(td, split='split_te.csv')['name'].tolist() == ['images/a.png', 'images/b.png']
        assert D.list_files(td, split='split_te.csv', stego_method='LSBR', alpha=0.4)['name'].tolist() == ['stego/x.png']
        assert D.list_files(td, split='split_te.csv')['device'].tolist() == ['007', '007']
