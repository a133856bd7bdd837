"""Golden fixtures for the rows either side of the hot path (a13 get_filter_residuals + get_processor,
a14 get_processor_2d + imread4_u8), produced by the UNMODIFIED reference in this container.
Run: python tests/golden/make_golden_defs.py   (needs /root/reference; writes tests/golden/defs_golden.npz and a PNG)."""
import pathlib
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import REF, import_reference  # noqa: E402


def main():
    from PIL import Image
    _defs, rfilters, _, _ = import_reference()
    g = {}
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (13, 18, 4), dtype=np.uint8)          # odd/even sizes exercise every inbayer trimming
    g['proc_img'] = img
    for inb in (None, '00', '01', '10', '11'):
        g[f'proc_{inb}'] = _defs.get_processor(channels=(3,), inbayer=inb)(img)
    g['proc2d_3'] = _defs.get_processor_2d(channels=(3,))(img)
    g['proc2d_02'] = _defs.get_processor_2d(channels=(0, 2))(img)
    # imread4_u8 on a colour PNG written here (RGB differ, so the luma conversion and channel order are both pinned)
    rgb = rng.integers(0, 256, (9, 11, 3), dtype=np.uint8)
    Image.fromarray(rgb).save(HERE / 'rgb_9x11.png')
    g['imread4_rgb'] = _defs.imread4_u8(str(HERE / 'rgb_9x11.png'))
    g['imread_u8_rgb'] = _defs.imread_u8(str(HERE / 'rgb_9x11.png'))
    # get_filter_residuals with the named 8x1 vectors on crops of the shipped images (channel 3 = Y)
    ws = np.load(HERE / 'ws_golden.npz')
    for tag in ('cover', 'lsbr04'):
        x4 = np.repeat(ws[f'img_{tag}'][..., None], 4, axis=2)
        for name in ('KB', 'AVG'):
            r = rfilters.evaluate.get_filter_residuals('mem', filter=rfilters.evaluate.NAMED_FILTERS[name],
                                                       process_image=_defs.get_processor(channels=(3,)), imread=lambda f: x4)
            assert r.dtype == np.float64 and r.shape == (126 * 158, 1)
            g[f'resid_{tag}_{name}'] = r
            g[f'mae_{tag}_{name}'] = np.array([np.nanmean(np.abs(r))])
    # a fitted (non-named) coefficient vector
    coef = rng.normal(0.125, 0.05, (8, 1))
    g['coef_ols'] = coef
    x4 = np.repeat(ws['img_cover'][..., None], 4, axis=2)
    g['resid_cover_ols'] = rfilters.evaluate.get_filter_residuals(
        'mem', filter=coef, process_image=_defs.get_processor(channels=(3,), inbayer='01'), imread=lambda f: x4)
    np.savez_compressed(HERE / 'defs_golden.npz', **g)
    print({k: v.shape for k, v in g.items()})


if __name__ == '__main__':
    main()
