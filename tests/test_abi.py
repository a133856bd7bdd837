"""CPU: the C-ABI library loads, exports every symbol include/wsunet.h declares, and fails loudly without a GPU."""
import ctypes
import pathlib
import re

import pytest
import torch

from ws_unet_b200 import _native

REPO = pathlib.Path(__file__).resolve().parents[1]


def header_symbols():
    text = (REPO / 'include' / 'wsunet.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(wsu_[a-z_0-9]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/wsunet.h but not exported'
    assert sorted(_native.PROTOTYPES) == syms, 'ctypes prototype table out of sync with the header'
    assert lib.wsu_version() >= 100


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', str(_native.LIB_PATH)], capture_output=True, text=True).stdout
    assert 'sm_100a' in out and not re.search(r'sm_(?!100a)\d+', out)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback_without_gpu():
    lib = _native.load()
    h = ctypes.c_void_p()
    rc = lib.wsu_create(ctypes.byref(h), 0, 2, 1, 1)
    assert rc == _native.WSU_ERR_CUDA
    assert 'no CUDA device' in _native.last_error()
    import ws_unet_b200 as W
    model = W.get_model('unet_2', 1)
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 1, 16, 16))
    with pytest.raises(RuntimeError):
        W.ws_estimate(torch.zeros(1, 1, 16, 16, dtype=torch.uint8), 'KB')
    with pytest.raises(RuntimeError):
        W.filters.get_filter_estimator('KB')(torch.zeros(8, 8, 1).numpy())


def test_argument_validation_needs_no_gpu():
    lib = _native.load()
    h = ctypes.c_void_p()
    assert lib.wsu_create(ctypes.byref(h), 0, 7, 1, 1) == _native.WSU_ERR_INVALID
    assert lib.wsu_create(ctypes.byref(h), 0, 2, 1, 3) == _native.WSU_ERR_INVALID
    assert lib.wsu_filter_predict(0, None, 0, 0, None, 1, 8, 8, None) == _native.WSU_ERR_INVALID
    with pytest.raises(ValueError):
        _native.check(_native.WSU_ERR_INVALID)
    with pytest.raises(RuntimeError):
        _native.check(_native.WSU_ERR_CUDA)
