"""ctypes binding of libwsunet.so (C ABI declared in include/wsunet.h).

The library is the product: there is no Python/PyTorch fallback for any compute entry point. If the shared
object is missing this module raises at import of the symbol table, and every compute call raises when no
CUDA device is present.
"""
from __future__ import annotations

import ctypes
import os
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = _HERE / "_lib" / "libwsunet.so"

WSU_OK, WSU_ERR_INVALID, WSU_ERR_CUDA, WSU_ERR_STATE = 0, -1, -2, -3
WSU_U8, WSU_F32 = 0, 1
PRED_KINDS = {"KB": 0, "AVG": 1, "AVG9": 2, "1": 3}

_c = ctypes
_vp, _i, _i64, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_size_t

# name -> (restype, argtypes); must list every symbol include/wsunet.h declares (tests/test_abi.py checks this)
PROTOTYPES = {
    "wsu_last_error": (_c.c_char_p, []),
    "wsu_version": (_i, []),
    "wsu_create": (_i, [_c.POINTER(_vp), _i, _i, _i, _i]),
    "wsu_destroy": (_i, [_vp]),
    "wsu_load_weights": (_i, [_vp, _c.c_char_p, _vp, _c.POINTER(_i64), _i]),
    "wsu_commit_weights": (_i, [_vp]),
    "wsu_set_option": (_i, [_vp, _c.c_char_p, _i64]),
    "wsu_unet_forward": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "wsu_unet_ws_estimate": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "wsu_unet_ws_estimate_host": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "wsu_filter_predict": (_i, [_i, _vp, _i, _i, _vp, _i, _i, _i, _vp]),
    "wsu_filter_ws_estimate": (_i, [_i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp]),
    "wsu_filter_ws_estimate_host": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i]),
    "wsu_ws_from_prediction": (_i, [_i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp]),
    "wsu_ws_grad_prediction": (_i, [_i, _vp, _i, _vp, _i, _c.c_float, _vp, _i, _i, _i, _vp]),
    "wsu_uniform_dropout": (_i, [_i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _c.c_uint32, _vp]),
    "wsu_filter_residual_rows": (_i, [_i, _vp, _i, _vp, _vp, _i64, _vp]),
    "wsu_debug_layer": (_i, [_vp, _c.c_char_p, _vp, _sz, _i, _c.POINTER(_i64), _vp]),
    "wsu_get_info": (_i, [_vp, _c.c_char_p, _c.POINTER(_i64)]),
    "wsu_profile_read": (_i, [_vp, _vp, _i]),
    "wsu_profile_name": (_c.c_char_p, [_vp, _i]),
    "wsu_launch_count": (_i64, [_i]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libwsunet.so (built by __graft_entry__.build() / `make -C ws_unet_b200/csrc`)."""
    global _lib
    if _lib is None:
        path = pathlib.Path(os.environ.get("WSUNET_LIB", LIB_PATH))
        if not path.exists():
            raise ImportError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(libwsunet has no Python fallback)")
        lib = ctypes.CDLL(str(path))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().wsu_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    """Map wsu_status to the exception types the reference's own code paths raise."""
    if rc == WSU_OK:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc == WSU_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)


def stream_ptr(device=None):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
