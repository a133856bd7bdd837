"""ORACLE / test infrastructure: import the UNMODIFIED reference package tree (src/ of uibk-uncover/ws-unet).

`__graft_entry__.build()` mirrors /root/reference/src into the git-ignored baseline/_ref/src when the mount is
present (the reference has no setup.py / pyproject, so there is nothing for pip to install; the copy travels to the
GPU box with the snapshot, /root/reference does not). Modules the hot path never touches but the reference imports at
module level (timm, conseal, jpeglib, seaborn, matplotlib, torchinfo) are replaced by inert stubs (SURVEY.md 8c recipe).
Only tests, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may call this.
"""
import importlib.machinery
import os
import pathlib
import sys
import types

REPO = pathlib.Path(__file__).resolve().parents[1]
VENDORED = REPO / 'baseline' / '_ref' / 'src'
MOUNT = pathlib.Path('/root/reference/src')
STUBS = ['timm', 'conseal', 'jpeglib', 'seaborn', 'matplotlib', 'matplotlib.pyplot', 'torchinfo']


def reference_src():
    """Directory of the reference's src/ tree, or None."""
    for p in (VENDORED, MOUNT):
        if (p / 'unet' / 'model' / 'unet.py').exists():
            return p
    return None


def import_reference(src=None):
    """Returns the reference's (_defs, filters, unet, ws) packages, imported from `src` (default: reference_src())."""
    src = pathlib.Path(src) if src is not None else reference_src()
    if src is None:
        raise ImportError('reference sources not found (baseline/_ref/src is created by __graft_entry__.build())')
    for name in STUBS:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            m.__path__ = []
            sys.modules[name] = m
    cwd = os.getcwd()
    sys.path.insert(0, str(src))
    try:
        os.chdir(src)   # the reference appends 'unet', 'detector', '.' to sys.path relative to cwd = src/
        import _defs, filters, unet, ws  # noqa: E401
        import unet.model  # noqa: F401
        import unet.evaluate  # noqa: F401
        import ws.estimate  # noqa: F401
    finally:
        os.chdir(cwd)
    return _defs, filters, unet, ws
