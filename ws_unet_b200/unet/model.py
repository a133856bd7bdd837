"""B200-native drop-in for the reference's UNet pixel predictor.

Mirrors the interface of src/unet/model/unet.py:54-199 (class UNet) and src/unet/model/__init__.py:8-49
(get_model): same constructor arguments, same sub-module names and construction order (so
`torch.manual_seed(s); get_model(...)` draws the same initial weights and `state_dict()` /
`load_state_dict()` are interchangeable with the reference), `to()` returning self,
`disable_center_pixels()`, `input_dropout` attribute. `forward` does not run PyTorch layers: it hands the
weights and the input to libwsunet (tcgen05 implicit-GEMM convolutions) through the C ABI.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from .. import _native


PRECISIONS = {'bf16x3': 0, 'fp16x2': 1, 'fp16x1': 2, 'fp16x1_f8': 3}   # wsu_set_option(h, "precision", .)


class UniformDropout(nn.Module):
    """Input dropout of the reference (unet.py:15-51): every pixel is kept with probability 1-drop_rate and otherwise
    replaced by its KB prediction (reflect padding). With drop_rate=0 (every shipped inference path,
    src/unet/evaluate.py:175-181) the keep-probability is 1 and the layer is the identity, so it is elided.
    For drop_rate>0 (training-time feature) the KB prediction and the blend run in libwsunet
    (`wsu_uniform_dropout`). Differences from the reference, both deliberate: the caller's tensor is NOT modified in place, and the mask
    comes from the device RNG (the reference draws it on the CPU), so masks are not reproducible across the two."""

    def __init__(self, p: float, drop_channel):
        super().__init__()
        self.p = 1 - p
        self.kb = torch.tensor([[[[-1, +2, -1], [+2, +0, +2], [-1, +2, -1]]]], dtype=torch.float32) / 4.
        self.drop_channel = drop_channel
        self.mask = None

    def forward(self, x):
        if self.p == 1:
            return x
        if not x.is_cuda:
            raise RuntimeError("ws_unet_b200 runs on a CUDA device only (no CPU fallback)")
        c = [int(v) for v in self.drop_channel]
        x = x.contiguous() if x.dtype == torch.uint8 else x.to(torch.float32).contiguous()
        B, C, H, W = x.shape
        # the keep-mask is drawn by torch's device RNG; the KB prediction and the blend run in libwsunet (wsu_uniform_dropout)
        self.mask = torch.empty(B, 1, H, W, device=x.device).bernoulli_(p=self.p)
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
        bits = 0
        for ch in c:
            bits |= 1 << ch
        with torch.cuda.device(x.device):
            _native.check(_native.load().wsu_uniform_dropout(
                x.device.index, ctypes.c_void_p(x.data_ptr()), _native.WSU_U8 if x.dtype == torch.uint8 else _native.WSU_F32,
                ctypes.c_void_p(self.mask.data_ptr()), ctypes.c_void_p(out.data_ptr()), B, C, H, W, bits,
                _native.stream_ptr(x.device)), 'wsu_uniform_dropout')
        return out


class UNet(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, nsteps: int, drop_rate: float, drop_channel):
        super().__init__()
        assert nsteps >= 0
        if nsteps > 4:
            raise NotImplementedError("nsteps > 4")
        self.nsteps = nsteps
        self.in_channels = in_channels
        self.out_channels = out_channels
        conv_kw = {'kernel_size': 3, 'padding': 1, 'padding_mode': 'reflect'}
        ups_kw = {'kernel_size': 2, 'stride': 2}
        # same order as unet.py:78-135 so parameter initialisation consumes the RNG identically
        if drop_rate is not None:
            self.input_dropout = UniformDropout(p=drop_rate, drop_channel=drop_channel)
        else:
            self.input_dropout = None
        self.e11 = nn.Conv2d(in_channels, 64, **conv_kw)
        self.e12 = nn.Conv2d(64, 64, **conv_kw)
        if nsteps >= 1:
            self.e21 = nn.Conv2d(64, 128, **conv_kw)
            self.e22 = nn.Conv2d(128, 128, **conv_kw)
        if nsteps >= 2:
            self.e31 = nn.Conv2d(128, 256, **conv_kw)
            self.e32 = nn.Conv2d(256, 256, **conv_kw)
        if nsteps >= 3:
            self.e41 = nn.Conv2d(256, 512, **conv_kw)
            self.e42 = nn.Conv2d(512, 512, **conv_kw)
        if nsteps >= 4:
            self.e51 = nn.Conv2d(512, 1024, **conv_kw)
            self.e52 = nn.Conv2d(1024, 1024, **conv_kw)
        if nsteps >= 4:
            self.upconv1 = nn.ConvTranspose2d(1024, 512, **ups_kw)
            self.d11 = nn.Conv2d(1024, 512, **conv_kw)
            self.d12 = nn.Conv2d(512, 512, **conv_kw)
        if nsteps >= 3:
            self.upconv2 = nn.ConvTranspose2d(512, 256, **ups_kw)
            self.d21 = nn.Conv2d(512, 256, **conv_kw)
            self.d22 = nn.Conv2d(256, 256, **conv_kw)
        if nsteps >= 2:
            self.upconv3 = nn.ConvTranspose2d(256, 128, **ups_kw)
            self.d31 = nn.Conv2d(256, 128, **conv_kw)
            self.d32 = nn.Conv2d(128, 128, **conv_kw)
        if nsteps >= 1:
            self.upconv4 = nn.ConvTranspose2d(128, 64, **ups_kw)
            self.d41 = nn.Conv2d(128, 64, **conv_kw)
            self.d42 = nn.Conv2d(64, 64, **conv_kw)
        self.outconv = nn.Conv2d(64, out_channels, kernel_size=1, padding_mode='reflect')
        self._handle = None       # wsu_handle (ctypes.c_void_p)
        self._handle_device = None
        self._weights_key = None
        self._precision = 'bf16x3'

    # ------------------------------------------------------------------ native handle management
    def __getstate__(self):  # joblib workers pickle the model (src/ws/estimate.py:139-146)
        state = self.__dict__.copy()
        state['_handle'] = None
        state['_handle_device'] = None
        state['_weights_key'] = None
        return state   # _precision travels with the pickle

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _release(self):
        if getattr(self, '_handle', None) is not None:
            _native.load().wsu_destroy(self._handle)
            self._handle = None

    def _key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def native_handle(self, device: torch.device):
        """Create (once per device) the libwsunet context and keep its packed weights in sync with the
        module's parameters (load_state_dict / in-place edits bump tensor versions)."""
        lib = _native.load()
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is None or self._handle_device != dev_index:
            self._release()
            h = ctypes.c_void_p()
            _native.check(lib.wsu_create(ctypes.byref(h), dev_index, self.nsteps, self.in_channels, self.out_channels),
                          'wsu_create')
            self._handle, self._handle_device, self._weights_key = h, dev_index, None
            _native.check(lib.wsu_set_option(h, b'precision', PRECISIONS[getattr(self, '_precision', 'bf16x3')]), 'precision')
        key = self._key()
        if key != self._weights_key:
            for name, t in self.state_dict().items():
                host = t.detach().to('cpu', torch.float32).contiguous()
                dims = (ctypes.c_int64 * host.dim())(*host.shape)
                _native.check(lib.wsu_load_weights(self._handle, name.encode(), ctypes.c_void_p(host.data_ptr()), dims,
                                                   host.dim()), f'wsu_load_weights({name})')
            _native.check(lib.wsu_commit_weights(self._handle), 'wsu_commit_weights')
            self._weights_key = key
        return self._handle

    # ------------------------------------------------------------------ reference interface
    def forward(self, x_in: torch.Tensor) -> torch.Tensor:
        if self.input_dropout is not None:
            x_in = self.input_dropout(x_in)
        if not x_in.is_cuda:
            raise RuntimeError("ws_unet_b200.UNet runs on a B200 only: move the input to a CUDA device "
                               "(there is no CPU fallback)")
        if x_in.dim() != 4 or x_in.shape[1] != self.in_channels:
            raise RuntimeError(f"expected input (B,{self.in_channels},H,W), got {tuple(x_in.shape)}")
        if x_in.dtype == torch.uint8:
            dtype = _native.WSU_U8
        else:
            dtype = _native.WSU_F32
            x_in = x_in.to(torch.float32)
        x_in = x_in.contiguous()
        B, _, H, W = x_in.shape
        if H % (1 << self.nsteps) or W % (1 << self.nsteps):
            # the reference fails inside torch.cat with a RuntimeError (unet.py:178)
            raise RuntimeError(f"Sizes of tensors must match: H={H}, W={W} not divisible by {1 << self.nsteps}")
        h = self.native_handle(x_in.device)
        y = torch.empty((B, 1, H, W), dtype=torch.float32, device=x_in.device)
        with torch.cuda.device(x_in.device):
            _native.check(_native.load().wsu_unet_forward(h, ctypes.c_void_p(x_in.data_ptr()), dtype,
                                                          ctypes.c_void_p(y.data_ptr()), B, H, W,
                                                          _native.stream_ptr(x_in.device)), 'wsu_unet_forward')
        return y

    def to(self, *args, **kw):
        super().to(*args, **kw)
        if self.input_dropout is not None:
            self.input_dropout.to(*args, **kw)
        return self

    def disable_center_pixels(self):
        self.e11.weight.data[:, :, 1, 1] = 0.
        if self.e11.weight.grad is not None:
            self.e11.weight.grad[:, :, 1, 1] = 0.
        self.refresh_weights()

    def refresh_weights(self):
        """Force a re-pack of the device weights on the next forward. Needed only after edits through `.data`
        (which do not bump tensor version counters); load_state_dict, optimizer steps and .to() are detected."""
        self._weights_key = None

    # ------------------------------------------------------------------ precision plan (no counterpart in the reference)
    def set_precision(self, mode: str, device=None):
        """Arithmetic of the layers whose input lives at UNet level >= 1 (e21.., d3x, up-convolutions):
        'bf16x3' (default) - split-bf16 operands, three MMAs per MAC, like the full-resolution layers e12 / d41 / d42;
        'fp16x2' / 'fp16x1' - ONE fp16 activation plane against fp16 (hi, lo) / fp16 weights, two / one MMA per MAC and
        half the activation bytes. Whether a reduced plan keeps the prediction inside the 1e-3 px bar depends on how much
        of the output the deep path carries for the weights at hand - use `calibrate_precision` to decide.
        'fp16x1_f8' - 'fp16x1', and the full-resolution layers (e12, d41's skip half, d42) compute a*w as one fp16 MMA plus
        ONE e4m3 MMA over both correction terms ((a - fp16 a) * w and a * (w - fp16 w), 2^-11 of the main term, so fp8's 4
        bits suffice): two MMA times instead of three at ~15 significant bits, independent of the weights."""
        if mode not in PRECISIONS:
            raise ValueError(f'precision must be one of {list(PRECISIONS)}')
        self._precision = mode
        if self._handle is not None:
            _native.check(_native.load().wsu_set_option(self._handle, b'precision', PRECISIONS[mode]), 'precision')
        return self

    def active_precision(self, device=None) -> str:
        """The plan the library actually runs (reduced plans need shared-memory-resident up-convolutions: unet_1, unet_2)."""
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        out = ctypes.c_int64()
        _native.check(_native.load().wsu_get_info(self.native_handle(dev), b'precision', ctypes.byref(out)))
        return {v: k for k, v in PRECISIONS.items()}[int(out.value)]

    @torch.no_grad()
    def calibrate_precision(self, images: torch.Tensor, budget_px: float = 2e-4,
                            modes=('fp16x1_f8', 'fp16x1', 'fp16x2')) -> dict:
        """Pick the cheapest precision plan whose predictions on `images` ((B,C,H,W) uint8 pixels or float32 in [0,1], on the
        device) stay within `budget_px` (max-abs, pixel units) of the three-term plan, which itself sits 2-5e-5 px from the
        FP32 reference. Leaves the chosen plan active and returns {'chosen', 'max_abs_px': {mode: err}, 'budget_px'}."""
        self.set_precision('bf16x3')
        ref = self(images)
        report = {'chosen': 'bf16x3', 'max_abs_px': {}, 'budget_px': budget_px, 'images': int(images.shape[0])}
        for mode in modes:
            self.set_precision(mode)
            if self.active_precision(images.device) != mode:
                report['max_abs_px'][mode] = None      # not available for this depth
                continue
            err = ((self(images) - ref).abs().max() * 255.).item()
            report['max_abs_px'][mode] = err
            if err == err and err <= budget_px:
                report['chosen'] = mode
                return report
        self.set_precision('bf16x3')
        return report

    def set_micro_batch(self, n: int, device=None):
        """Images per pass through the layer chain (0 = auto from free HBM budget)."""
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        _native.check(_native.load().wsu_set_option(self.native_handle(dev), b'micro_batch', int(n)))


def get_model(name: str, in_channels: int, out_channels: int = 1, channel=[0], drop_rate: float = 0.) -> nn.Module:
    """src/unet/model/__init__.py:8-49."""
    if name.lower().startswith('unet'):
        nsteps = int(name.split('_')[1])
        return UNet(in_channels=in_channels, out_channels=out_channels, nsteps=nsteps, drop_channel=channel,
                    drop_rate=drop_rate)
    raise NotImplementedError(name)
