"""ORACLE (test infrastructure): numpy restatement of the reference's WS estimator and linear predictors.

Each function cites the reference lines it follows. Two flavours of the 3x3 predictor are kept:
`filter_predict_fft` goes through scipy.signal.convolve on x/255 exactly like the reference (FFT rounding noise
~1e-4 px), `filter_predict_exact` is the direct float64 stencil used to separate our error from
the reference's.
"""
import numpy as np
import scipy.signal

# src/ws/estimate.py:31-52 == src/filters/evaluate.py:29-50
NAMED_FILTERS = {
    'KB': np.array([[[-1, +2, -1], [+2, 0, +2], [-1, +2, -1]]], dtype='float32').T / 4.,
    'AVG': np.array([[[1, 1, 1], [1, 0, 1], [1, 1, 1]]], dtype='float32').T / 8.,
    'AVG9': np.array([[[1, 1, 1], [1, 1, 1], [1, 1, 1]]], dtype='float32').T / 9.,
    '1': np.array([[[0, 0, 0], [0, 1, 0], [0, 0, 0]]], dtype='float32').T / 1.,
}


def lsb_flip(x_u8: np.ndarray) -> np.ndarray:
    """x_bar = x ^ 1 on the integer pixels (src/ws/estimate.py:83, src/unet/evaluate.py:128)."""
    return np.asarray(x_u8, dtype=np.uint8) ^ 1


def filter_predict_fft(x: np.ndarray, name: str) -> np.ndarray:
    """src/filters/evaluate.py:136-141 infere_single: (H,W,C) float32 px -> (H-2,W-2,1) float32 px."""
    model = NAMED_FILTERS[name]
    y = scipy.signal.convolve(x / 255., model[..., ::-1], mode='valid')[..., :1]
    return y * 255.


def filter_predict_exact(x: np.ndarray, name: str) -> np.ndarray:
    """Same predictor as a direct float64 cross-correlation (all shipped kernels are symmetric)."""
    k = NAMED_FILTERS[name][..., 0].astype(np.float64)
    x0 = np.asarray(x, dtype=np.float64)[..., 0]
    H, W = x0.shape
    y = np.zeros((H - 2, W - 2))
    for dy in range(3):
        for dx in range(3):
            y += k[dy, dx] * x0[dy:H - 2 + dy, dx:W - 2 + dx]
    return y[..., None]


def local_variance(x: np.ndarray, exact: bool = True) -> np.ndarray:
    """src/ws/estimate.py:94-96: var = AVG(x^2) - AVG(x)^2 over the 8 neighbours."""
    if exact:
        mu = filter_predict_exact(x, 'AVG')
        mu2 = filter_predict_exact(np.asarray(x, dtype=np.float64) ** 2, 'AVG')
    else:
        avg = NAMED_FILTERS['AVG']
        mu = scipy.signal.convolve(x[..., :1], avg[..., ::-1], mode='valid')
        mu2 = scipy.signal.convolve(x[..., :1] ** 2, avg[..., ::-1], mode='valid')
    return mu2 - mu ** 2


def attack(x_u8: np.ndarray, pixel_estimator, correct_bias: bool = False, weighted: int = 1, exact: bool = True,
           clip: bool = True):
    """src/ws/estimate.py:80-128 on an in-memory uint8 (H,W,1) image. `pixel_estimator` is a callable
    (H,W,1) float -> (H-2,W-2,1) or a filter name. exact=True evaluates in float64 with direct stencils;
    exact=False follows the reference's dtypes and FFT convolutions."""
    x_u8 = np.asarray(x_u8, dtype=np.uint8)
    if x_u8.ndim == 2:
        x_u8 = x_u8[..., None]
    ft = np.float64 if exact else np.float32
    x_bar = lsb_flip(x_u8).astype(ft)                                   # :83,:87
    x = x_u8.astype(ft)                                                 # :86
    if isinstance(pixel_estimator, str):
        name = pixel_estimator
        pixel_estimator = (lambda v: filter_predict_exact(v, name)) if exact else (lambda v: filter_predict_fft(v, name))
    x1_hat = np.asarray(pixel_estimator(x), dtype=ft)                   # :90
    if abs(int(weighted)) == 1:                                         # :93-106
        var = local_variance(x, exact=exact)
        weights = 1 / (5 + var) if int(weighted) == 1 else 5 + var
        weights = weights / np.sum(weights)
    else:                                                               # :109-110
        weights = np.ones_like(x1_hat) / x1_hat.size
    x1, x1_bar = x[1:-1, 1:-1, :1], x_bar[1:-1, 1:-1, :1]               # :113-114
    beta_hat = np.sum(weights * (x1 - x1_bar) * (x1 - x1_hat))          # :118-120
    if clip:
        beta_hat = np.clip(beta_hat, 0, None)                           # :121
    if correct_bias:                                                    # :126-128
        x_bias = np.asarray(pixel_estimator(x_bar - x), dtype=ft)
        beta_hat = beta_hat - beta_hat * np.sum(weights * (x1 - x1_bar) * x_bias)
    return ft(beta_hat)


def predict_unet(x_u8: np.ndarray, x_hat: np.ndarray, exact: bool = True):
    """src/unet/evaluate.py:125-132: beta_hat = mean((x - x_bar)(x - x_hat)) (unclipped), l1 = mean|x - x_hat|
    over the interior; x_hat is the (H-2,W-2[,1]) prediction in pixel units."""
    ft = np.float64 if exact else np.float32
    x = np.asarray(x_u8, dtype=np.uint8)
    x = x[..., 0] if x.ndim == 3 else x
    xi = x[1:-1, 1:-1]
    x_bar = (xi ^ 1).astype(ft)
    xf = xi.astype(ft)
    xh = np.asarray(x_hat, dtype=ft).reshape(xf.shape)
    return ft(np.mean((xf - x_bar) * (xf - xh))), ft(np.mean(np.abs(xf - xh)))


def wsloss_betas(outputs01: np.ndarray, inputs01: np.ndarray, crop: int = 0, exact: bool = True):
    """Batched beta_hat of src/_defs/losses.py:46-61 (crop=0, relu) and src/_defs/metrics.py:122-137 (crop=1, clip):
    images (B,1,H,W) in [0,1]."""
    ft = np.float64 if exact else np.float32
    x = np.asarray(inputs01, dtype=ft) * 255.
    xh = np.asarray(outputs01, dtype=ft) * 255.
    if crop:
        x, xh = x[:, :, 1:-1, 1:-1], xh[:, :, 1:-1, 1:-1]
    x_bar = (np.round(x).astype('int64') ^ 1).astype(ft)
    w = 1.0 / np.prod(x.shape[1:])
    return np.clip(np.sum(w * (x - x_bar) * (x - xh), axis=(1, 2, 3)), 0, None)
