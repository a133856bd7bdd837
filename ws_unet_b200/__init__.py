"""ws_unet_b200 - B200-native (sm_100a) implementation of the UNet -> Weighted-Stego hot path of
uibk-uncover/ws-unet behind the reference's own predictor / estimator interfaces."""
from . import _native  # noqa: F401
from . import filters, unet, ws  # noqa: F401
from .unet import get_model, UNet  # noqa: F401
from .ws import ws_estimate, ws_estimate_host, ws_from_prediction, attack  # noqa: F401
from . import dataset, defs, metrics, parallel  # noqa: F401

__all__ = ['filters', 'unet', 'ws', 'dataset', 'defs', 'metrics', 'parallel', 'get_model', 'UNet', 'ws_estimate', 'ws_estimate_host',
           'ws_from_prediction', 'attack']
