"""UNet pixel predictor package - mirrors src/unet/__init__.py:9-11,110-121 of the reference."""
from .model import UNet, UniformDropout, get_model
from .evaluate import infere_single, predict_unet, get_pretrained


def get_unet_estimator(*args, **kw):
    """src/unet/__init__.py:110-121: closure (H,W,C) float32 pixels -> (510,510,1) float32 pixels."""
    model = get_pretrained(*args, **kw)

    def predict(x):
        return infere_single(x, model=model)

    return predict


def make_unet_estimator(model):
    """Same closure for an already constructed model (random-init parity tests)."""
    return lambda x: infere_single(x, model=model)
