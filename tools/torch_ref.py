"""Plain PyTorch fp32 restatement of UNet.forward (src/unet/model/unet.py:137-189) over the parameters of a
ws_unet_b200 UNet module. Test/debug infrastructure only (tests/, tools/): the product path never calls it."""
import torch
import torch.nn.functional as F


def _conv(m, x):
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), m.weight, m.bias)


def reference_forward(model, x, keep=False):
    """x: (B,C,H,W) float32 in [0,1] on any device. Returns sigmoid output and (optionally) every feature map
    keyed like wsu_debug_layer names."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n = model.nsteps
    acts = {}
    enc = []
    h = x
    for l in range(n + 1):
        a = F.relu(_conv(getattr(model, f'e{l + 1}1'), h))
        b = F.relu(_conv(getattr(model, f'e{l + 1}2'), a))
        acts[f'e{l + 1}1'], acts[f'e{l + 1}2'] = a, b
        enc.append(b)
        if l < n:
            h = F.max_pool2d(b, 2, 2)
            acts[f'p{l + 1}'] = h
    h = enc[-1]
    for l in range(n - 1, -1, -1):
        k = 4 - l
        up = getattr(model, f'upconv{k}')
        u = F.conv_transpose2d(h, up.weight, up.bias, stride=2)
        acts[f'u{k}'] = u
        a = F.relu(_conv(getattr(model, f'd{k}1'), torch.cat([u, enc[l]], dim=1)))
        h = F.relu(_conv(getattr(model, f'd{k}2'), a))
        acts[f'd{k}1'], acts[f'd{k}2'] = a, h
    y = torch.sigmoid(F.conv2d(h, model.outconv.weight, model.outconv.bias))
    return (y, acts) if keep else y
