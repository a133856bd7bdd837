"""Consumers of the beta_hat vector and the batched training-side definitions, mirrored from the reference:
  * WSLoss / L1WSLoss / WSMeter - src/_defs/losses.py:27-115, src/_defs/metrics.py:116-142; the losses are differentiable
    with respect to the predictor output (beta_hat is linear in it: wsu_ws_grad_prediction), so they can drive an
    autograd graph that ends in a torch predictor
  * produce_roc       - src/ws/roc.py:198-283 (501-threshold ROC, AUC, P_E from beta_hat)
The WS arithmetic runs through libwsunet (wsu_ws_from_prediction); ROC is a few thousand scalar ops on the host.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

import ctypes

from . import _native, ws


class _WSBetas(torch.autograd.Function):
    """beta_hat (unclipped) of a batch as a differentiable function of the predictor output in [0,1]."""

    @staticmethod
    def forward(ctx, outputs, inputs, crop):
        ctx.save_for_backward(inputs)
        ctx.crop = int(crop)
        ctx.out_shape = outputs.shape
        return ws.ws_from_prediction(inputs, outputs.detach().reshape(outputs.shape[0], *outputs.shape[-2:]) * 255.,
                                     weighted=0, clip=False, crop=int(crop))

    @staticmethod
    def backward(ctx, grad_beta):
        (inputs,) = ctx.saved_tensors
        images, dtype = ws._prep_images(inputs)
        B, _, H, W = images.shape
        dev = images.device
        coef = grad_beta.to(torch.float32).contiguous()
        grad = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.check(_native.load().wsu_ws_grad_prediction(
                dev.index, ctypes.c_void_p(images.data_ptr()), dtype, ctypes.c_void_p(coef.data_ptr()), ctx.crop, 255.,
                ctypes.c_void_p(grad.data_ptr()), B, H, W, _native.stream_ptr(dev)), 'wsu_ws_grad_prediction')
        return grad.reshape(ctx.out_shape), None, None


def ws_betas_hat(outputs: torch.Tensor, inputs: torch.Tensor, crop: int = 0) -> torch.Tensor:
    """betas_hat of WSLoss._error (crop=0, relu; losses.py:46-61) or WSMeter.update (crop=1, clip; metrics.py:122-137):
    images (B,1,H,W) float32 in [0,1] on a CUDA device; outputs = predictor output in [0,1]. Differentiable in outputs."""
    return torch.relu(_WSBetas.apply(outputs, inputs, crop))


class L1Loss:
    """src/_defs/losses.py:27-36: mean |cover - output|."""

    def __call__(self, outputs, targets, *args, **kw):
        covers, _ = targets
        return torch.mean(torch.abs(covers - outputs))


class WSLoss:
    """src/_defs/losses.py:45-89: mean |relu(beta_hat) - alpha/2| over the batch (whole image, uniform weights)."""

    def _error(self, outputs, inputs, betas):
        return torch.abs(ws_betas_hat(outputs, inputs, crop=0) - betas.to(outputs.device))

    def __call__(self, outputs, targets, inputs):
        _, alphas = targets
        return torch.mean(self._error(outputs, inputs, alphas / 2.))


class L1WSLoss:
    """src/_defs/losses.py:92-115: prediction MAE + WS MAE."""

    def __init__(self):
        self.l1_loss, self.ws_loss = L1Loss(), WSLoss()

    def __call__(self, outputs, targets, inputs):
        return self.l1_loss(outputs, targets) + self.ws_loss(outputs, targets, inputs)


class WSMeter:
    """src/_defs/metrics.py:116-142 (AverageMeter over per-batch mean |beta_hat - alpha/2|, 1-px crop)."""
    name = 'ws'

    def __init__(self):
        self.sum, self.count = 0.0, 0

    def update(self, x, x_hat, alphas):
        x = torch.as_tensor(x)
        x_hat = torch.as_tensor(x_hat)
        if not x.is_cuda:
            x, x_hat = x.cuda(), x_hat.cuda()
        betas_hat = ws_betas_hat(x_hat, x, crop=1).cpu().numpy()
        self.sum += float(np.mean(np.abs(betas_hat - np.asarray(alphas) / 2.)))
        self.count += 1

    @property
    def avg(self):
        return self.sum / max(self.count, 1)


def produce_roc(df_ws: pd.DataFrame) -> pd.DataFrame:
    """src/ws/roc.py:198-283: per (stego_method, model_name) against the 'Cover' rows, thresholds tau in
    linspace(0,1,501) on clip(beta_hat, 0) (model names containing 'B0' use the detector's `score` column against alpha,
    roc.py:211-213), AUC by the reference's FPR-bin weighting, P_E = min (1-TPR+FPR)/2.

    `tpr_50` reproduces the reference as written: roc.py:251 divides TP(tau=.5) by TP(tau=.5) + FN where FN is the value
    left over from the LAST threshold of the loop (tau = 0), i.e. the positives with beta_hat <= 0, not the positives with
    beta_hat <= .5. Kept so that a results table regenerated through this package matches the reference's column."""
    out = []
    for (stego_method, model_name), _ in df_ws.groupby(['stego_method', 'model_name']):
        if stego_method == 'Cover':
            continue
        d = df_ws[(df_ws['model_name'] == model_name) & df_ws['stego_method'].isin([stego_method, 'Cover'])]
        if 'B0' in model_name:
            y_hat = d['score'].to_numpy()
            y = d['alpha'].to_numpy()
        else:
            y_hat = np.clip(d['beta_hat'].to_numpy(), 0, None)
            y = d['alpha'].to_numpy() / 2
        taus = np.array(list(reversed(np.linspace(0, 1, 501, endpoint=True))))
        gt = y_hat[None, :] > taus[:, None]
        pos, neg = (y > 0.)[None, :], (y <= 0.)[None, :]
        TP, FP = (gt & pos).sum(1), (gt & neg).sum(1)
        TN, FN = (~gt & neg).sum(1), (~gt & pos).sum(1)
        with np.errstate(divide='ignore', invalid='ignore'):
            tpr, fpr = TP / (TP + FN), FP / (FP + TN)
            bins = np.diff(fpr, prepend=fpr[0])
            bins = bins / bins.sum()
            auc = np.sum(bins * tpr)
            err = (1 - tpr + fpr) / 2
            i0 = int(np.argmin(err))
            g50 = y_hat > .5
            TP50, FP50, TN50 = (g50 & (y > 0.)).sum(), (g50 & (y <= 0.)).sum(), (~g50 & (y <= 0.)).sum()
            fpr50 = FP50 / (FP50 + TN50)
            tpr50 = TP50 / (TP50 + FN[-1])          # roc.py:251: FN of the loop's last threshold (tau = 0)
        out.append(pd.DataFrame({
            'stego_method': stego_method, 'model_name': model_name, 'tau': taus, 'tpr': tpr, 'fpr': fpr, 'p_e': err[i0],
            'tau0': taus[i0], 'fpr_tau0': fpr[i0], 'tpr_tau0': tpr[i0], 'auc': auc, 'fpr_50': fpr50, 'tpr_50': tpr50,
            'label': model_name if 'B0' in model_name else f'WS-{model_name}',
        }))
    return pd.concat(out)
