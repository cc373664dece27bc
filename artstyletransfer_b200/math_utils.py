"""Drop-in for the reference's math_utils.py: same names, arguments and error behaviour
(prepare_model :9-23, gram_matrix :26-34, total_variation :37-41, regularization :44-47), with the
arithmetic running in libast_sm100.so.  Additive: bicubic resize helpers named by the north star."""
from __future__ import annotations

import math
from functools import reduce

import torch

from . import ops
from .neural_nets import Vgg19


def prepare_model(model, device):
    """math_utils.py:9-23 — returns (net.eval() on device, content index, style indices)."""
    if model == 'vgg19':
        model = Vgg19(requires_grad=False, show_progress=True)
    else:
        raise ValueError(f'{model} not supported.')
    content_feature_maps_index = model.content_feature_maps_index
    style_feature_maps_indices = model.style_feature_maps_indices
    return model.to(device).eval(), content_feature_maps_index, style_feature_maps_indices


def gram_matrix(x, should_normalize=True, precision=None):
    """math_utils.py:26-34 — (b, ch, h, w) -> (b, ch, ch), G = F F^T / (ch*h*w).  Differentiable.
    Split-K tcgen05 kernel (TF32 operands rounded to nearest, fp32 accumulate) or the exact fp32 path
    (precision='fp32')."""
    return ops.gram_matrix(x, should_normalize, precision)


def total_variation(y):
    """math_utils.py:37-41 — mean|dx|^2 + mean|dy|^2 (squares of the two means).  Differentiable."""
    return ops.total_variation(y)


def regularization(y):
    """math_utils.py:44-47 — unused by the reference; kept for API parity (plain torch, not on the hot path)."""
    els = reduce(lambda a, b: a * b, y.shape)
    return torch.sum(torch.pow(y / 128.0, 10)) / math.pow(els, 10)


# ---- additive helpers (not in the reference; named by BASELINE.json's north_star) -------------------------
def bicubic_half(x):
    """F.interpolate(x, size=(H//2, W//2), mode='bicubic') with a deterministic adjoint
    (the in-loop pyramid step, neural_style_transfer.py:173-176)."""
    return ops.bicubic_half(x)


def bicubic_resize(x, out_h, out_w, layout='chw', coord='cv2'):
    """General-ratio Keys bicubic (cv2.resize INTER_CUBIC / F.interpolate semantics), no autograd."""
    return ops.bicubic_resize(x, out_h, out_w, layout, coord)
