"""Drop-in for the reference's neural_nets.py (Vgg19, :10-68) plus the additive StyleLoss / ContentLoss
modules named by the north star.  The VGG19 convolutions stay on torch/cuDNN (out of scope); what changes is
everything computed FROM the feature maps."""
from __future__ import annotations

from collections import namedtuple
from typing import Optional

import torch
from torchvision import models

from . import ops


class Vgg19(torch.nn.Module):
    """Same constructor, attributes, slices and forward() result as the reference (neural_nets.py:17-68).
    torchvision's ReLUs are in-place, so the tensor exposed as 'conv4_2' is mutated by slice6's leading ReLU
    and is really relu4_2; that aliasing is preserved because the slices are torchvision's own
    modules in the same order."""

    def __init__(self, requires_grad=False, show_progress=False, use_relu=True):
        super().__init__()
        vgg_pretrained_features = models.vgg19(pretrained=True, progress=show_progress).features
        if use_relu:
            self.layer_names = ['relu1_1', 'relu2_1', 'relu3_1', 'relu4_1', 'conv4_2', 'relu5_1']
            self.offset = 1
        else:
            self.layer_names = ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv4_2', 'conv5_1']
            self.offset = 0
        self.content_feature_maps_index = 4
        self.style_feature_maps_indices = list(range(len(self.layer_names)))
        self.style_feature_maps_indices.remove(4)
        bounds = [(0, 1 + self.offset), (1 + self.offset, 6 + self.offset), (6 + self.offset, 11 + self.offset),
                  (11 + self.offset, 20 + self.offset), (20 + self.offset, 22), (22, 29 + self.offset)]
        for n, (lo, hi) in enumerate(bounds, start=1):
            seq = torch.nn.Sequential()
            for x in range(lo, hi):
                seq.add_module(str(x), vgg_pretrained_features[x])
            setattr(self, f'slice{n}', seq)
        if not requires_grad:
            for param in self.parameters():
                param.requires_grad = False
        self._outputs = namedtuple('VggOutputs', self.layer_names)

    def forward(self, x):
        feats = []
        for n in range(1, 7):
            x = getattr(self, f'slice{n}')(x)
            feats.append(x)
        return self._outputs(*feats)


class StyleLoss(torch.nn.Module):
    """mean((A - G(x))^2) for one feature map, G = F F^T / (C*HW): the per-layer term the reference forms with
    gram_matrix + MSELoss (neural_style_transfer.py:100-104), as one fused kernel pair; backward (G - A) F."""

    def __init__(self, target_gram: torch.Tensor, precision: Optional[str] = None):
        super().__init__()
        t = target_gram.detach()
        if t.dim() == 3:
            t = t[0]
        self.register_buffer('target', t.contiguous())
        self.precision = precision
        self._ws = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        ch, hw = x.shape[-3], x.shape[-2] * x.shape[-1]
        need = ops.L.load().ast_gram_workspace_bytes(ch, hw)
        if self._ws is None or self._ws.nbytes < need or self._ws.buf.device != x.device:
            self._ws = ops.Workspace(need, x.device)
        return ops.StyleLossFn.apply(x, self.target, self._ws, ops._prec(self.precision))


class ContentLoss(torch.nn.Module):
    """mean((T - x)^2) (neural_style_transfer.py:95) as one streaming reduction; backward 2(x - T)/n."""

    def __init__(self, target: torch.Tensor):
        super().__init__()
        self.register_buffer('target', target.detach().contiguous())
        self._ws = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._ws is None or self._ws.buf.device != x.device:
            self._ws = ops.reduce_workspace(x.device)
        return ops.ContentLossFn.apply(x, self.target, self._ws)
