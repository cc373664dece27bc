"""Channels-last execution of the reference's VGG19 feature path and of one pyramid level's loss on top of it.

What the reference does per level and closure (neural_style_transfer.py:84-112 + autograd): Vgg19.forward
(neural_nets.py:53-68) through torch modules, five gram_matrix + MSELoss, content MSELoss, total_variation, and
`backward()` through all of it.  Here the same arithmetic runs as an explicit schedule with no autograd graph:

  forward   image (1,3,H,W) planar -> (H,W,3)            ast_chw_to_hwc
            13 x cuDNN conv + bias + ReLU (torch op,     aten.cudnn_convolution_relu on channels_last tensors: the TF32
                 out of scope)                           tensor-core kernels run without NCHW<->NHWC transposes and
                                                         apply the epilogue (sharded path: conv.out + ast_bias_relu_nhwc)
            4  x 2x2 max-pool                            ast_maxpool2x2_nhwc (no index tensor)
            5  x Gram + MSE on (HW, C) taps              ast_gram_mse_fwd_nhwc  (tcgen05, split-K, fused finalize)
            content MSE, TV, weighted sum                ast_mse_fwd, ast_tv_fwd, ast_level_combine
  backward  deepest tap first: dF (+)= s (G-A) F         ast_gram_bwd_nhwc (accumulates into the running gradient
                                                         with TMA reduce-add: no separate add kernels)
            content: dX += 2 w (X-T)/n                   ast_mse_bwd(accumulate)
            ReLU backward: fused into the tap kernels    relu_mask of ast_gram_bwd_nhwc / ast_mse_bwd (6 tap layers),
            above, into the pool backward, else in place ast_maxpool2x2_bwd_nhwc (4), ast_relu_bwd (3)
            conv backward-data (cuDNN, torch op)         aten.convolution_backward(output_mask = input only)
            (H,W,3) -> planar image gradient, += TV grad ast_hwc_to_chw, ast_tv_bwd(accumulate)

The taps are the reference's: relu1_1, relu2_1, relu3_1, relu4_1, "conv4_2" (post-ReLU because torchvision's
ReLUs are in place) and relu5_1.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch

from . import ops

_CL = torch.channels_last
# cuDNN's fused conv + bias + ReLU (torch.cudnn_convolution_relu) runs at the speed of the bare convolution on B200 and is
# bit-identical to conv followed by ast_bias_relu_nhwc (tests/tools/conv_fused_probe.py, profiles/r01_conv_fused.jsonl),
# so the unsharded path lets cuDNN apply the epilogue.  '0' = separate bias/ReLU kernel (what the sharded path uses:
# its convolutions write straight into padded band buffers through cudnn_convolution.out).
FUSED_CONV_RELU = os.environ.get('AST_FUSED_CONV_RELU', '1') != '0'
# On (the default): cuDNN times its engines once per convolution shape (torch.backends.cudnn.benchmark semantics,
# applied only around this path's own calls, the global flag is restored) instead of trusting its heuristics: the
# search costs ~1 s in the first closure of a job and buys 15 % of every later one at L=3 on a B200.  The reference
# leaves the flag at torch's default (off); AST_CUDNN_BENCHMARK=0 does the same.  bench.py measures the default.
CUDNN_BENCHMARK = os.environ.get('AST_CUDNN_BENCHMARK', '1') != '0'
# widest tap whose ReLU backward is fused into ast_gram_bwd_nhwc's epilogue (wider taps run ast_relu_bwd afterwards)
FUSED_TAP_RELU_MAX_C = int(os.environ.get('AST_FUSED_TAP_RELU_MAX_C', '256'))


class _cudnn_mode:
    """torch.backends.cudnn.benchmark = CUDNN_BENCHMARK for the duration of one cuDNN call of this path."""
    __slots__ = ('old',)

    def __enter__(self):
        self.old = torch.backends.cudnn.benchmark
        if CUDNN_BENCHMARK and not self.old:
            torch.backends.cudnn.benchmark = True

    def __exit__(self, *exc):
        if torch.backends.cudnn.benchmark != self.old:
            torch.backends.cudnn.benchmark = self.old


class FeaturePlan:
    """The frozen Vgg19's modules flattened into conv(+bias+ReLU) / pool steps with the tap positions."""

    def __init__(self, net):
        mods, ends = [], []
        for n in range(1, 7):
            sl = getattr(net, f'slice{n}')
            mods += list(sl)
            ends.append(len(mods) - 1)
        self.steps = []          # ('conv', weight_cl, bias, cin, cout) | ('pool',)
        step_of_module = {}
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, torch.nn.Conv2d):
                ok = (m.kernel_size == (3, 3) and m.stride == (1, 1) and m.padding == (1, 1) and m.dilation == (1, 1)
                      and m.groups == 1 and m.bias is not None and m.padding_mode == 'zeros')
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                if not ok or not (isinstance(nxt, torch.nn.ReLU) and nxt.inplace):
                    raise ValueError('feature path: expected 3x3/1/1 Conv2d followed by an in-place ReLU')
                if m.weight.requires_grad or m.bias.requires_grad:
                    raise ValueError('feature path: the network must be frozen (requires_grad=False)')
                w = m.weight.detach().contiguous(memory_format=_CL)
                self.steps.append(('conv', w, m.bias.detach().contiguous(), m.in_channels, m.out_channels))
                # a tap on the conv output aliases the in-place ReLU's output
                step_of_module[i] = step_of_module[i + 1] = len(self.steps) - 1
                i += 2
            elif isinstance(m, torch.nn.MaxPool2d):
                ks = m.kernel_size if isinstance(m.kernel_size, tuple) else (m.kernel_size, m.kernel_size)
                st = m.stride if isinstance(m.stride, tuple) else (m.stride, m.stride)
                if ks != (2, 2) or st != (2, 2) or m.padding not in (0, (0, 0)) or m.ceil_mode:
                    raise ValueError('feature path: expected 2x2/2 floor-mode MaxPool2d')
                self.steps.append(('pool',))
                step_of_module[i] = len(self.steps) - 1
                i += 1
            else:
                raise ValueError(f'feature path: unsupported module {type(m).__name__}')
        self.tap_step = [step_of_module[e] for e in ends]           # step whose output is tap k
        for k, e in enumerate(ends):
            if self.steps[self.tap_step[k]][0] != 'conv':
                raise ValueError('feature path: taps must be ReLU outputs')
        self.taps_at = {}
        for k, sidx in enumerate(self.tap_step):
            self.taps_at.setdefault(sidx, []).append(k)
        self.n_steps_needed = max(self.tap_step) + 1
        self.device = self.steps[0][1].device

    @staticmethod
    def try_build(net) -> Optional['FeaturePlan']:
        try:
            plan = FeaturePlan(net)
        except (ValueError, AttributeError, IndexError):
            return None
        return plan if plan.device.type == 'cuda' else None


def plan_for(net) -> Optional[FeaturePlan]:
    """Cached FeaturePlan of a network (None when it is not the frozen torchvision-style Vgg19 on CUDA)."""
    plan = getattr(net, '_ast_feature_plan', False)
    if plan is False:
        plan = FeaturePlan.try_build(net)
        try:
            object.__setattr__(net, '_ast_feature_plan', plan)
        except Exception:
            pass
    if plan is not None and plan.steps[0][1].device != next(net.parameters()).device:
        plan = FeaturePlan.try_build(net)
        object.__setattr__(net, '_ast_feature_plan', plan)
    return plan


def _ckey(kind, x, w):
    return (kind, int(w.shape[1]), int(w.shape[0]), int(x.shape[2]), int(x.shape[3]))


def _conv_fwd(x, w):
    with ops.timed(x.device, _ckey('cudnn_conv_fwd', x, w)):
        return torch.ops.aten.cudnn_convolution(x, w, [1, 1], [1, 1], [1, 1], 1, CUDNN_BENCHMARK, False,
                                                torch.backends.cudnn.allow_tf32)


def _conv_relu_fwd(x, w, b):
    with ops.timed(x.device, _ckey('cudnn_conv_bias_relu_fwd', x, w)), _cudnn_mode():
        y = torch.cudnn_convolution_relu(x, w, b, [1, 1], [1, 1], [1, 1], 1)
    return y if y.is_contiguous(memory_format=_CL) else y.contiguous(memory_format=_CL)


def _conv_bwd_data(g, x, w):
    with ops.timed(x.device, _ckey('cudnn_conv_dgrad', x, w)), _cudnn_mode():
        gi = torch.ops.aten.convolution_backward(g, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                 [True, False, False])[0]
    return gi if gi.is_contiguous(memory_format=_CL) else gi.contiguous(memory_format=_CL)


def features_forward(plan: FeaturePlan, img: torch.Tensor, keep: bool):
    """img: (1,3,H,W) planar fp32.  Returns (taps, saved): taps[k] is the channels_last (1,C,h,w) tap tensor;
    saved (when keep) holds per step (input, output) for the backward."""
    ops._require_cuda(img)
    if img.dim() != 4 or img.shape[0] != 1 or img.shape[1] != plan.steps[0][3]:
        raise ValueError(f'feature path expects (1, {plan.steps[0][3]}, H, W); got {tuple(img.shape)}')
    img = img.contiguous()
    _, c0, h, w = img.shape
    x = torch.empty((1, c0, h, w), dtype=torch.float32, device=img.device, memory_format=_CL)
    ops.chw_to_hwc(img, x, c0, h * w)
    taps: List[Optional[torch.Tensor]] = [None] * len(plan.tap_step)
    saved = []
    for sidx in range(plan.n_steps_needed):
        st = plan.steps[sidx]
        if st[0] == 'conv':
            if FUSED_CONV_RELU:
                y = _conv_relu_fwd(x, st[1], st[2])
            else:
                y = _conv_fwd(x, st[1])
                ops.bias_relu_(y, st[2])
        else:
            y = torch.empty((1, x.shape[1], x.shape[2] // 2, x.shape[3] // 2), dtype=torch.float32, device=x.device,
                            memory_format=_CL)
            ops.maxpool2x2(x, y)
        if keep:
            saved.append((x, y))
        for k in plan.taps_at.get(sidx, ()):
            taps[k] = y
        x = y
    return taps, saved


def features_backward(plan: FeaturePlan, saved, tap_grad, d_img: torch.Tensor, accumulate: bool) -> None:
    """Back-propagates through the feature path.  tap_grad(k, tap, g, relu_mask) must add tap k's loss gradient
    into g (a channels_last tensor like `tap`), allocating it when g is None, and return it; with relu_mask it
    must also apply the backward of the ReLU that produced `tap` (fused into the same kernel).  The image gradient
    is written (or added) into the planar (1,3,H,W) tensor d_img."""
    g = None
    masked = False       # the ReLU backward of the step below has already been applied (fused into pool backward)
    for sidx in range(plan.n_steps_needed - 1, -1, -1):
        st = plan.steps[sidx]
        x, y = saved[sidx]
        taps_here = plan.taps_at.get(sidx, ())
        for n, k in enumerate(taps_here):
            # the last tap gradient added at a conv step also applies that step's ReLU backward
            fuse = st[0] == 'conv' and n == len(taps_here) - 1 and not masked
            g = tap_grad(k, y, g, fuse)
            masked = masked or fuse
        if g is None:
            continue
        if st[0] == 'conv':
            if not masked:
                ops.relu_bwd_(g, y)
            masked = False
            g = _conv_bwd_data(g, x, st[1])
        else:
            gx = torch.empty_like(x, memory_format=_CL)
            fuse = sidx > 0 and plan.steps[sidx - 1][0] == 'conv' and (sidx - 1) not in plan.taps_at
            ops.maxpool2x2_bwd(g, x, gx, fuse)
            masked = fuse
            g = gx
    if g is None:
        if not accumulate:
            d_img.zero_()
        return
    ops.hwc_to_chw(g, d_img, d_img.shape[1], d_img.shape[2] * d_img.shape[3], accumulate)


class LevelTargets:
    """Per-level targets in the layout of the path: channels_last content map and the style Grams."""

    def __init__(self, content_cl: torch.Tensor, grams: Sequence[torch.Tensor]):
        self.content_cl = content_cl
        self.grams = list(grams)


def build_targets(plan: FeaturePlan, content_img: torch.Tensor, style_img: torch.Tensor, content_idx: int,
                  style_idx: Sequence[int], wss: ops.LevelWorkspaces) -> LevelTargets:
    """Targets of one level (neural_style_transfer.py:78-82) through the same path as the optimizing image."""
    with torch.no_grad():
        taps, _ = features_forward(plan, content_img, keep=False)
        content_cl = taps[content_idx].clone(memory_format=torch.preserve_format)
        del taps
        taps, _ = features_forward(plan, style_img, keep=False)
        grams = []
        for j, k in enumerate(style_idx):
            f = taps[k]
            c, hw = f.shape[1], f.shape[2] * f.shape[3]
            g = torch.empty((c, c), dtype=torch.float32, device=f.device)
            ws = ops.Workspace(ops.L.load().ast_gram_workspace_bytes(c, hw), f.device)
            ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), None, g, None, ws)
            grams.append(g)
        return LevelTargets(content_cl, grams)


class LevelPathFn(torch.autograd.Function):
    """One pyramid level of the Gatys loss, image in -> (total, content, style, tv) out, as ONE autograd node
    whose forward and backward are the explicit schedules above."""

    @staticmethod
    def forward(ctx, cfg, img):
        plan, targets, content_idx, style_idx, weights, wss = cfg[:6]
        bf16 = bool(cfg[6]) if len(cfg) > 6 else False
        tv_pre = cfg[7] if len(cfg) > 7 else None      # (sums2, tv) from the fused pyramid step that read this image
        dev = ops._require_cuda(img)
        cw, sw, tvw = (float(v) for v in weights)
        img = img.contiguous()
        need_grad = ctx.needs_input_grad[1]
        taps, saved = features_forward(plan, img, keep=need_grad)
        n_style = len(style_idx)
        vals = torch.empty(n_style + 2, dtype=torch.float32, device=dev)   # style mse[n] | content | tv
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        ds = {}
        for j, k in enumerate(style_idx):
            f = taps[k]
            c, hw = f.shape[1], f.shape[2] * f.shape[3]
            d = ops.new_d(c, bf16, dev)
            ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), targets.grams[j], d, vals[j], wss.for_gram(j, c, hw, dev),
                                  round_out=ops.d_round_mode(c, bf16))     # D only feeds the backward's tensor-core operand
            ds[k] = d
        xc = taps[content_idx]
        if xc.numel() != targets.content_cl.numel():
            raise ValueError('content feature map and target differ in size')
        ops.mse_fwd(xc, targets.content_cl, 1.0 / xc.numel(), vals[n_style], wss.for_reduce('content', dev))
        if tv_pre is not None:
            sums2, tv_val = tv_pre
        else:
            sums2 = torch.empty(2, dtype=torch.float32, device=dev)
            tv_val = vals[n_style + 1]
            ops.tv_fwd(img, sums2, tv_val, wss.for_reduce('tv', dev))
        ops._launch(dev, ('combine',), 'ast_level_combine', vals.data_ptr(), n_style, vals[n_style].data_ptr(),
                    tv_val.data_ptr(), cw, sw, tvw, out4.data_ptr())
        if need_grad:
            ctx.save_for_backward(img)
            ctx.pack = (plan, targets, content_idx, tuple(style_idx), (cw, sw, tvw), saved, ds, sums2)
        total, content, style, tv = out4[0], out4[1], out4[2], out4[3]
        ctx.mark_non_differentiable(content, style, tv)
        return total, content, style, tv

    @staticmethod
    def backward(ctx, g_total, g_content, g_style, g_tv):
        (img,) = ctx.saved_tensors
        plan, targets, content_idx, style_idx, (cw, sw, tvw), saved, ds, sums2 = ctx.pack
        ctx.pack = None
        dev = img.device
        gsc = ops._gscale(g_total, dev)
        n = len(style_idx)

        def tap_grad(k, tap, g, relu_mask):
            acc = g is not None
            if not acc:
                g = torch.empty_like(tap, memory_format=_CL)
            c, hw = tap.shape[1], tap.shape[2] * tap.shape[3]
            style, content = k in ds, k == content_idx
            # the Gram kernel's fused ReLU backward pays off up to C = 256 (measured, profiles/r01_gram_nhwc_sweep.jsonl)
            fuse_gram = relu_mask and not content and c <= FUSED_TAP_RELU_MAX_C
            if style:
                ops.gram_bwd_nhwc_auto(ds[k], tap, c, hw, (sw / n) * 4.0 / (float(c) * c * c * hw), gsc, g, acc,
                                       relu_mask=fuse_gram)
            if content:
                ops.mse_bwd(tap, targets.content_cl, cw * 2.0 / tap.numel(), gsc, g, acc or style, relu_mask)
            elif not style and not acc:
                g.zero_()
            if relu_mask and not content and not fuse_gram:
                ops.relu_bwd_(g, tap)
            return g

        d_img = torch.empty_like(img)
        ops.tv_bwd(img, sums2, tvw, gsc, d_img, False)
        features_backward(plan, saved, tap_grad, d_img, accumulate=True)
        return None, d_img
