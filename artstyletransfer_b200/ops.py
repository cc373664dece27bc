"""torch <-> C-ABI glue: tensor checks, caller-owned workspaces and the autograd Functions that put the
sm_100a kernels behind the reference's call surface.  PyTorch is plumbing here (device memory, streams,
autograd graph); every arithmetic op on the path is a kernel in libast_sm100.so.
"""
from __future__ import annotations

import os
import threading
from typing import Optional, Sequence

import torch

from . import _lib as L

DEFAULT_PRECISION = os.environ.get('AST_PRECISION', 'tf32')


def _prec(precision: Optional[str]) -> int:
    p = precision or DEFAULT_PRECISION
    if p not in L.PRECISIONS:
        raise ValueError(f'unknown precision {p!r}; choose from {sorted(L.PRECISIONS)}')
    return L.PRECISIONS[p]


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('artstyletransfer_b200 ops need CUDA tensors: the hot path is sm_100a CUDA only '
                               '(no CPU fallback); got a tensor on ' + str(t.device))
        if t.dtype != torch.float32:
            raise TypeError(f'expected float32, got {t.dtype}')
        dev = dev or t.device
        if t.device != dev:
            raise RuntimeError(f'tensors on different devices: {t.device} vs {dev}')
    return dev


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class KernelStats:
    """Optional instrumentation for bench.py: counts this library's kernel launches and, when `timing` is on,
    brackets every C-ABI call with CUDA events on the launching stream (resolved after a synchronize)."""

    def __init__(self):
        self.enabled = False
        self.timing = False
        self.launches = 0
        self.events = []      # (key, start_event, stop_event)

    def reset(self, enabled=True, timing=False):
        self.enabled, self.timing, self.launches, self.events = enabled, timing, 0, []

    def summary(self):
        out = {}
        for key, a, b in self.events:
            d = out.setdefault(key, [0, 0.0])
            d[0] += 1
            d[1] += a.elapsed_time(b)
        return {k: {'calls': v[0], 'ms_total': v[1], 'ms_avg': v[1] / v[0]} for k, v in out.items()}


STATS = KernelStats()

# kernels launched per C-ABI entry point
_KERNELS_PER_CALL = {'ast_gram_mse_fwd': 2, 'ast_gram_mse_fwd_nhwc': 2, 'ast_gram_bwd_nhwc': 1, 'ast_gram_bwd_nhwc_bf16': 1, 'ast_gram_finalize': 1, 'ast_gram_finalize_batch': 1, 'ast_gram_bwd': 1, 'ast_mse_fwd': 1, 'ast_mse_bwd': 1,
                     'ast_tv_fwd': 1, 'ast_tv_bwd': 1, 'ast_tv_bwd_rows': 1, 'ast_level_combine': 1, 'ast_bicubic_down2x': 1, 'ast_bicubic_down2x_tv': 1,
                     'ast_bicubic_down2x_adj': 1, 'ast_bicubic_resize': 1, 'ast_bicubic_resize_adj': 1,
                     'ast_noise_init': 1, 'ast_bias_relu_nhwc': 1, 'ast_relu_bwd': 1, 'ast_maxpool2x2_nhwc': 1,
                     'ast_maxpool2x2_bwd_nhwc': 1, 'ast_chw_to_hwc': 1, 'ast_hwc_to_chw': 1, 'ast_unprepare_hwc': 1, 'ast_halo_exchange': 1, 'ast_band_announce': 1, 'ast_band_gather': 1}


def _launch(dev: torch.device, key, name: str, *args) -> None:
    """Call an entry point on `dev`'s current stream (last ABI argument)."""
    st = torch.cuda.current_stream(dev)
    with _on(dev):
        if STATS.enabled:
            STATS.launches += _KERNELS_PER_CALL[name]
            if STATS.timing:
                a = torch.cuda.Event(enable_timing=True)
                b = torch.cuda.Event(enable_timing=True)
                a.record(st)
                L.call(name, *args, st.cuda_stream)
                b.record(st)
                STATS.events.append((key, a, b))
                return
        L.call(name, *args, st.cuda_stream)


class timed:
    """`with ops.timed(dev, key):` — CUDA events around a region that is NOT one of this library's kernels (a cuDNN
    convolution, a collective) when bench.py's per-kernel pass is on; free otherwise."""
    __slots__ = ('dev', 'key', 'a')

    def __init__(self, dev: torch.device, key):
        self.dev, self.key, self.a = dev, key, None

    def __enter__(self):
        if STATS.enabled and STATS.timing:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record(torch.cuda.current_stream(self.dev))

    def __exit__(self, *exc):
        if self.a is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record(torch.cuda.current_stream(self.dev))
            STATS.events.append((self.key, self.a, b))


class _on:
    """Make `dev` current for the duration of a launch (the C ABI launches on the current device)."""
    __slots__ = ('dev', 'prev')

    def __init__(self, dev: torch.device):
        self.dev = dev.index if dev.index is not None else torch.cuda.current_device()
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.dev:
            self.prev = cur
            torch.cuda.set_device(self.dev)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


class Workspace:
    """Caller-owned zero-initialised scratch (the C ABI keeps no hidden state).  Not shareable between
    host threads or between ops that may be in flight at once."""

    def __init__(self, nbytes: int, device: torch.device):
        self.buf = torch.zeros(max(int(nbytes), 16), dtype=torch.uint8, device=device)

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr()

    @property
    def nbytes(self) -> int:
        return self.buf.numel()


def gram_workspace(C: int, HW: int, device: torch.device) -> Workspace:
    return Workspace(L.load().ast_gram_workspace_bytes(C, HW), device)


def reduce_workspace(device: torch.device) -> Workspace:
    return Workspace(L.load().ast_reduce_workspace_bytes(), device)


_tls = threading.local()


def _thread_ws(kind: str, nbytes: int, device: torch.device) -> Workspace:
    """Per-thread workspace cache for the functional entry points (gram_matrix, total_variation, ...)."""
    cache = getattr(_tls, 'ws', None)
    if cache is None:
        cache = _tls.ws = {}
    key = (kind, device)
    ws = cache.get(key)
    if ws is None or ws.nbytes < nbytes:
        ws = cache[key] = Workspace(nbytes, device)
    return ws


# ------------------------------------------------------------------------------------------------------
# raw launches
# ------------------------------------------------------------------------------------------------------
def effective_precision(precision: int, ptr: int, C: int, HW: int, ld: int) -> int:
    """TF32 (tcgen05/TMA) when the operand is addressable by TMA, else the exact fp32 CUDA kernels — still this
    library's sm_100a code, never a CPU or torch fallback.  Only toy shapes (HW % 4 != 0) take the second branch."""
    if precision == L.AST_PREC_BF16:       # BF16 operands exist for the channels-last 512-channel backward only
        precision = L.AST_PREC_TF32
    if precision == L.AST_PREC_TF32 and not L.load().ast_gram_tf32_supported(ptr, C, HW, ld):
        return L.AST_PREC_FP32
    return precision


def gram_mse_fwd(feat: torch.Tensor, C: int, HW: int, scale: float, target: Optional[torch.Tensor],
                 out: torch.Tensor, loss: Optional[torch.Tensor], ws: Workspace, precision: int,
                 ld: Optional[int] = None, offset: int = 0) -> None:
    """feat: base tensor; the (C, HW) operand starts `offset` elements in and has row pitch `ld` (default HW)."""
    precision = effective_precision(precision, feat.data_ptr() + 4 * offset, C, HW, HW if ld is None else ld)
    _launch(feat.device, ('gram_fwd', C, HW, precision), 'ast_gram_mse_fwd', feat.data_ptr() + 4 * offset, C, HW,
            HW if ld is None else ld, scale,
            target.data_ptr() if target is not None else None, out.data_ptr(),
            loss.data_ptr() if loss is not None else None, ws.ptr, ws.nbytes, precision)


def gram_finalize(g_raw: torch.Tensor, C: int, scale: float, target: Optional[torch.Tensor], out: torch.Tensor,
                  loss: Optional[torch.Tensor], ws: Workspace, round_out: bool = False) -> None:
    _launch(g_raw.device, ('gram_finalize', C), 'ast_gram_finalize', g_raw.data_ptr(), C, scale,
            target.data_ptr() if target is not None else None, out.data_ptr(),
            loss.data_ptr() if loss is not None else None, ws.ptr, ws.nbytes, int(round_out))


def gram_finalize_batch(items, ws: Workspace) -> None:
    """items: sequence of (g_raw, C, scale, target or None, out, loss or None, round_out) — ast_gram_finalize for all of
    them in one launch; ws from finalize_batch_workspace(len(items), device)."""
    if len(items) > L.AST_FINALIZE_MAX_ITEMS:
        raise ValueError(f'at most {L.AST_FINALIZE_MAX_ITEMS} Grams per batched finalize; got {len(items)}')
    arr = (L.FinalizeItem * len(items))()
    for i, (g_raw, c, scale, target, out, loss, round_out) in enumerate(items):
        arr[i].G_raw, arr[i].A = g_raw.data_ptr(), (target.data_ptr() if target is not None else None)
        arr[i].out, arr[i].loss = out.data_ptr(), (loss.data_ptr() if loss is not None else None)
        arr[i].scale, arr[i].C, arr[i].round_out = float(scale), int(c), int(round_out)
    _launch(items[0][0].device, ('gram_finalize_batch', len(items)), 'ast_gram_finalize_batch', arr, len(items), ws.ptr,
            ws.nbytes)


def finalize_batch_workspace(n_items: int, device: torch.device) -> Workspace:
    return Workspace(L.load().ast_finalize_batch_workspace_bytes(n_items), device)


def gram_bwd(D: torch.Tensor, feat: torch.Tensor, C: int, HW: int, scale: float, gscale: Optional[torch.Tensor],
             dF: torch.Tensor, accumulate: bool, precision: int, ld: Optional[int] = None, offset: int = 0) -> None:
    precision = effective_precision(precision, feat.data_ptr() + 4 * offset, C, HW, HW if ld is None else ld)
    if (dF.data_ptr() + 4 * offset) % 16 or D.data_ptr() % 16:
        precision = L.AST_PREC_FP32
    _launch(feat.device, ('gram_bwd', C, HW, precision), 'ast_gram_bwd', D.data_ptr(), feat.data_ptr() + 4 * offset, C,
            HW, HW if ld is None else ld, scale, gscale.data_ptr() if gscale is not None else None,
            dF.data_ptr() + 4 * offset, int(accumulate), precision)


def gram_mse_fwd_nhwc(feat: torch.Tensor, C: int, HW: int, scale: float, target: Optional[torch.Tensor],
                      out: torch.Tensor, loss: Optional[torch.Tensor], ws: Workspace, offset: int = 0,
                      round_out: bool = False) -> None:
    """feat: (HW, C) row-major storage (torch channels_last); the operand starts `offset` elements in.
    round_out: store `out` (= D) rounded to TF32 for gram_bwd_nhwc(..., d_prerounded=True)."""
    _launch(feat.device, ('gram_fwd_nhwc', C, HW), 'ast_gram_mse_fwd_nhwc', feat.data_ptr() + 4 * offset, C, HW, scale,
            target.data_ptr() if target is not None else None, out.data_ptr(),
            loss.data_ptr() if loss is not None else None, ws.ptr, ws.nbytes, int(round_out))


def gram_bwd_nhwc(D: torch.Tensor, feat: torch.Tensor, C: int, HW: int, scale: float, gscale: Optional[torch.Tensor],
                  dF: torch.Tensor, accumulate: bool, offset: int = 0, d_prerounded: bool = False,
                  relu_mask: bool = False) -> None:
    """relu_mask: feat is a ReLU output; write the gradient w.r.t. the ReLU's input (threshold_backward fused)."""
    _launch(feat.device, ('gram_bwd_nhwc', C, HW, int(accumulate) + 2 * int(relu_mask)), 'ast_gram_bwd_nhwc',
            D.data_ptr(), feat.data_ptr() + 4 * offset, C, HW, scale,
            gscale.data_ptr() if gscale is not None else None, dF.data_ptr() + 4 * offset, int(accumulate),
            int(d_prerounded), int(relu_mask))


BF16_MIN_C = 512      # AST_PREC_BF16: layers at least this wide take bfloat16 operands in the backward (the tensor-bound ones)


def d_round_mode(C: int, bf16: bool) -> int:
    """round_out of the finalize kernels for a D that only feeds the backward: 2 = bfloat16 (BF16 mode, C = 512),
    1 = TF32-representable fp32."""
    return 2 if (bf16 and C >= BF16_MIN_C) else 1


def new_d(C: int, bf16: bool, device: torch.device) -> torch.Tensor:
    return torch.empty((C, C), dtype=torch.bfloat16 if (bf16 and C >= BF16_MIN_C) else torch.float32, device=device)


def gram_bwd_nhwc_auto(D: torch.Tensor, feat: torch.Tensor, C: int, HW: int, scale: float, gscale, dF: torch.Tensor,
                       accumulate: bool, relu_mask: bool = False, offset: int = 0) -> None:
    """Backward with a D produced by the finalize kernels (round_out from d_round_mode): bfloat16 D -> BF16 kernel."""
    if D.dtype == torch.bfloat16:
        _launch(feat.device, ('gram_bwd_nhwc_bf16', C, HW, int(accumulate) + 2 * int(relu_mask)), 'ast_gram_bwd_nhwc_bf16',
                D.data_ptr(), feat.data_ptr() + 4 * offset, C, HW, scale, gscale.data_ptr() if gscale is not None else None,
                dF.data_ptr() + 4 * offset, int(accumulate), int(relu_mask))
    else:
        gram_bwd_nhwc(D, feat, C, HW, scale, gscale, dF, accumulate, offset=offset, d_prerounded=True, relu_mask=relu_mask)


def _gscale(g: Optional[torch.Tensor], dev: torch.device) -> Optional[torch.Tensor]:
    if g is None:
        return None
    if g.dtype != torch.float32 or g.device != dev or not g.is_contiguous():
        g = g.to(device=dev, dtype=torch.float32).contiguous()
    return g


# ------------------------------------------------------------------------------------------------------
# glue around the cuDNN convolutions (csrc/vgg_glue.cu); tensors are torch channels_last (1, C, H, W)
# ------------------------------------------------------------------------------------------------------
def bias_relu_(y: torch.Tensor, bias: torch.Tensor) -> None:
    c = y.shape[1]
    _launch(y.device, ('bias_relu', c, y.numel()), 'ast_bias_relu_nhwc', y.data_ptr(), bias.data_ptr(), c,
            y.numel() // c)


def relu_bwd_(g: torch.Tensor, r: torch.Tensor) -> None:
    _launch(g.device, ('relu_bwd', g.numel()), 'ast_relu_bwd', g.data_ptr(), r.data_ptr(), g.numel())


def maxpool2x2(x: torch.Tensor, y: torch.Tensor) -> None:
    _, c, h, w = x.shape
    _launch(x.device, ('maxpool', c, h, w), 'ast_maxpool2x2_nhwc', x.data_ptr(), c, h, w, y.data_ptr())


def maxpool2x2_bwd(gy: torch.Tensor, x: torch.Tensor, gx: torch.Tensor, relu_mask: bool) -> None:
    _, c, h, w = x.shape
    _launch(x.device, ('maxpool_bwd', c, h, w), 'ast_maxpool2x2_bwd_nhwc', gy.data_ptr(), x.data_ptr(), c, h, w,
            int(relu_mask), gx.data_ptr())


def chw_to_hwc(x: torch.Tensor, y: torch.Tensor, c: int, hw: int, plane: Optional[int] = None, x_off: int = 0,
               y_off: int = 0) -> None:
    """planar -> interleaved; offsets in elements, `plane` = stride between the planar side's channel planes."""
    _launch(x.device, ('chw_to_hwc', c, hw), 'ast_chw_to_hwc', x.data_ptr() + 4 * x_off, c, hw,
            hw if plane is None else plane, y.data_ptr() + 4 * y_off)


def hwc_to_chw(x: torch.Tensor, y: torch.Tensor, c: int, hw: int, accumulate: bool, plane: Optional[int] = None,
               x_off: int = 0, y_off: int = 0) -> None:
    _launch(x.device, ('hwc_to_chw', c, hw), 'ast_hwc_to_chw', x.data_ptr() + 4 * x_off, c, hw,
            y.data_ptr() + 4 * y_off, hw if plane is None else plane, int(accumulate))


def unprepare_hwc(x: torch.Tensor, y: torch.Tensor, mean) -> None:
    """(1,3,H,W) planar 'x*255 - mean' image -> (H,W,3) in [0,1] (neural_style_transfer.py:388-393), one pass."""
    hw = x.shape[-2] * x.shape[-1]
    _launch(x.device, ('unprepare', hw), 'ast_unprepare_hwc', x.data_ptr(), hw, float(mean[0]), float(mean[1]),
            float(mean[2]), y.data_ptr())


# ------------------------------------------------------------------------------------------------------
# math_utils.gram_matrix (math_utils.py:26-34), differentiable
# ------------------------------------------------------------------------------------------------------
class GramMatrixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, should_normalize: bool, precision: int):
        _require_cuda(x)
        if x.dim() != 4:
            raise ValueError(f'gram_matrix expects (b, ch, h, w); got {tuple(x.shape)}')
        x = x.contiguous()
        b, ch, h, w = x.shape
        hw = h * w
        scale = 1.0 / (ch * h * w) if should_normalize else 1.0
        out = torch.empty((b, ch, ch), dtype=torch.float32, device=x.device)
        ws = _thread_ws('gram', L.load().ast_gram_workspace_bytes(ch, hw), x.device)
        for i in range(b):
            gram_mse_fwd(x[i], ch, hw, scale, None, out[i], None, ws, precision)
        ctx.save_for_backward(x)
        ctx.scale = scale
        ctx.precision = precision
        return out

    @staticmethod
    def backward(ctx, g_out: torch.Tensor):
        (x,) = ctx.saved_tensors
        b, ch, h, w = x.shape
        # dL/dF = scale * (gG + gG^T) F
        dsym = (g_out + g_out.transpose(1, 2)).contiguous()
        dx = torch.empty_like(x)
        for i in range(b):
            gram_bwd(dsym[i], x[i], ch, h * w, ctx.scale, None, dx[i], False, ctx.precision)
        return dx, None, None


def gram_matrix(x: torch.Tensor, should_normalize: bool = True, precision: Optional[str] = None) -> torch.Tensor:
    return GramMatrixFn.apply(x, bool(should_normalize), _prec(precision))


# ------------------------------------------------------------------------------------------------------
# StyleLoss: mean((A - G(F))^2) fused into the Gram kernel's finalize; backward (G - A) F
# ------------------------------------------------------------------------------------------------------
class StyleLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, target: torch.Tensor, ws: Workspace, precision: int):
        _require_cuda(x, target)
        x = x.contiguous()
        ch = x.shape[-3]
        hw = x.shape[-2] * x.shape[-1]
        if x.numel() != ch * hw:
            raise ValueError('StyleLoss expects a single feature map (batch 1)')
        if tuple(target.shape[-2:]) != (ch, ch):
            raise ValueError(f'target Gram {tuple(target.shape)} does not match {ch} channels')
        d = torch.empty((ch, ch), dtype=torch.float32, device=x.device)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        gram_mse_fwd(x, ch, hw, 1.0 / (ch * hw), target.contiguous(), d, loss, ws, precision)
        ctx.save_for_backward(x, d)
        ctx.precision = precision
        return loss

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        x, d = ctx.saved_tensors
        ch = x.shape[-3]
        hw = x.shape[-2] * x.shape[-1]
        dx = torch.empty_like(x)
        gram_bwd(d, x, ch, hw, 4.0 / (float(ch) * ch * ch * hw), _gscale(g, x.device), dx, False, ctx.precision)
        return dx, None, None, None


# ------------------------------------------------------------------------------------------------------
# ContentLoss: mean((T - X)^2)   (neural_style_transfer.py:95)
# ------------------------------------------------------------------------------------------------------
def mse_fwd(x, t, scale, loss, ws):
    _launch(x.device, ('mse_fwd', x.numel()), 'ast_mse_fwd', x.data_ptr(), t.data_ptr(), x.numel(), scale,
            loss.data_ptr(), ws.ptr, ws.nbytes)


def mse_bwd(x, t, scale, gscale, dx, accumulate, relu_mask=False):
    _launch(x.device, ('mse_bwd', x.numel(), int(bool(accumulate))), 'ast_mse_bwd', x.data_ptr(), t.data_ptr(), x.numel(), scale,
            gscale.data_ptr() if gscale is not None else None, dx.data_ptr(), int(accumulate), int(relu_mask))


class ContentLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, target: torch.Tensor, ws: Workspace):
        _require_cuda(x, target)
        x = x.contiguous()
        target = target.contiguous()
        if x.numel() != target.numel():
            raise ValueError(f'content shapes differ: {tuple(x.shape)} vs {tuple(target.shape)}')
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        mse_fwd(x, target, 1.0 / x.numel(), loss, ws)
        ctx.save_for_backward(x, target)
        return loss

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        x, target = ctx.saved_tensors
        dx = torch.empty_like(x)
        mse_bwd(x, target, 2.0 / x.numel(), _gscale(g, x.device), dx, False)
        return dx, None, None


# ------------------------------------------------------------------------------------------------------
# total variation (math_utils.py:37-41)
# ------------------------------------------------------------------------------------------------------
def tv_fwd(y, sums2, tv, ws):
    c = y.numel() // (y.shape[-2] * y.shape[-1])
    _launch(y.device, ('tv_fwd', y.numel()), 'ast_tv_fwd', y.data_ptr(), c, y.shape[-2], y.shape[-1], sums2.data_ptr(),
            tv.data_ptr() if tv is not None else None, ws.ptr, ws.nbytes)


def tv_bwd(y, sums2, weight, gscale, dy, accumulate, rows=None):
    """rows = (r0, r1): only those rows of every plane (ast_tv_bwd_rows); the scale is the whole image's."""
    dev = y.device
    h, w = y.shape[-2], y.shape[-1]
    c = y.numel() // (h * w)
    nx, ny = float(c) * h * (w - 1), float(c) * (h - 1) * w
    if rows is not None:
        r0, r1 = int(rows[0]), int(rows[1])
        _launch(dev, ('tv_bwd_rows', c * (r1 - r0) * w), 'ast_tv_bwd_rows', y.data_ptr(), c, h, w, r0, r1, sums2.data_ptr(),
                2.0 * weight / (nx * nx), 2.0 * weight / (ny * ny), gscale.data_ptr() if gscale is not None else None,
                dy.data_ptr(), int(accumulate))
        return
    _launch(dev, ('tv_bwd', y.numel()), 'ast_tv_bwd', y.data_ptr(), c, h, w, sums2.data_ptr(), 2.0 * weight / (nx * nx),
            2.0 * weight / (ny * ny), gscale.data_ptr() if gscale is not None else None, dy.data_ptr(),
            int(accumulate))


class TotalVariationFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y: torch.Tensor):
        _require_cuda(y)
        y = y.contiguous()
        if y.dim() < 2 or y.shape[-1] < 2 or y.shape[-2] < 2:
            raise ValueError(f'total_variation needs H, W >= 2; got {tuple(y.shape)}')
        sums2 = torch.empty(2, dtype=torch.float32, device=y.device)
        tv = torch.empty((), dtype=torch.float32, device=y.device)
        tv_fwd(y, sums2, tv, _thread_ws('reduce', L.load().ast_reduce_workspace_bytes(), y.device))
        ctx.save_for_backward(y, sums2)
        return tv

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        y, sums2 = ctx.saved_tensors
        dy = torch.empty_like(y)
        tv_bwd(y, sums2, 1.0, _gscale(g, y.device), dy, False)
        return dy


def total_variation(y: torch.Tensor) -> torch.Tensor:
    return TotalVariationFn.apply(y)


# ------------------------------------------------------------------------------------------------------
# bicubic pyramid step + adjoint (neural_style_transfer.py:173-176)
# ------------------------------------------------------------------------------------------------------
def _planes(x: torch.Tensor):
    h, w = x.shape[-2], x.shape[-1]
    return x.numel() // (h * w), h, w


def bicubic_down_raw(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """F.interpolate(x, size=(out_h, out_w), mode='bicubic') for (..., H, W) CUDA float32."""
    _require_cuda(x)
    x = x.contiguous()
    c, h, w = _planes(x)
    y = torch.empty(x.shape[:-2] + (out_h, out_w), dtype=torch.float32, device=x.device)
    if h == 2 * out_h and w == 2 * out_w:
        _launch(x.device, ('down2x', c, h, w), 'ast_bicubic_down2x', x.data_ptr(), c, h, w, y.data_ptr())
    else:
        _launch(x.device, ('resize', c, h, w, out_h, out_w), 'ast_bicubic_resize', x.data_ptr(), c, h, w, y.data_ptr(),
                out_h, out_w, L.AST_LAYOUT_CHW, L.AST_COORD_TORCH)
    return y


def bicubic_down_tv_raw(x: torch.Tensor, wss: 'LevelWorkspaces'):
    """Exact 2x down-sampling of (..., H, W) (even H, W) fused with total_variation(x): returns (y, sums2, tv) where
    sums2 / tv are what tv_fwd(x) would have produced — one pass over x instead of two."""
    _require_cuda(x)
    x = x.contiguous()
    c, h, w = _planes(x)
    if h % 2 or w % 2:
        raise ValueError(f'fused pyramid step + TV needs even sizes; got {h}x{w}')
    dev = x.device
    y = torch.empty(x.shape[:-2] + (h // 2, w // 2), dtype=torch.float32, device=dev)
    sums2 = torch.empty(2, dtype=torch.float32, device=dev)
    tv = torch.empty((), dtype=torch.float32, device=dev)
    ws = wss.for_named('down_tv', L.load().ast_bicubic_down2x_tv_workspace_bytes(c, h, w), dev)
    _launch(dev, ('down2x_tv', c, h, w), 'ast_bicubic_down2x_tv', x.data_ptr(), c, h, w, y.data_ptr(), sums2.data_ptr(),
            tv.data_ptr(), ws.ptr, ws.nbytes)
    return y, sums2, tv


def bicubic_down_adj_raw(gy: torch.Tensor, in_h: int, in_w: int, gx: Optional[torch.Tensor] = None,
                         accumulate: bool = False) -> torch.Tensor:
    _require_cuda(gy)
    gy = gy.contiguous()
    c, oh, ow = _planes(gy)
    if gx is None:
        gx = torch.empty(gy.shape[:-2] + (in_h, in_w), dtype=torch.float32, device=gy.device)
        accumulate = False
    if in_h == 2 * oh and in_w == 2 * ow:
        _launch(gy.device, ('down2x_adj', c, in_h, in_w), 'ast_bicubic_down2x_adj', gy.data_ptr(), c, in_h, in_w,
                gx.data_ptr(), int(accumulate))
    else:
        _launch(gy.device, ('resize_adj', c, in_h, in_w, oh, ow), 'ast_bicubic_resize_adj', gy.data_ptr(), c, in_h, in_w,
                oh, ow, gx.data_ptr(), int(accumulate), L.AST_COORD_TORCH)
    return gx


class BicubicHalfFn(torch.autograd.Function):
    """level_i = F.interpolate(level_{i-1}, size=(H//2, W//2), mode='bicubic') with a deterministic adjoint."""

    @staticmethod
    def forward(ctx, x: torch.Tensor):
        ctx.in_hw = (x.shape[-2], x.shape[-1])
        return bicubic_down_raw(x, x.shape[-2] // 2, x.shape[-1] // 2)

    @staticmethod
    def backward(ctx, gy: torch.Tensor):
        return bicubic_down_adj_raw(gy, *ctx.in_hw)


def bicubic_half(x: torch.Tensor) -> torch.Tensor:
    return BicubicHalfFn.apply(x)


class BicubicHalfTvFn(torch.autograd.Function):
    """bicubic_half(x) that also returns total_variation(x)'s two sums and value (non-differentiable by-products: the
    TV term's gradient is taken by the level's loss node from sums2, as before)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, wss):
        ctx.in_hw = (x.shape[-2], x.shape[-1])
        y, sums2, tv = bicubic_down_tv_raw(x, wss)
        ctx.mark_non_differentiable(sums2, tv)
        return y, sums2, tv

    @staticmethod
    def backward(ctx, gy, g_sums2, g_tv):
        return bicubic_down_adj_raw(gy, *ctx.in_hw), None


def bicubic_half_tv(x: torch.Tensor, wss: 'LevelWorkspaces'):
    return BicubicHalfTvFn.apply(x, wss)


def bicubic_resize(x: torch.Tensor, out_h: int, out_w: int, layout: str = 'chw', coord: str = 'cv2') -> torch.Tensor:
    """General-ratio bicubic (no autograd).  layout 'chw': (..., H, W); 'hwc': (H, W, C)."""
    _require_cuda(x)
    x = x.contiguous()
    dev = x.device
    cm = {'torch': L.AST_COORD_TORCH, 'cv2': L.AST_COORD_CV2}[coord]
    if layout == 'hwc':
        h, w, c = x.shape
        y = torch.empty((out_h, out_w, c), dtype=torch.float32, device=dev)
        lay = L.AST_LAYOUT_HWC
    else:
        c, h, w = _planes(x)
        y = torch.empty(x.shape[:-2] + (out_h, out_w), dtype=torch.float32, device=dev)
        lay = L.AST_LAYOUT_CHW
    _launch(dev, ('resize', c, h, w, out_h, out_w), 'ast_bicubic_resize', x.data_ptr(), c, h, w, y.data_ptr(), out_h,
            out_w, lay, cm)
    return y


# ------------------------------------------------------------------------------------------------------
# One pyramid level of the Gatys loss as ONE autograd node (LossBuilder.build, nst.py:84-112)
# ------------------------------------------------------------------------------------------------------
class LevelWorkspaces:
    """Per-LossBuilder scratch: one Gram workspace per style layer + reduce workspaces (content, tv)."""

    def __init__(self):
        self.gram = {}
        self.content = None
        self.tv = None

    def for_gram(self, k: int, C: int, HW: int, dev: torch.device) -> Workspace:
        need = L.load().ast_gram_workspace_bytes(C, HW)
        ws = self.gram.get(k)
        if ws is None or ws.nbytes < need or ws.buf.device != dev:
            ws = self.gram[k] = Workspace(need, dev)
        return ws

    def for_named(self, name: str, nbytes: int, dev: torch.device) -> Workspace:
        ws = self.gram.get(name)
        if ws is None or ws.nbytes < nbytes or ws.buf.device != dev:
            ws = self.gram[name] = Workspace(nbytes, dev)
        return ws

    def for_reduce(self, which: str, dev: torch.device) -> Workspace:
        ws = getattr(self, which)
        if ws is None or ws.buf.device != dev:
            ws = reduce_workspace(dev)
            setattr(self, which, ws)
        return ws


class LevelLossFn(torch.autograd.Function):
    """inputs: image, content feature map, then the style feature maps.
    outputs: (total, content, style, tv) — only `total` is differentiable (the others are what the reference
    prints).  Backward launches one kernel per term with the upstream gradient read on the device."""

    @staticmethod
    def forward(ctx, cfg, img, content_feat, *style_feats):
        (target_content, target_grams, weights, wss, precision) = cfg
        dev = _require_cuda(img, content_feat, *style_feats)
        cw, sw, tvw = (float(v) for v in weights)
        n_style = len(style_feats)
        img = img.contiguous()
        content_feat = content_feat.contiguous()
        style_feats = [f.contiguous() for f in style_feats]
        vals = torch.empty(n_style + 2, dtype=torch.float32, device=dev)   # style mse[n] | content | tv
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        ds = []
        for k, (f, a) in enumerate(zip(style_feats, target_grams)):
            ch, hw = f.shape[-3], f.shape[-2] * f.shape[-1]
            d = torch.empty((ch, ch), dtype=torch.float32, device=dev)
            gram_mse_fwd(f, ch, hw, 1.0 / (ch * hw), a, d, vals[k], wss.for_gram(k, ch, hw, dev), precision)
            ds.append(d)
        if content_feat.numel() != target_content.numel():
            raise ValueError('content feature map and target differ in size')
        mse_fwd(content_feat, target_content, 1.0 / content_feat.numel(), vals[n_style], wss.for_reduce('content', dev))
        sums2 = torch.empty(2, dtype=torch.float32, device=dev)
        tv_fwd(img, sums2, vals[n_style + 1], wss.for_reduce('tv', dev))
        _launch(dev, ('combine',), 'ast_level_combine', vals.data_ptr(), n_style, vals[n_style].data_ptr(),
                vals[n_style + 1].data_ptr(), cw, sw, tvw, out4.data_ptr())
        ctx.save_for_backward(img, content_feat, target_content, sums2, *style_feats, *ds)
        ctx.n_style = n_style
        ctx.weights = (cw, sw, tvw)
        ctx.precision = precision
        total, content, style, tv = out4[0], out4[1], out4[2], out4[3]
        ctx.mark_non_differentiable(content, style, tv)
        return total, content, style, tv

    @staticmethod
    def backward(ctx, g_total, g_content, g_style, g_tv):
        saved = ctx.saved_tensors
        img, content_feat, target_content, sums2 = saved[:4]
        n = ctx.n_style
        style_feats, ds = saved[4:4 + n], saved[4 + n:4 + 2 * n]
        cw, sw, tvw = ctx.weights
        dev = img.device
        g = _gscale(g_total, dev)
        need = ctx.needs_input_grad    # (cfg, img, content_feat, *style_feats)
        d_img = d_content = None
        if need[1]:
            d_img = torch.empty_like(img)
            tv_bwd(img, sums2, tvw, g, d_img, False)
        if need[2]:
            d_content = torch.empty_like(content_feat)
            mse_bwd(content_feat, target_content, cw * 2.0 / content_feat.numel(), g, d_content, False)
        d_style = []
        for k in range(n):
            if not need[3 + k]:
                d_style.append(None)
                continue
            f = style_feats[k]
            ch, hw = f.shape[-3], f.shape[-2] * f.shape[-1]
            df = torch.empty_like(f)
            gram_bwd(ds[k], f, ch, hw, (sw / n) * 4.0 / (float(ch) * ch * ch * hw), g, df, False, ctx.precision)
            d_style.append(df)
        return (None, d_img, d_content, *d_style)


# ------------------------------------------------------------------------------------------------------
# structured-noise init (neural_style_transfer.py:265-362)
# ------------------------------------------------------------------------------------------------------
def noise_init(content_hwc: Optional[torch.Tensor], H: int, W: int, levels: Sequence[dict], noise_factor: float,
               mode: str, use_gradient_map: bool, blur_w0: float, blur_w1: float, device: torch.device) -> torch.Tensor:
    """levels: dicts with kind (0/1/2), lowres (device (lh,lw,3) tensor or None), gy, gx (device float64 vectors),
    center, central, peripheral."""
    import ctypes as C
    if len(levels) > L.AST_NOISE_MAX_LEVELS:
        raise RuntimeError(f'at most {L.AST_NOISE_MAX_LEVELS} noise levels are supported')
    arr = (L.NoiseLevel * max(len(levels), 1))()
    keep = []
    for i, lv in enumerate(levels):
        low = lv.get('lowres')
        if low is not None:
            _require_cuda(low)
            low = low.contiguous()
            keep.append(low)
            arr[i].lowres, arr[i].lh, arr[i].lw = low.data_ptr(), low.shape[0], low.shape[1]
        arr[i].kind = int(lv['kind'])
        if lv.get('gy') is not None:
            gy, gx = lv['gy'].contiguous(), lv['gx'].contiguous()
            if gy.dtype != torch.float64 or gx.dtype != torch.float64 or gy.numel() != H or gx.numel() != W:
                raise ValueError('envelope vectors must be float64 of length H / W')
            keep += [gy, gx]
            arr[i].gy, arr[i].gx = gy.data_ptr(), gx.data_ptr()
            arr[i].center = float(lv['center'])
        arr[i].central, arr[i].peripheral = float(lv.get('central', 0.0)), float(lv.get('peripheral', 0.0))
    out = torch.empty((H, W, 3), dtype=torch.float32, device=device)
    m = {'random': L.AST_INIT_RANDOM, 'content+noise': L.AST_INIT_CONTENT_NOISE}[mode]
    if content_hwc is not None:
        _require_cuda(content_hwc)
        content_hwc = content_hwc.contiguous()
    _launch(device, ('noise_init', H, W), 'ast_noise_init', content_hwc.data_ptr() if content_hwc is not None else None,
            H, W, C.cast(arr, C.POINTER(L.NoiseLevel)), len(levels), float(noise_factor), m, int(use_gradient_map),
            float(blur_w0), float(blur_w1), out.data_ptr())
    return out
