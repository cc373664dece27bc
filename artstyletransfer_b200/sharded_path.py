"""Row-band sharded levels on the channels-last feature path, with per-layer halo exchange.

A rank owns a band of rows of a level (which rows of which level: parallel.PyramidBands; band edges are multiples of
16 so the four 2x2 max-pools never straddle ranks; a rank may own nothing of a level) and ONLY computes those plus
its halo rows: a 3x3 convolution needs one activation row of each neighbour (parallel.halo_exchange; rows are
contiguous in NHWC, so they go over NVLink in place, no packing).  A band carries D halo rows per side
(parallel.halo_schedule): with D = 2 one exchange serves two stacked convolutions — the first one leaves one valid,
redundantly computed halo row — so only every second convolution waits for the neighbours.  The backward uses the
SAME exchange on the gradient w.r.t. each convolution's output: with the neighbours' edge gradient rows in the halos,
an ordinary symmetric-padding backward-data convolution over the padded band yields complete gradients for the owned
rows (and, with D = 2, for one halo row per side, on which the next step applies its tap gradients and ReLU mask
itself) — cuDNN's asymmetric-padding backward-data, which the mirrored "send halo gradients back" scheme needs, runs
1.5x slower on B200.  The first convolution needs no exchange in the forward: every rank holds the whole (3-channel)
level image.

Activations live in padded bands (D halo rows above and below the owned rows; at the image border they are zero =
the convolution's zero padding).  The convolutions run as cuDNN's fused conv + bias + ReLU over the whole padded band
(FUSED_BAND_CONV): the outermost output row per side saw the band's zero padding instead of a real row and is junk —
with D = 1 exactly the halo row the next exchange overwrites, with D = 2 the row the second convolution of a pair
does not need.

Per closure (D = 2): 6 + 7 exchanges for ALL levels of the rank (two rows, <= 1.6 MB per side), per-tap all-reduces
(sum) of the packed raw Grams + content SSE of all levels overlapped with the layers above the tap, then every rank
finalises identically.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch

from . import feature_path as fp
from . import ops
from . import parallel as par

_CL = torch.channels_last
MULTI_STREAM = os.environ.get('AST_LEVEL_STREAMS', '1') != '0'
LANE_PRIORITY = int(os.environ.get('AST_LANE_PRIORITY', '0'))
# '1': the band convolutions run as cuDNN's fused conv + bias + ReLU over the WHOLE padded band with symmetric
# padding — (h + 2) rows in, (h + 2) rows out; the h owned rows only ever see real rows (their halos), the two outer
# output rows saw the zero padding, are junk, and are exactly the halo rows of the next layer: the next exchange
# overwrites them (at the image border they are zeroed).  Costs 2 / (h + 2) extra convolution rows and removes the
# separate bias + ReLU pass (8 B per activation element, 3 ms of an unsharded L=3 closure).  '0': the convolution
# writes the owned rows through cudnn_convolution.out into a persistent band, ast_bias_relu_nhwc follows.
FUSED_BAND_CONV = os.environ.get('AST_SHARD_FUSED_CONV', '1') != '0'
FUSED_PYRAMID_TV = os.environ.get('AST_FUSED_PYRAMID_TV', '1') != '0'
# '1': every tap's partial Grams (all levels) are all-reduced as soon as they exist, on a side stream under the
# convolutions above the tap; '0': one all-reduce of the whole packed buffer after the forward (round 1).
OVERLAP_GRAM_ALLREDUCE = os.environ.get('AST_OVERLAP_GRAM_ALLREDUCE', '1') != '0'


class Lanes:
    """Fork / join of the per-level work between two exchange points.  Lane 0 (the largest level) is the current
    stream; every other level gets its own stream that waits for the fork point and is joined before the next
    grouped exchange.  The pyramid levels' kernels are independent between exchanges, and the small levels' tiny
    launches (a 32-row band of the 256x384 level is a handful of CTAs) disappear in the shadow of the big one's.
    Inside a CUDA-graph capture the forks become parallel branches of the graph."""

    def __init__(self, device: torch.device, n: int):
        self.device = device
        # AST_LANE_PRIORITY=-1 runs the small levels' lanes at high priority; measured on a B200 it changes nothing
        # (tests/tools/rank_time_probe.py: 4.05 vs 4.06 ms for the two-band rank), so the default stays 0
        self.side = [torch.cuda.Stream(device, priority=LANE_PRIORITY) for _ in range(max(n - 1, 0))]

    def each(self, indices, fn) -> None:
        """fn(li) for every li in indices; li == indices[0] on the current stream, the rest on side streams."""
        indices = list(indices)
        if len(indices) <= 1 or not self.side:
            for li in indices:
                fn(li)
            return
        main = torch.cuda.current_stream(self.device)
        fork = torch.cuda.Event()
        fork.record(main)
        joins = []
        for n, li in enumerate(indices[1:]):
            st = self.side[n % len(self.side)]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                fn(li)
                ev = torch.cuda.Event()
                ev.record(st)
            joins.append(ev)
        fn(indices[0])
        for ev in joins:
            main.wait_event(ev)


_SERIAL = Lanes.__new__(Lanes)
_SERIAL.device, _SERIAL.side = None, []


def _rows(t: torch.Tensor) -> torch.Tensor:
    """(1, C, h, w) channels_last -> contiguous (h, w, C) view."""
    return t.permute(0, 2, 3, 1)[0]


class ShardedPathLevel:
    """Replaces LossBuilder.build for one level when sharding is on (same return contract)."""

    def __init__(self, group, plan: fp.FeaturePlan, content_img: torch.Tensor, style_img: torch.Tensor,
                 content_idx: int, style_idx: Sequence[int], weights, height: int, width: int, band=None,
                 halo_depth: int = None):
        """band: (r0, r1, up, dn) from parallel.PyramidBands — the owned rows and the ranks holding the rows above /
        below (None at the image border); r0 == r1 when this rank does not work on the level (it still joins the
        all-reduce and finalises the loss).  Default: `world` equal bands, neighbours rank -+ 1.
        halo_depth: halo rows per side of every band (parallel.halo_schedule; the SAME on every rank and level of a
        lock-step pyramid: PyramidBands.halo_depth()).  Default: parallel.HALO_DEPTH for equal bands that are tall
        enough, 1 for an explicit band."""
        self.group, self.plan = group, plan
        self.cidx, self.sidx = content_idx, list(style_idx)
        self.weights = tuple(float(w) for w in weights)
        self.H, self.W = height, width
        if band is None:
            bp = par.BandPlan(height, group.rank, group.world, halo=0)
            band = (bp.r0, bp.r1, group.rank - 1 if group.rank > 0 else None,
                    group.rank + 1 if group.rank + 1 < group.world else None)
            if halo_depth is None:
                halo_depth = par.HALO_DEPTH if (height // group.world) // par.ALIGN >= par.HALO_DEPTH else 1
        self.D = 1 if halo_depth is None or not FUSED_BAND_CONV else int(halo_depth)
        if self.D not in (1, 2):
            raise ValueError(f'halo depth {self.D}: 1 or 2 rows')
        self.r0, self.r1, self.up, self.dn = band
        self.hb = self.r1 - self.r0
        if self.hb and self.hb // par.ALIGN < self.D:
            raise ValueError(f'a band of {self.hb} rows cannot hand {self.D} halo rows to its neighbours at stride '
                             f'{par.ALIGN}')
        from . import neural_style_transfer as _nst
        self.bf16 = ops._prec(_nst.PRECISION) == ops.L.AST_PREC_BF16
        if self.hb and (self.r0 % par.ALIGN or self.r1 % par.ALIGN or not 0 <= self.r0 < self.r1 <= height):
            raise ValueError(f'band rows [{self.r0}, {self.r1}) of a {height}-row level must be multiples of {par.ALIGN}')
        dev = self.device = content_img.device
        self.wss = ops.LevelWorkspaces()
        # targets: every rank runs the whole content / style image once at set-up (replicated, not on the hot path)
        tg = fp.build_targets(plan, content_img, style_img, content_idx, style_idx, self.wss)
        self.target_grams = tg.grams
        cs = par.LAYER_STRIDE[content_idx]
        crow = _rows(tg.content_cl)
        self.target_content_band = crow[self.r0 // cs:self.r1 // cs].contiguous()
        # the same band with one more row per side (zeros outside the image): the target under a content tap whose
        # gradient is applied on the owned rows +- 1 (a backward step that runs without an exchange)
        self._target_content_ext = None
        if self.hb:
            ext = torch.zeros((self.hb // cs + 2, *crow.shape[1:]), dtype=crow.dtype, device=crow.device)
            lo, hi = max(self.r0 // cs - 1, 0), min(self.r1 // cs + 1, crow.shape[0])
            ext[lo - (self.r0 // cs - 1):lo - (self.r0 // cs - 1) + hi - lo] = crow[lo:hi]
            self._target_content_ext = ext
        self.content_numel_global = tg.content_cl.numel()
        self.channels = [g.shape[-1] for g in tg.grams]
        self.offs, self.content_slot, self.n_packed = par.pack_layout(self.channels)
        self.fin_ws = [ops.reduce_workspace(dev) for _ in self.sidx]
        self.generation = 0
        self.bufs: List[torch.Tensor] = []
        if not self.hb:
            self.xin = None
            return
        # persistent padded activation bands: xin (the image band) and one per step
        c0 = plan.steps[0][3]
        self.xin = torch.empty((1, c0, self.hb + 2 * self.D, width), dtype=torch.float32, device=dev,
                               memory_format=_CL).zero_()
        c, h, w = c0, self.hb, width
        for sidx in range(plan.n_steps_needed):
            st = plan.steps[sidx]
            if st[0] == 'conv':
                c = st[4]
            else:
                h, w = h // 2, w // 2
            if st[0] == 'conv' and FUSED_BAND_CONV:
                self.bufs.append(None)            # the fused convolution allocates its padded output every closure
            else:
                self.bufs.append(torch.empty((1, c, h + 2 * self.D, w), dtype=torch.float32, device=dev,
                                             memory_format=_CL).zero_())

    def content_target(self, ext: int) -> torch.Tensor:
        """Content target rows under the owned rows +- ext (ext 0 or 1) of the tapped layer."""
        return self.target_content_band if ext == 0 else self._target_content_ext

    def build(self, level_img: torch.Tensor):
        return ShardPathFn.apply(self, level_img)


def _interior(buf: torch.Tensor, depth: int = 1, ext: int = 0) -> torch.Tensor:
    """The owned rows (+- ext) of a band padded with `depth` halo rows per side."""
    return buf[:, :, depth - ext:buf.shape[2] - depth + ext, :]


def _conv_fwd_into(x_pad, w, out):
    with ops.timed(x_pad.device, fp._ckey('cudnn_conv_fwd', out, w)):
        torch.ops.aten.cudnn_convolution.out(x_pad, w, [0, 1], [1, 1], [1, 1], 1, fp.CUDNN_BENCHMARK, False,
                                             torch.backends.cudnn.allow_tf32, out=out)


def _conv_relu_fwd_padded(x_pad, w, b):
    """cuDNN conv + bias + ReLU over the whole padded band (see FUSED_BAND_CONV): rows 1..h of the result are the
    owned rows, rows 0 and h + 1 are to be overwritten."""
    with ops.timed(x_pad.device, fp._ckey('cudnn_conv_bias_relu_fwd', x_pad, w)), fp._cudnn_mode():
        y = torch.cudnn_convolution_relu(x_pad, w, b, [1, 1], [1, 1], [1, 1], 1)
    return y if y.is_contiguous(memory_format=_CL) else y.contiguous(memory_format=_CL)


def _conv_bwd_data_padded(g_pad, x_pad, w):
    """Backward-data over the whole padded band (h + 2 rows in, h + 2 rows out, symmetric padding)."""
    with ops.timed(x_pad.device, fp._ckey('cudnn_conv_dgrad', g_pad, w)), fp._cudnn_mode():
        gi = torch.ops.aten.convolution_backward(g_pad, x_pad, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                 [True, False, False])[0]
    return gi if gi.is_contiguous(memory_format=_CL) else gi.contiguous(memory_format=_CL)


class ShardPathFn(torch.autograd.Function):
    """input: the full level image (every rank has it).  outputs as LevelPathFn; the returned image gradient is
    this rank's contribution (its band rows, plus the TV gradient on rank 0) — summed by sync_image_grad."""

    @staticmethod
    def forward(ctx, sh: ShardedPathLevel, level_img):
        out4, state = sharded_forward(sh, level_img)
        if ctx.needs_input_grad[1]:
            ctx.state = state
        total, content, style, tv = out4[0], out4[1], out4[2], out4[3]
        ctx.mark_non_differentiable(content, style, tv)
        return total, content, style, tv

    @staticmethod
    def backward(ctx, g_total, g_content, g_style, g_tv):
        state, ctx.state = ctx.state, None
        return None, sharded_backward(state, g_total)


def sharded_forward(sh: ShardedPathLevel, level_img: torch.Tensor):
    """Forward schedule of one sharded level.  Returns (out4 = [total, content, style, tv], state for backward)."""
    out4s, state = pyramid_forward([sh], [level_img])
    return out4s[0], state


def sharded_backward(state, g_total) -> torch.Tensor:
    """Backward schedule: this rank's contribution to the level image's gradient (g_total: upstream scalar or None)."""
    return pyramid_backward(state, g_total)[0]


class ShardedPyramid:
    """All sharded levels of a job evaluated in LOCK-STEP: the pyramid levels' feature paths do not depend on one
    another, so step s of every level runs before step s+1 of any, and the halo rows of all levels travel in ONE
    grouped exchange per step (6 + 7 per closure with two-row halos, not per level), the raw Grams of all levels in
    ONE all-reduce.  A send/recv group costs ~30 us of latency on NVLink whatever it carries (rows are <= 0.8 MB)."""

    def __init__(self, levels: List[ShardedPathLevel]):
        self.levels = list(levels)
        if len({sh.D for sh in levels}) != 1:
            raise ValueError('lock-step levels must carry the same halo depth')
        self.lanes = Lanes(levels[0].device, len(levels)) if MULTI_STREAM else _SERIAL
        self.comm = torch.cuda.Stream(levels[0].device)          # the partial-Gram all-reduces (pyramid_forward)
        # image gradient: an all-gather of the ranks' disjoint rows through NVLink peer memory instead of an all-reduce
        # of the whole image — with the peer-memory group (AST_HALO=peer) unless AST_GRAD_GATHER=0
        self.gather = None
        grp = levels[0].group
        if isinstance(grp, par.PeerHaloGroup) and grp.world > 1 and os.environ.get('AST_GRAD_GATHER', '1') != '0':
            gather = err = None
            try:
                gather = par.PeerGradGather(grp, levels[0].device, [(sh.H, sh.W) for sh in levels],
                                            [(sh.r0, sh.r1) for sh in levels])
            except Exception as e:                # e.g. more ranks than the kernel's descriptor holds
                err = e
            if par.all_ranks_ok(err is None, levels[0].device):
                self.gather = gather
            else:
                import warnings
                warnings.warn(f'peer-memory gradient gather unavailable ({err!r}); all-reducing the image gradient')

    def evaluate(self, optimizing_img: torch.Tensor):
        """image leaf -> summed loss over the levels (neural_style_transfer.py:168-185), differentiable."""
        return PyramidFn.apply(self, optimizing_img)


class PyramidFn(torch.autograd.Function):
    """image -> total loss of the whole pyramid: chained bicubic 2x down (:168-176), every level's loss in
    lock-step, plain sum in level order (:180-185, previous_loss_importance = 1); backward adds the bicubic
    adjoint chain.  Per-level (total, content, style, tv) are kept in `last_out4` for verbose printing."""

    @staticmethod
    def forward(ctx, pyr: ShardedPyramid, img):
        ops._require_cuda(img)
        if pyr.gather is not None and ctx.needs_input_grad[1]:
            pyr.gather.announce()                  # my previous gradient has been consumed: peers may overwrite it
        imgs = [img.contiguous()]
        tvs = [None] * len(pyr.levels)
        for i in range(1, len(pyr.levels)):
            prev = imgs[-1]
            if FUSED_PYRAMID_TV and prev.shape[-2] % 2 == 0 and prev.shape[-1] % 2 == 0:
                nxt, sums2, tv = ops.bicubic_down_tv_raw(prev, pyr.levels[i - 1].wss)   # + total_variation(prev)
                tvs[i - 1] = (sums2, tv)
            else:
                nxt = ops.bicubic_down_raw(prev, prev.shape[-2] // 2, prev.shape[-1] // 2)
            imgs.append(nxt)
        out4s, state = pyramid_forward(pyr.levels, imgs, pyr.lanes, tvs, getattr(pyr, 'comm', None))
        total = out4s[0][0]
        for o in out4s[1:]:
            total = 1.0 * total + o[0]
        pyr.last_out4 = out4s
        if ctx.needs_input_grad[1]:
            ctx.state = state
            ctx.lanes = pyr.lanes
            ctx.gather = pyr.gather
            ctx.shapes = [tuple(t.shape[-2:]) for t in imgs]
        return total

    @staticmethod
    def backward(ctx, g_total):
        state, ctx.state = ctx.state, None
        d_imgs = pyramid_backward(state, g_total, ctx.lanes, getattr(ctx, 'gather', None))
        for i in range(len(d_imgs) - 1, 0, -1):          # adjoint of the down-sampling chain, coarse -> fine
            ops.bicubic_down_adj_raw(d_imgs[i], *ctx.shapes[i - 1], gx=d_imgs[i - 1], accumulate=True)
        g0 = d_imgs[0]
        if getattr(ctx, 'gather', None) is not None:
            g0 = g0.view(g0.shape)      # a fresh tensor object over the symmetric buffer: autograd can adopt it as .grad without a copy
        return None, g0


def pyramid_forward(levels: Sequence[ShardedPathLevel], imgs: Sequence[torch.Tensor], lanes: Lanes = _SERIAL,
                    tvs: Sequence = None, comm: 'torch.cuda.Stream' = None):
    """Forward schedule of several sharded levels in lock-step (one level = the plain sharded level).
    Returns ([out4 per level], state for pyramid_backward)."""
    dev = ops._require_cuda(*imgs)
    sh0 = levels[0]
    plan, grp = sh0.plan, sh0.group
    imgs = [t.contiguous() for t in imgs]
    for sh, im in zip(levels, imgs):
        if sh.plan is not plan or sh.group is not grp:
            raise ValueError('lock-step levels must share the feature plan and the process group')
        if tuple(im.shape) != (1, plan.steps[0][3], sh.H, sh.W):
            raise ValueError(f'sharded level expects {(1, plan.steps[0][3], sh.H, sh.W)}; got {tuple(im.shape)}')
        sh.generation += 1
        if not sh.hb:
            continue
        if sh.D != sh0.D:
            raise ValueError('lock-step levels must carry the same halo depth')
        # image band + D rows above / below straight from the replicated image (rows outside the image stay 0)
        lo, hi = max(sh.r0 - sh.D, 0), min(sh.r1 + sh.D, sh.H)
        c0 = im.shape[1]
        ops.chw_to_hwc(im, sh.xin, c0, (hi - lo) * sh.W, plane=sh.H * sh.W, x_off=lo * sh.W,
                       y_off=(lo - (sh.r0 - sh.D)) * sh.W * c0)
    D = sh0.D
    inner = lambda buf: _interior(buf, D)
    exchange_before, _ = par.halo_schedule([st[0] for st in plan.steps[:plan.n_steps_needed]], plan.taps_at, D)
    active = [li for li, sh in enumerate(levels) if sh.hb]       # the levels this rank owns rows of
    xs = [sh.xin for sh in levels]
    taps = [[None] * len(plan.tap_step) for _ in levels]
    # Raw partial Grams + partial content MSE of every level, TAP-MAJOR: [relu1_1 of all levels | relu2_1 ... | relu5_1
    # of all levels | content of all levels] (every slot 16-byte aligned).  The slots of one tap are one contiguous
    # block, all-reduced as soon as that tap's partial Grams exist — on the `comm` stream, under the convolutions of the
    # layers above it (OVERLAP_GRAM_ALLREDUCE); only the last block (relu5_1 + content, ~4 MB at L=3) is exposed.  Every
    # rank finalises identically afterwards.
    sidx0, cidx0 = sh0.sidx, sh0.cidx
    for sh in levels:
        if sh.sidx != sidx0 or sh.cidx != cidx0:
            raise ValueError('lock-step levels must tap the same layers')
    pad4 = lambda n: (n + 3) // 4 * 4
    tap_off, o = [], 0
    for j in range(len(sidx0)):
        tap_off.append(o)
        o += sum(pad4(sh.channels[j] ** 2) for sh in levels)
    content_off = o
    n_packed = o + pad4(len(levels))
    packed_all = torch.zeros(n_packed, dtype=torch.float32, device=dev)
    gram_slot, content_slot = [[None] * len(sidx0) for _ in levels], []
    for j in range(len(sidx0)):
        o = tap_off[j]
        for li, sh in enumerate(levels):
            c2 = sh.channels[j] ** 2
            gram_slot[li][j] = packed_all[o:o + c2]
            o += pad4(c2)
    for li in range(len(levels)):
        content_slot.append(packed_all[content_off + li:content_off + li + 1])
    overlap = comm is not None and OVERLAP_GRAM_ALLREDUCE
    main = torch.cuda.current_stream(dev)
    pending = []                                   # completion events of the all-reduces in flight on `comm`

    def all_reduce_block(lo, hi, last):
        """Sum packed_all[lo:hi] over the ranks; every rank calls this for the same blocks in the same order."""
        block = packed_all[lo:hi]
        with ops.timed(dev, ('allreduce_packed_grams', hi - lo)):
            if overlap and not last:
                ready = torch.cuda.Event()
                ready.record(main)
                comm.wait_event(ready)
                with torch.cuda.stream(comm):
                    grp.all_reduce_sum(block)
                    done = torch.cuda.Event()
                    done.record(comm)
                pending.append(done)
            else:
                grp.all_reduce_sum(block)

    reduced_to = 0                                 # packed_all[:reduced_to] has been handed to an all-reduce
    for sidx in range(plan.n_steps_needed):
        st = plan.steps[sidx]
        if sidx in exchange_before:
            with ops.timed(dev, ('halo_exchange_fwd', len(active), sidx)):
                par.halo_exchange(grp, [(_rows(xs[li]), levels[li].up, levels[li].dn, li) for li in active], depth=D)
        feeds_conv = sidx + 1 < plan.n_steps_needed and plan.steps[sidx + 1][0] == 'conv'

        def step(li, st=st, sidx=sidx, feeds_conv=feeds_conv):
            sh = levels[li]
            x, y = xs[li], sh.bufs[sidx]
            if st[0] == 'conv' and FUSED_BAND_CONV:
                y = sh.bufs[sidx] = _conv_relu_fwd_padded(x, st[1], st[2])
                if feeds_conv:                      # border halos are the next convolution's zero padding
                    rows = _rows(y)
                    if sh.up is None:
                        rows[:D].zero_()
                    if sh.dn is None:
                        rows[-D:].zero_()
            elif st[0] == 'conv':
                yi = inner(y)
                _conv_fwd_into(x, st[1], yi)
                ops.bias_relu_(yi, st[2])
            else:
                ops.maxpool2x2(inner(x), inner(y))
            for k in plan.taps_at.get(sidx, ()):
                tap = taps[li][k] = inner(y)
                if k in sh.sidx:                    # this band's partial Gram of the tap, straight into its slot
                    j = sh.sidx.index(k)
                    c, hw_band = tap.shape[1], tap.shape[2] * tap.shape[3]
                    ops.gram_mse_fwd_nhwc(tap, c, hw_band, 1.0, None, gram_slot[li][j], None,
                                          sh.wss.for_gram(j, c, hw_band, dev))
                if k == sh.cidx:
                    # partial content MSE already carries 1/numel of the WHOLE map: the reduced slot is the level's loss
                    ops.mse_fwd(tap, sh.target_content_band, 1.0 / sh.content_numel_global, content_slot[li],
                                sh.wss.for_reduce('content', dev))
            xs[li] = y

        lanes.each(active, step)
        # style taps are produced in slot order: everything up to the end of the last finished tap can be reduced now
        done_taps = [sidx0.index(k) for k in plan.taps_at.get(sidx, ()) if k in sidx0]
        if done_taps:
            j = max(done_taps)
            hi = tap_off[j + 1] if j + 1 < len(sidx0) else n_packed      # the last tap takes the content block along
            last = j + 1 == len(sidx0)
            if hi > reduced_to and (overlap or last):
                all_reduce_block(reduced_to, hi, last)
                reduced_to = hi
    if reduced_to < n_packed:                      # defensive: taps not in ascending slot order
        all_reduce_block(reduced_to, n_packed, True)
    for ev in pending:
        main.wait_event(ev)
    # every rank finalises identically: ALL Grams of ALL levels in one launch (D = G/(C HW) - A rounded to TF32 for the
    # backward's operand, per-layer MSE), then per level TV + the weighted sum
    items, ds_all, vals_all = [], [], []
    for li, sh in enumerate(levels):
        n_style = len(sh.sidx)
        vals = torch.empty(n_style + 1, dtype=torch.float32, device=dev)   # style mse[n] | tv
        ds = {}
        for j, k in enumerate(sh.sidx):
            c = sh.channels[j]
            st_ = par.LAYER_STRIDE[k]
            hw_global = (sh.H // st_) * (sh.W // st_)
            d = ops.new_d(c, sh.bf16, dev)
            items.append((gram_slot[li][j], c, 1.0 / (c * hw_global), sh.target_grams[j], d, vals[j],
                          ops.d_round_mode(c, sh.bf16)))
            ds[k] = (d, hw_global)
        ds_all.append(ds)
        vals_all.append(vals)
    sh0 = levels[0]
    fin_ws = getattr(sh0, 'fin_batch_ws', None)
    if fin_ws is None or fin_ws[0] < len(items):
        fin_ws = sh0.fin_batch_ws = (len(items), ops.finalize_batch_workspace(len(items), dev))
    for a in range(0, len(items), ops.L.AST_FINALIZE_MAX_ITEMS):
        ops.gram_finalize_batch(items[a:a + ops.L.AST_FINALIZE_MAX_ITEMS], fin_ws[1])
    out4s, per_level = [], []
    for li, sh in enumerate(levels):
        im, vals = imgs[li], vals_all[li]
        cw, sw, tvw = sh.weights
        n_style = len(sh.sidx)
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        if tvs is not None and tvs[li] is not None:        # the fused pyramid step already has total_variation(im)
            sums2, tv_val = tvs[li]
        else:
            sums2 = torch.empty(2, dtype=torch.float32, device=dev)
            tv_val = vals[n_style]
            ops.tv_fwd(im, sums2, tv_val, sh.wss.for_reduce('tv', dev))
        ops._launch(dev, ('combine',), 'ast_level_combine', vals.data_ptr(), n_style, content_slot[li].data_ptr(),
                    tv_val.data_ptr(), cw, sw, tvw, out4.data_ptr())
        out4s.append(out4)
        per_level.append((sh, sh.generation, ds_all[li], im, sums2))
    return out4s, per_level


def pyramid_backward(state, g_total, lanes: Lanes = _SERIAL, gather=None) -> List[torch.Tensor]:
    """Backward schedule in lock-step: this rank's contribution to every level image's gradient (its band rows,
    plus the TV gradient on rank 0).  g_total: upstream scalar gradient (device tensor) or None."""
    levels = [st[0] for st in state]
    sh0 = levels[0]
    plan, grp = sh0.plan, sh0.group
    dev = state[0][3].device
    gsc = ops._gscale(g_total, dev)
    for sh, generation, *_ in state:
        if generation != sh.generation:
            raise RuntimeError('sharded level: backward() after a newer forward() of the same level — its activation '
                               'bands are persistent buffers; run forward and backward of a closure back to back')

    D = sh0.D
    inner = lambda buf, ext=0: _interior(buf, D, ext)
    _, schedule = par.halo_schedule([st[0] for st in plan.steps[:plan.n_steps_needed]], plan.taps_at, D)

    def new_grad_band(like):
        """Padded gradient band for a padded activation band of the same shape."""
        return torch.empty(like.shape, dtype=torch.float32, device=dev, memory_format=_CL)

    def tap_grad(li, k, buf, gp, relu_mask, ext=0):
        """Gradient of tap k (the owned rows +- ext of the padded activation band `buf`) into the gradient band."""
        sh, _, ds, _, _ = state[li]
        cw, sw, tvw = sh.weights
        n = len(sh.sidx)
        acc = gp is not None
        if not acc:
            if ext:
                raise RuntimeError('the first tap gradient of a band has no valid halo rows to extend into')
            gp = new_grad_band(buf)
        tap, g = inner(buf, ext), inner(gp, ext)
        c, hw_band = tap.shape[1], tap.shape[2] * tap.shape[3]
        style, content = k in ds, k == sh.cidx
        fuse_gram = relu_mask and not content and c <= fp.FUSED_TAP_RELU_MAX_C
        if style:
            d, hw_global = ds[k]
            ops.gram_bwd_nhwc_auto(d, tap, c, hw_band, (sw / n) * 4.0 / (float(c) * c * c * hw_global), gsc, g, acc,
                                   relu_mask=fuse_gram)
        if content:
            ops.mse_bwd(tap, sh.content_target(ext), cw * 2.0 / sh.content_numel_global, gsc, g, acc or style,
                        relu_mask)
        elif not style and not acc:
            g.zero_()
        if relu_mask and not content and not fuse_gram:
            ops.relu_bwd_(g, tap)
        return gp

    d_imgs = []
    for li, (sh, _, _, im, sums2) in enumerate(state):
        if gather is not None:
            # peer-memory gather: this rank's rows go straight into the symmetric gradient buffer, the TV gradient of
            # those rows is added in place, and everybody's rows arrive there with gather.gather()
            d = gather.views[li]
        elif grp.rank == 0:                     # TV is replicated: count its gradient once
            d = torch.empty_like(im)
            ops.tv_bwd(im, sums2, sh.weights[2], gsc, d, False)
        else:
            d = torch.zeros_like(im)
        d_imgs.append(d)
    gps = [None] * len(levels)                  # padded gradient band w.r.t. the current step's output, per level
    flowing = False                             # a tap at or below this step has started the gradient
    masked = [False] * len(levels)
    for sidx in range(plan.n_steps_needed - 1, -1, -1):
        st = plan.steps[sidx]
        has_tap = bool(plan.taps_at.get(sidx))
        if not (has_tap or flowing):             # the same decision on every rank, whatever rows it owns
            continue
        flowing = True
        live = [li for li in range(len(levels)) if levels[li].hb]
        if st[0] == 'conv':
            how, ext = schedule[sidx]

            def pre(li, sidx=sidx, ext=ext):     # tap gradients into the band + the ReLU's backward (fused or in place)
                # ext = 1 (a step without an exchange): on the owned rows +- 1 — the gradient band is valid there
                # after the previous backward-data, and so is the forward band (it fed a convolution); rows outside
                # the image hold activation 0, so the mask zeroes their gradient
                taps_here = plan.taps_at.get(sidx, ())
                for n, k in enumerate(taps_here):
                    fuse = n == len(taps_here) - 1 and not masked[li]
                    gps[li] = tap_grad(li, k, levels[li].bufs[sidx], gps[li], fuse, ext)
                    masked[li] = masked[li] or fuse
                if not masked[li]:
                    ops.relu_bwd_(inner(gps[li], ext), inner(levels[li].bufs[sidx], ext))
                masked[li] = False

            lanes.each(live, pre)
            if how == 'exchange':
                with ops.timed(dev, ('halo_exchange_bwd', len(live), sidx)):
                    par.halo_exchange(grp, [(_rows(gps[li]), levels[li].up, levels[li].dn, li) for li in live],
                                      zero_border=True, depth=D)

            def dgrad(li, st=st, sidx=sidx):
                sh = levels[li]
                x = sh.bufs[sidx - 1] if sidx > 0 else sh.xin
                gxp = _conv_bwd_data_padded(gps[li], x, st[1])   # owned rows complete; its halo rows are not used
                if sidx > 0:
                    gps[li] = gxp
                else:
                    c0 = gxp.shape[1]
                    ops.hwc_to_chw(gxp, d_imgs[li], c0, sh.hb * sh.W, gather is None, plane=sh.H * sh.W,
                                   x_off=D * sh.W * c0, y_off=sh.r0 * sh.W)
                    if gather is not None:      # the TV gradient of MY rows travels with them (the image is replicated)
                        _, _, _, im, sums2 = state[li]
                        ops.tv_bwd(im, sums2, sh.weights[2], gsc, d_imgs[li], True, rows=(sh.r0, sh.r1))
                    gps[li] = None

            lanes.each(live, dgrad)
        else:
            fuse = sidx > 0 and plan.steps[sidx - 1][0] == 'conv' and (sidx - 1) not in plan.taps_at

            def pool(li, sidx=sidx, fuse=fuse):
                sh = levels[li]
                for k in plan.taps_at.get(sidx, ()):
                    gps[li] = tap_grad(li, k, sh.bufs[sidx], gps[li], False)
                xb = sh.bufs[sidx - 1] if sidx > 0 else sh.xin
                gxp = new_grad_band(xb)
                ops.maxpool2x2_bwd(inner(gps[li]), inner(xb), inner(gxp), fuse)
                masked[li] = fuse
                gps[li] = gxp

            lanes.each(live, pool)
    if gather is not None:
        with ops.timed(dev, ('gather_image_grad', len(d_imgs))):
            gather.gather()                     # every rank's rows of every level (TV gradient included), in place
    return d_imgs
