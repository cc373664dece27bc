"""CPU, world_size 2 over gloo: the host-side logic of the row-band sharding (artstyletransfer_b200/parallel.py)
— band plan, halo sufficiency, packed all-reduce layout, image-gradient sync — with the band arithmetic done by
plain torch on the oracle's VGG (the CUDA kernels are covered by tests/test_gpu_sharding.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from artstyletransfer_b200 import parallel
        from oracle import gatys_oracle as O
        torch.set_num_threads(2)
        parallel.init_sharding()
        assert parallel.world() == (rank, world)
        H, W = 384, 32
        net, cidx, sidx = O.make_vgg19(1234)
        g = torch.Generator().manual_seed(0)
        x = torch.rand((1, 3, H, W), generator=g) * 255 - 120
        tgt = torch.rand((1, 3, H, W), generator=g) * 255 - 120
        plan = parallel.BandPlan(H, rank, world)
        with torch.no_grad():
            full = net(x)
            band = net(x[:, :, plan.lo:plan.hi, :])
            tc_full = net(tgt)[cidx][0]
        chans = [full[k].shape[1] for k in sidx]
        offs, slot, n = parallel.pack_layout(chans)
        packed = torch.zeros(n)
        for i, k in enumerate(sidx):
            a, b, rows = plan.feat_rows(parallel.LAYER_STRIDE[k])
            assert rows == band[k].shape[2]
            fb = band[k][0, :, a:b, :].reshape(chans[i], -1).double()
            packed[offs[i]:offs[i] + chans[i] ** 2] = (fb @ fb.t()).float().reshape(-1)
        ca, cb, _ = plan.feat_rows(parallel.LAYER_STRIDE[cidx])
        ga, gb = plan.global_feat_rows(parallel.LAYER_STRIDE[cidx])
        packed[slot] = ((band[cidx][0, :, ca:cb, :] - tc_full[:, ga:gb, :]).double() ** 2).sum().float()
        parallel._GROUP.all_reduce_sum(packed)
        errs = []
        for i, k in enumerate(sidx):
            f = full[k][0].reshape(chans[i], -1).double()
            ref = (f @ f.t())
            got = packed[offs[i]:offs[i] + chans[i] ** 2].reshape(chans[i], chans[i]).double()
            errs.append(float((got - ref).norm() / ref.norm()))
        sse_ref = float(((full[cidx][0] - tc_full).double() ** 2).sum())
        errs.append(abs(float(packed[slot]) - sse_ref) / sse_ref)
        img = torch.zeros(1, 3, 8, 8, requires_grad=True)
        if rank == 0:
            img.grad = torch.full_like(img, 1.0)       # rank 1 has no gradient yet: must still join
        parallel.sync_image_grad(img)
        out[rank] = (errs, float(img.grad.mean()))
    finally:
        dist.destroy_process_group()


def test_band_sums_equal_full_grams_over_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]
    for rank in range(world):
        errs, gmean = out[rank]
        assert max(errs) < 1e-5, errs          # halo cropping is exact: only fp32 summation-order noise remains
        assert gmean == 1.0
    assert out[0][0] == out[1][0]              # all ranks hold identical reduced values


def _halo_worker(rank, world, port, out, edges=None):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import torch.nn.functional as F
        from artstyletransfer_b200 import parallel
        torch.set_num_threads(2)
        parallel.init_sharding()
        grp = parallel._GROUP
        H, W, C1, C2 = 24, 10, 4, 6
        if edges is None:                          # equal bands, neighbours rank -+ 1 (halo_exchange's default)
            edges = [r * (H // world) for r in range(world + 1)]
            nb = None
        else:                                      # explicit row edges; a rank may own nothing (parallel.PyramidBands)
            owners = [r for r in range(world) if edges[r + 1] > edges[r]]
            i = owners.index(rank) if rank in owners else None
            nb = (None, None) if i is None else (owners[i - 1] if i > 0 else None,
                                                 owners[i + 1] if i + 1 < len(owners) else None)
        r0, r1 = edges[rank], edges[rank + 1]
        hb = r1 - r0

        def exchange(t, **kw):
            if nb is None:
                parallel.halo_exchange(grp, t, **kw)
            else:
                parallel.halo_exchange(grp, [(t, nb[0], nb[1])] if hb else [], **kw)
        g = torch.Generator().manual_seed(0)
        x = torch.randn((1, C1, H, W), generator=g, dtype=torch.float64)
        w1 = torch.randn((C2, C1, 3, 3), generator=g, dtype=torch.float64)
        w2 = torch.randn((C1, C2, 3, 3), generator=g, dtype=torch.float64)
        gout = torch.randn((1, C1, H, W), generator=g, dtype=torch.float64)
        # reference: two stacked 3x3 convolutions on the whole image, gradient of <y, gout> w.r.t. x
        xr = x.clone().requires_grad_(True)
        yr = F.conv2d(F.conv2d(xr, w1, padding=1), w2, padding=1)
        (gref,) = torch.autograd.grad(yr, xr, gout)
        # sharded: each rank owns hb rows; padded NHWC bands, one halo row exchanged per convolution

        def padded(c):
            return torch.zeros((hb + 2, W, c), dtype=torch.float64)

        def conv_band(pad_rows, wgt):      # (hb+2, W, Cin) -> (hb, W, Cout), zero padding left/right only
            xin = pad_rows.permute(2, 0, 1)[None]
            return F.conv2d(xin, wgt, padding=(0, 1))[0].permute(1, 2, 0).contiguous()

        if not hb:                                 # a rank without rows still walks through the four exchanges
            for _ in range(4):
                exchange(None)
            out[rank] = (0.0, 0.0)
            return
        a0 = padded(C1)
        a0[1:-1] = x[0, :, r0:r1].permute(1, 2, 0)
        exchange(a0)
        a1 = padded(C2)
        a1[1:-1] = conv_band(a0, w1)
        exchange(a1)
        y = conv_band(a1, w2)
        fwd_err = float((y - yr[0, :, r0:r1].permute(1, 2, 0)).abs().max())

        def conv_band_bwd(gpad, wgt, cin):
            """Symmetric-padding backward-data over the whole padded gradient band (hb+2 rows in and out); with the
            neighbours' edge gradient rows in the halos its owned rows are complete."""
            xin = torch.zeros((1, cin, hb + 2, W), dtype=torch.float64, requires_grad=True)
            yy = F.conv2d(xin, wgt, padding=1)
            (gx,) = torch.autograd.grad(yy, xin, gpad.permute(2, 0, 1)[None])
            return gx[0].permute(1, 2, 0).contiguous()

        g1 = torch.full((hb + 2, W, C1), float('nan'), dtype=torch.float64)      # recycled buffer: halos are garbage
        g1[1:-1] = gout[0, :, r0:r1].permute(1, 2, 0)
        exchange(g1, zero_border=True)
        g0 = conv_band_bwd(g1, w2, C2)            # gradient w.r.t. the first convolution's output band (padded)
        exchange(g0, zero_border=True)
        gx = conv_band_bwd(g0, w1, C1)
        bwd_err = float((gx[1:-1] - gref[0, :, r0:r1].permute(1, 2, 0)).abs().max())
        out[rank] = (fwd_err, bwd_err)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_halo_exchange_reproduces_full_convolutions_over_gloo(world):
    """Host logic of the per-layer halo exchange (parallel.halo_exchange, activations and gradients) with real send/recv between
    processes: two stacked 3x3 convolutions on row bands == the same convolutions on the whole image, forward and
    backward (fp64, exact up to rounding)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_halo_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert sorted(out.keys()) == list(range(world))
    for rank in range(world):
        fwd_err, bwd_err = out[rank]
        assert fwd_err < 1e-10 and bwd_err < 1e-10, (rank, fwd_err, bwd_err)


def _halo2_worker(rank, world, port, out):
    """Two halo rows per exchange (parallel.halo_schedule, depth 2): ONE exchange serves two stacked convolutions, in
    the forward and in the backward — the scheme of sharded_path.pyramid_forward / pyramid_backward in fp64."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import torch.nn.functional as F
        from artstyletransfer_b200 import parallel
        torch.set_num_threads(2)
        parallel.init_sharding()
        grp = parallel._GROUP
        H, W, C1, C2, D = 24, 10, 4, 6, 2
        r0, r1 = rank * (H // world), (rank + 1) * (H // world)
        hb = r1 - r0
        top, bottom = rank == 0, rank == world - 1
        g = torch.Generator().manual_seed(0)
        x = torch.randn((1, C1, H, W), generator=g, dtype=torch.float64)
        w1 = torch.randn((C2, C1, 3, 3), generator=g, dtype=torch.float64)
        w2 = torch.randn((C1, C2, 3, 3), generator=g, dtype=torch.float64)
        gout = torch.randn((1, C1, H, W), generator=g, dtype=torch.float64)
        xr = x.clone().requires_grad_(True)
        yr = F.conv2d(F.conv2d(xr, w1, padding=1), w2, padding=1)
        (gref,) = torch.autograd.grad(yr, xr, gout)

        def whole_band_conv(rows, wgt):            # (hb + 4, W, Cin) -> (hb + 4, W, Cout), symmetric zero padding
            return F.conv2d(rows.permute(2, 0, 1)[None], wgt, padding=1)[0].permute(1, 2, 0).contiguous()

        def whole_band_bwd(grows, wgt, cin):
            xin = torch.zeros((1, cin, hb + 2 * D, W), dtype=torch.float64, requires_grad=True)
            (gx,) = torch.autograd.grad(F.conv2d(xin, wgt, padding=1), xin, grows.permute(2, 0, 1)[None])
            return gx[0].permute(1, 2, 0).contiguous()

        def zero_outside_image(rows):              # the next convolution's zero padding / no gradient out there
            if top:
                rows[:D].zero_()
            if bottom:
                rows[-D:].zero_()

        a0 = torch.zeros((hb + 2 * D, W, C1), dtype=torch.float64)
        a0[D:-D] = x[0, :, r0:r1].permute(1, 2, 0)
        parallel.halo_exchange(grp, a0, depth=D)                  # v = 2
        a1 = whole_band_conv(a0, w1)                              # v = 1: the outermost rows are junk
        zero_outside_image(a1)
        y = whole_band_conv(a1, w2)                               # v = 0: only the owned rows are right
        fwd_err = float((y[D:-D] - yr[0, :, r0:r1].permute(1, 2, 0)).abs().max())
        g1 = torch.full((hb + 2 * D, W, C1), float('nan'), dtype=torch.float64)
        g1[D:-D] = gout[0, :, r0:r1].permute(1, 2, 0)
        parallel.halo_exchange(grp, g1, zero_border=True, depth=D)   # u = 2
        g0 = whole_band_bwd(g1, w2, C2)                           # u = 1
        zero_outside_image(g0)                                    # (the ReLU mask does this on the product path)
        gx = whole_band_bwd(g0, w1, C1)                           # u = 0
        bwd_err = float((gx[D:-D] - gref[0, :, r0:r1].permute(1, 2, 0)).abs().max())
        out[rank] = (fwd_err, bwd_err)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_two_row_halos_serve_two_convolutions_over_gloo(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_halo2_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert sorted(out.keys()) == list(range(world))
    for rank in range(world):
        fwd_err, bwd_err = out[rank]
        assert fwd_err < 1e-10 and bwd_err < 1e-10, (rank, fwd_err, bwd_err)


def _band_schedule_worker(rank, world, port, out, depth):
    """The whole band scheme of sharded_path.pyramid_forward / pyramid_backward restated with torch CPU ops in
    float64 and run between gloo processes: conv+bias+ReLU over the whole padded band, border zeroing, max-pool of the
    owned rows, taps, and the backward with tap gradients and ReLU masks on the owned rows +- ext, all driven by
    parallel.halo_schedule.  Compared with autograd on the whole image."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import torch.nn.functional as F
        from artstyletransfer_b200 import parallel
        torch.set_num_threads(2)
        parallel.init_sharding()
        grp = parallel._GROUP
        D = depth
        H, W = 16 * world, 10
        kinds = ['conv', 'conv', 'pool', 'conv', 'conv']
        chans = [3, 4, 5, 5, 6, 4]                       # channels of the input and of every step's output
        taps = {0: 0.7, 3: -1.3, 4: 0.9}                   # step -> weight of the tap loss 0.5 * w * sum(a_c * y^2)
        gen = torch.Generator().manual_seed(5)
        rnd = lambda *shape: torch.randn(shape, generator=gen, dtype=torch.float64)
        img = rnd(1, chans[0], H, W)
        ws = {s: (rnd(chans[s + 1], chans[s], 3, 3) * 0.4, rnd(chans[s + 1]) * 0.5 + 0.3)
              for s, k in enumerate(kinds) if k == 'conv'}
        a = {s: rnd(chans[s + 1]).abs() + 0.1 for s in taps}

        # ---- reference: the whole image through autograd ------------------------------------------------------
        xr = img.clone().requires_grad_(True)
        y, loss_ref = xr, 0.0
        for s, k in enumerate(kinds):
            y = F.relu(F.conv2d(y, *ws[s], padding=1)) if k == 'conv' else F.max_pool2d(y, 2)
            if s in taps:
                loss_ref = loss_ref + 0.5 * taps[s] * (a[s][None, :, None, None] * y * y).sum()
        (gref,) = torch.autograd.grad(loss_ref, xr)

        # ---- sharded ------------------------------------------------------------------------------------------
        fwd_x, bwd_plan = parallel.halo_schedule(kinds, taps, D)
        top, bottom = rank == 0, rank == world - 1
        hb = H // world
        r0 = rank * hb
        nchw = lambda rows: rows.permute(2, 0, 1)[None]     # (h, W, C) band rows <-> (1, C, h, W)
        rows_of = lambda t: t[0].permute(1, 2, 0).contiguous()

        def interior(rows, ext=0):
            return rows[D - ext:rows.shape[0] - D + ext]

        x = torch.zeros((hb + 2 * D, W, chans[0]), dtype=torch.float64)
        lo, hi = max(r0 - D, 0), min(r0 + hb + D, H)
        x[lo - (r0 - D):lo - (r0 - D) + hi - lo] = rows_of(img[:, :, lo:hi])
        bufs, loss_part, n_fwd = [], torch.zeros(1, dtype=torch.float64), 0
        for s, k in enumerate(kinds):
            if s in fwd_x:
                parallel.halo_exchange(grp, x, depth=D)
                n_fwd += 1
            if k == 'conv':
                yb = rows_of(F.relu(F.conv2d(nchw(x), *ws[s], padding=1)))       # the whole padded band
                if s + 1 < len(kinds) and kinds[s + 1] == 'conv':               # border halos = the next zero padding
                    if top:
                        yb[:D] = 0
                    if bottom:
                        yb[-D:] = 0
            else:
                h2 = (x.shape[0] - 2 * D) // 2
                yb = torch.full((h2 + 2 * D, W // 2, x.shape[2]), 1e6, dtype=torch.float64)   # halos: junk until exchanged
                if top:
                    yb[:D] = 0
                if bottom:
                    yb[-D:] = 0
                yb[D:-D] = rows_of(F.max_pool2d(nchw(interior(x)), 2))
            if s in taps:
                yi = interior(yb)
                loss_part += 0.5 * taps[s] * (a[s] * yi * yi).sum()
            bufs.append(yb)
            x = yb
        dist.all_reduce(loss_part)
        # backward
        g, n_bwd, xin = None, 0, None
        inputs = [None] + bufs[:-1]
        for s in range(len(kinds) - 1, -1, -1):
            if s > max(taps):
                continue
            yb = bufs[s]
            if kinds[s] == 'conv':
                how, ext = bwd_plan[s]
                if s in taps:
                    if g is None:
                        g = torch.full(yb.shape, 1e6, dtype=torch.float64)        # recycled buffer: halos are junk
                        interior(g)[:] = 0
                    interior(g, ext)[:] += taps[s] * a[s] * interior(yb, ext)
                interior(g, ext)[:] *= (interior(yb, ext) > 0)                    # ReLU backward, also on the ext rows
                if how == 'exchange':
                    parallel.halo_exchange(grp, g, zero_border=True, depth=D)
                    n_bwd += 1
                x_shape = (1, chans[s], yb.shape[0], yb.shape[1])
                g = rows_of(torch.nn.grad.conv2d_input(x_shape, ws[s][0], nchw(g).contiguous(), padding=1))
            else:
                xb = inputs[s]
                if s in taps:
                    raise AssertionError('taps are ReLU outputs')
                xi = nchw(interior(xb)).clone().requires_grad_(True)
                (gi,) = torch.autograd.grad(F.max_pool2d(xi, 2), xi, nchw(interior(g)).contiguous())
                g = torch.full(xb.shape, 1e6, dtype=torch.float64)
                interior(g)[:] = rows_of(gi)
        mine = nchw(interior(g))
        want = gref[:, :, r0:r0 + hb]
        loss_ref = float(loss_ref.detach())
        out[rank] = (abs(float(loss_part) - loss_ref) / abs(loss_ref),
                     float((mine - want).abs().max() / gref.abs().max()), n_fwd, n_bwd)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,depth', [(2, 2), (3, 2), (3, 1)])
def test_band_scheme_with_relu_pool_and_taps_matches_autograd_over_gloo(world, depth):
    """Algorithm-level CPU check of the sharded path's band logic (the CUDA path is held to the same statement on the
    GPU by tests/test_gpu_sharding.py): loss and image gradient equal whole-image autograd to float64 rounding, with
    3 + 4 exchanges at depth 1 and 1 + 2 at depth 2 for this five-step network."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_band_schedule_worker, args=(world, _free_port(), out, depth), nprocs=world, join=True)
    assert sorted(out.keys()) == list(range(world))
    for rank in range(world):
        lerr, gerr, n_fwd, n_bwd = out[rank]
        assert lerr < 1e-12 and gerr < 1e-12, (rank, lerr, gerr)
        assert (n_fwd, n_bwd) == ((1, 2) if depth == 2 else (3, 4))


def test_halo_schedule_of_the_vgg19_path():
    """parallel.halo_schedule: with one halo row every convolution but the first waits for its neighbours (12 forward +
    13 backward exchanges up to conv5_1); with two rows every second one does (6 + 7), and the steps in between run
    their tap gradients on the owned rows +- 1."""
    from artstyletransfer_b200.parallel import PyramidBands, halo_schedule
    kinds = ['conv', 'conv', 'pool', 'conv', 'conv', 'pool', 'conv', 'conv', 'conv', 'conv', 'pool',
             'conv', 'conv', 'conv', 'conv', 'pool', 'conv']
    taps = {0: (0,), 3: (1,), 6: (2,), 11: (3,), 12: ('content',), 16: (4,)}
    convs = [s for s, k in enumerate(kinds) if k == 'conv']
    f1, b1 = halo_schedule(kinds, taps, 1)
    assert f1 == set(convs[1:]) and all(b1[s] == ('exchange', 0) for s in convs) and len(b1) == 13
    f2, b2 = halo_schedule(kinds, taps, 2)
    assert f2 == {3, 6, 8, 11, 13, 16}                                          # c2_1 c3_1 c3_3 c4_1 c4_3 c5_1
    assert sorted(s for s in convs if b2[s][0] == 'exchange') == [1, 4, 7, 9, 12, 14, 16]
    assert all(b2[s] == ('local', 1) for s in (0, 3, 6, 8, 11, 13))
    # every step that runs locally in the backward reads a forward band that fed a convolution (its halo row is valid)
    assert all(kinds[s + 1] == 'conv' for s, (how, _) in b2.items() if how == 'local')
    # taps deeper than the path: nothing flows above the deepest tap
    _, b = halo_schedule(kinds, {3: (1,)}, 2)
    assert sorted(b) == [0, 1, 3] and b[3] == ('exchange', 0) and b[1] == ('exchange', 0) and b[0] == ('local', 1)
    with pytest.raises(ValueError):
        halo_schedule(kinds, taps, 3)
    # the depth a plan can carry: every band must own `depth` rows at stride 16
    assert PyramidBands([(2048, 3072), (1024, 1536), (512, 768), (256, 384)], 8).halo_depth(2) == 2
    p = PyramidBands([(128, 64)], 4, uniform=True)
    assert p.halo_depth(2) == 2 and p.halo_depth(1) == 1
    p.bounds = [[0, 16, 48, 96, 128]]                  # a hand-made plan with a 16-row band: one row at stride 16
    assert p.halo_depth(2) == 1


def test_halo_schedule_keeps_every_read_row_valid_for_any_network():
    """Property test of parallel.halo_schedule on random conv / pool sequences with random taps: replay the validity
    bookkeeping the docstring describes and check that (forward) no convolution ever reads an invalid halo row into an
    owned row, (backward) every step that runs without an exchange finds both the gradient band and the forward band
    valid on the rows it touches, and that depth 2 never exchanges more often than depth 1."""
    from hypothesis import given, settings, strategies as st
    from artstyletransfer_b200.parallel import halo_schedule

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.sampled_from(['conv', 'conv', 'conv', 'pool']), min_size=1, max_size=24), st.data())
    def check(kinds, data):
        kinds = ['conv'] + kinds                                   # the path starts with a convolution of the image
        taps = set(data.draw(st.lists(st.integers(0, len(kinds) - 1), min_size=1, max_size=6)))
        taps = {s for s in taps if kinds[s] == 'conv'} or {0}       # taps are ReLU outputs
        counts = {}
        for depth in (1, 2):
            fwd, bwd = halo_schedule(kinds, {s: (0,) for s in taps}, depth)
            # forward replay: valid[s] = valid halo rows of step s's OUTPUT band when the backward reads it
            v, valid = depth, {}
            for s, kind in enumerate(kinds):
                if kind == 'conv':
                    if s in fwd:
                        assert s > 0 and v < 1
                        valid[s - 1] = depth                        # the exchange overwrote the input band's halos
                        v = depth
                    assert v >= 1, (kinds, s)                       # the owned rows' neighbours are real rows
                    v -= 1
                else:
                    assert s not in fwd
                    v = 0
                valid[s] = v
            # backward replay
            u, flowing, deepest = 0, False, max(taps)
            for s in range(len(kinds) - 1, -1, -1):
                if s > deepest:
                    assert s not in bwd
                    continue
                if kinds[s] == 'conv':
                    how, ext = bwd[s]
                    if how == 'exchange':
                        assert ext == 0 and u < 1
                        u = depth
                    else:
                        assert 1 <= ext <= u and valid[s] >= ext, (kinds, taps, s)   # gradient AND activation rows exist
                        assert s < deepest                          # the first tap gradient has no halo to extend into
                    u -= 1
                else:
                    assert s not in bwd
                    u = 0
            assert set(bwd) == {s for s in range(deepest + 1) if kinds[s] == 'conv'}
            counts[depth] = (len(fwd), sum(1 for h, _ in bwd.values() if h == 'exchange'))
        assert counts[2][0] <= counts[1][0] and counts[2][1] <= counts[1][1]

    check()


@pytest.mark.parametrize('edges', [[0, 16, 24], [0, 8, 8, 24], [0, 0, 10, 24]])
def test_halo_exchange_with_unequal_and_empty_bands_over_gloo(edges):
    """The level-aware plan (parallel.PyramidBands): bands of different heights, and ranks that own no rows of a
    level — their neighbours exchange across them, they join every step with an empty list."""
    world = len(edges) - 1
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_halo_worker, args=(world, _free_port(), out, edges), nprocs=world, join=True)
    assert sorted(out.keys()) == list(range(world))
    for rank in range(world):
        fwd_err, bwd_err = out[rank]
        assert fwd_err < 1e-10 and bwd_err < 1e-10, (rank, fwd_err, bwd_err)


def test_pyramid_bands_plan():
    """Host logic of the level-aware band plan: every row of every level owned exactly once, 16-row edges, no band
    under 32 rows, neighbours skip ranks without rows, and the modelled load balance at BASELINE's L=3 sizes."""
    from artstyletransfer_b200.parallel import ALIGN, MIN_BAND_ROWS, PyramidBands
    sizes = [(2048 >> i, 3072 >> i) for i in range(4)]
    for world in (1, 2, 3, 4, 5, 8):
        p = PyramidBands(sizes, world)
        for li, (h, _) in enumerate(sizes):
            edges = p.bounds[li]
            assert len(edges) == world + 1 and edges[0] == 0 and edges[-1] == h
            owners = []
            for r in range(world):
                r0, r1 = p.band(li, r)
                assert r0 <= r1 and r0 % ALIGN == 0 and r1 % ALIGN == 0
                assert r1 == r0 or r1 - r0 >= MIN_BAND_ROWS
                if r1 > r0:
                    owners.append(r)
            for i, r in enumerate(owners):
                assert p.neighbours(li, r) == (owners[i - 1] if i else None, owners[i + 1] if i + 1 < len(owners) else None)
            for r in set(range(world)) - set(owners):
                assert p.neighbours(li, r) == (None, None)
        one = sum(PyramidBands(sizes, 1).loads())
        assert one / max(p.loads()) > 0.85 * world, (world, p.describe())       # modelled strong-scaling efficiency
        assert PyramidBands(sizes, world).bounds == p.bounds                   # deterministic: every rank plans alike
    p8 = PyramidBands(sizes, 8)
    assert sum(1 for r in range(8) if p8.band(0, r)[1] > p8.band(0, r)[0]) == 6   # the 2048x3072 level on six ranks
    assert p8.band(3, 7) == (0, 256)                                           # the 256x384 level whole on the last
    u = PyramidBands(sizes, 8, uniform=True)
    assert [u.band(li, 3) for li in range(4)] == [(768, 1024), (384, 512), (192, 256), (96, 128)]
    assert PyramidBands([(256, 64)], 8).bounds == [[32 * r for r in range(9)]]
    with pytest.raises(ValueError):
        PyramidBands([(250, 64)], 2)
    with pytest.raises(ValueError):
        PyramidBands([(96, 64)], 4, uniform=True)


def test_init_sharding_requires_a_group():
    from artstyletransfer_b200 import parallel
    parallel.disable_sharding()
    with pytest.raises(RuntimeError, match='process group'):
        parallel.init_sharding()
    assert parallel.world() == (0, 1)


def test_pyramid_bands_random_pyramids():
    """Property test of the level-aware plan over random pyramids and world sizes: complete, aligned, no slivers,
    never worse (in the model) than cutting every level into equal bands when that is possible."""
    from hypothesis import given, settings, strategies as st
    from artstyletransfer_b200.parallel import ALIGN, MIN_BAND_ROWS, BandPlan, PyramidBands

    @settings(max_examples=150, deadline=None)
    @given(st.integers(2, 40), st.integers(1, 24), st.integers(1, 4), st.integers(1, 8),
           st.sampled_from([0.0, 0.015, 0.05]))
    def check(h16, w16, n_levels, world, overhead):
        h0, w0 = h16 * ALIGN << (n_levels - 1), w16 * ALIGN << (n_levels - 1)
        sizes = [(h0 >> i, w0 >> i) for i in range(n_levels)]
        p = PyramidBands(sizes, world, overhead=overhead)
        for li, (h, _) in enumerate(sizes):
            edges = p.bounds[li]
            assert edges[0] == 0 and edges[-1] == h and len(edges) == world + 1
            for r in range(world):
                rows = edges[r + 1] - edges[r]
                assert rows >= 0 and edges[r] % ALIGN == 0 and (rows == 0 or rows >= MIN_BAND_ROWS)
        if all(BandPlan.shardable(h, world) for h, _ in sizes):
            u = PyramidBands(sizes, world, overhead=overhead, uniform=True)
            assert max(p.loads()) <= max(u.loads()) + 1e-9, (sizes, world, p.describe())

    check()


def test_peer_halo_group_pairs_slots_and_counters(monkeypatch):
    """Host logic of parallel.PeerHaloGroup.exchange_rows (no GPU, no symmetric memory: fake peer pointers): what rank
    A pushes for its lower neighbour B lands in exactly the slot pair / counter B waits on, and vice versa, per
    level; border halos become zero-fill rows; one step is one launch (more than 16 rows are refused)."""
    from artstyletransfer_b200 import _lib as L, ops, parallel
    launched = {}

    def fake_launch(dev, key, name, arr, n):
        assert name == 'ast_halo_exchange' and 1 <= n <= L.AST_HALO_MAX_ROWS
        launched.setdefault('rows', []).extend(
            {f: getattr(arr[i], f) for f, _ in L.HaloRow._fields_} for i in range(n))
        launched['calls'] = launched.get('calls', 0) + 1

    monkeypatch.setattr(ops, '_launch', fake_launch)
    base = {0: 0x10000000, 1: 0x20000000, 2: 0x30000000}
    slot, flags_bytes = 4096, 256 * 16

    def group(rank):
        g = object.__new__(parallel.PeerHaloGroup)
        g.rank, g.world, g._hdl = rank, 3, object()
        g._ptrs, g._slot, g._flags_bytes = [base[0], base[1], base[2]], slot, flags_bytes
        g._state = torch.zeros(16 * 64, dtype=torch.int32)
        return g

    def rows_of(rank, entries, zero_border):
        launched.clear()
        parallel.halo_exchange(group(rank), entries, zero_border=zero_border)
        return launched['rows'], launched['calls']

    band = {r: [torch.zeros(6, 8, 4) for _ in range(2)] for r in range(3)}       # two levels per rank
    # level 0 lives on ranks 0, 1; level 1 on ranks 1, 2 (rank 1 owns rows of both: up = 0 at level 0, dn = 2 at level 1)
    a, _ = rows_of(0, [(band[0][0], None, 1, 0)], True)
    b, _ = rows_of(1, [(band[1][0], 0, None, 0), (band[1][1], None, 2, 1)], True)
    c, _ = rows_of(2, [(band[2][1], 1, None, 1)], False)
    assert [bool(r['src']) for r in a] == [False, True] and a[0]['halo'] == band[0][0][0].data_ptr()
    a_dn = a[1]
    b_up, b_l0_border, b_l1_border, b_dn = b
    assert not b_l0_border['src'] and not b_l1_border['src']
    assert a_dn['src'] == band[0][0][4].data_ptr() and a_dn['halo'] == band[0][0][5].data_ptr()
    assert a_dn['bytes'] == 8 * 4 * 4 and a_dn['slot_stride'] == slot
    # A -> B at level 0: A writes B's "from above" pair and counter, B reads exactly those; and the reverse
    assert a_dn['dst_remote'] == b_up['stage'] and a_dn['flag_remote'] == b_up['flag_local']
    assert b_up['dst_remote'] == a_dn['stage'] and b_up['flag_remote'] == a_dn['flag_local']
    assert base[1] <= b_up['stage'] < base[1] + flags_bytes + 32 * slot and base[0] <= a_dn['stage'] < base[1]
    # B -> C at level 1 uses other slots than level 0
    (c_up,) = c
    assert b_dn['dst_remote'] == c_up['stage'] and c_up['dst_remote'] == b_dn['stage']
    assert b_dn['flag_remote'] == c_up['flag_local'] and c_up['flag_remote'] == b_dn['flag_local']
    assert len({b_up['stage'], b_dn['stage']}) == 2 and len({b_up['state'], b_dn['state']}) == 2
    assert abs(b_up['stage'] - b_dn['stage']) >= 2 * slot
    # a step never spans two launches (neighbours could wait on each other's second launch): 18 rows are refused
    with pytest.raises(ValueError, match='halo rows in one step'):
        rows_of(1, [(torch.zeros(4, 2, 4), 0, 2, lv % 8) for lv in range(9)], False)
    rows, calls = rows_of(1, [(torch.zeros(4, 2, 4), 0, 2, lv) for lv in range(8)], False)
    assert len(rows) == 16 and calls == 1
    with pytest.raises(ValueError, match='exceeds the staging slot'):
        rows_of(1, [(torch.zeros(4, 64, 32), 0, 2, 0)], False)
    # two halo rows per exchange: the two edge rows travel as ONE contiguous piece of twice the bytes
    wide = torch.zeros(10, 8, 4)                                  # 6 owned rows + 2 halo rows per side
    launched.clear()
    parallel.halo_exchange(group(1), [(wide, 0, 2, 0)], zero_border=True, depth=2)
    up2, dn2 = launched['rows']
    assert up2['src'] == wide[2].data_ptr() and up2['halo'] == wide[0].data_ptr() and up2['bytes'] == 2 * 8 * 4 * 4
    assert dn2['src'] == wide[6].data_ptr() and dn2['halo'] == wide[8].data_ptr() and dn2['bytes'] == 2 * 8 * 4 * 4
    with pytest.raises(ValueError, match='cannot hand 2 rows'):
        parallel.halo_exchange(group(1), [(torch.zeros(5, 8, 4), 0, 2, 0)], depth=2)


@pytest.mark.parametrize('world', [2, 3, 4, 8])
def test_grad_gather_segments_tile_the_gradient_pyramid(world):
    """PeerGradGather's host logic: over all ranks, the segments every rank pushes into its peers' symmetric buffers
    cover every byte of every level's (3, H, W) gradient exactly once (disjoint rows -> a gather, not a reduction)."""
    import numpy as np
    from artstyletransfer_b200 import parallel
    sizes = [(2048 >> i, 3072 >> i) for i in range(4)]
    plan = parallel.PyramidBands(sizes, world)
    offs, flags_off, _ = parallel.grad_gather_layout(sizes, [(0, 0)] * 4)
    cover = np.zeros(flags_off // 16, dtype=np.int32)
    for rank in range(world):
        _, _, segs = parallel.grad_gather_layout(sizes, [plan.band(i, rank) for i in range(4)])
        assert len(segs) <= 24
        for off, nbytes in segs:
            assert off % 16 == 0 and nbytes % 16 == 0
            cover[off // 16:(off + nbytes) // 16] += 1
    for i, (h, w) in enumerate(sizes):
        lvl = cover[offs[i] // 16:(offs[i] + 3 * h * w * 4) // 16]
        assert lvl.min() == 1 and lvl.max() == 1
    assert cover.sum() == sum(3 * h * w * 4 for h, w in sizes) // 16
