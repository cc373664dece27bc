"""GPU: the row-band sharded level (artstyletransfer_b200/parallel.py) on ONE device — R ranks emulated as R host
threads sharing the default stream, with an in-process collective — vs the unsharded level.  Checks the pitched
partial-Gram loads, the packed all-reduce layout, ast_gram_finalize, halo cropping and the gradient sum."""
import threading

import numpy as np
import pytest
import torch

from oracle import gatys_oracle as O

pytestmark = pytest.mark.gpu
WEIGHTS = (1e3, 4e5, 1e2)


def dev():
    return torch.device('cuda', 0)


def root_causes(errors):
    """The emulated ranks' failures with the aborted-barrier followers last (they only echo the first failure)."""
    errors = sorted((str(e) for e in errors), key=lambda e: 'BrokenBarrierError' in e)
    for e in errors[:2]:
        print(e)                     # pytest's assertion repr truncates long strings; the captured stdout does not
    return errors[:2]


class ThreadGroup:
    """all_reduce_sum across `world` threads of this process (fixed rank order -> deterministic)."""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self.shared = rank, world, shared

    def all_reduce_sum(self, t):
        sh = self.shared
        # the product may call this on a side stream (overlapped partial-Gram all-reduce): the emulation reads the other
        # threads' buffers from THIS thread's stream, so every thread first waits for its own producers on the host
        torch.cuda.synchronize()
        sh['bufs'][self.rank] = t
        sh['barrier'].wait()
        total = sh['bufs'][0].clone()
        for r in range(1, self.world):
            total = total + sh['bufs'][r]
        sh['barrier'].wait()
        t.copy_(total)
        torch.cuda.synchronize()
        sh['barrier'].wait()

    def exchange(self, sends, recvs):
        """In-process stand-in for the grouped NCCL send/recv: post, barrier, copy, barrier (all threads enqueue on
        the same stream, so the copies are ordered after the producers' kernels)."""
        sh = self.shared
        for peer in {p for _, p in sends}:
            sh['mail'][(self.rank, peer)] = [t for t, p in sends if p == peer]     # matched in issue order, like NCCL
        sh['barrier'].wait()
        taken = {}
        for t, peer in recvs:
            i = taken.get(peer, 0)
            t.copy_(sh['mail'][(peer, self.rank)][i])
            taken[peer] = i + 1
        sh['barrier'].wait()


@pytest.fixture(autouse=True)
def exact_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic = old


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
@pytest.mark.parametrize('world', [2, 4])
def test_sharded_level_matches_unsharded(seeded_vgg, world, precision):
    from artstyletransfer_b200 import math_utils, neural_style_transfer as nst, ops, parallel
    H, W = 512, 64
    content, style = O.synthetic_images(H, W, seed=5)
    init = np.clip(content * 0.5 + np.random.default_rng(6).uniform(0, 1, size=content.shape) * 0.5, 0, 1).astype(np.float32)
    nst.PRECISION = precision
    try:
        net, cidx, sidx = math_utils.prepare_model('vgg19', dev())
        lb = nst.LossBuilder(cidx, sidx, nst.prepare_img(content, dev()), nst.prepare_img(style, dev()), net, *WEIGHTS)
        img = nst.prepare_img(init, dev()).requires_grad_(True)
        total, c, s, tv = lb.build(img)
        total.backward()
        ref = [v.item() for v in (total, c, s, tv)]
        ref_grad = img.grad.clone()

        shared = {'bufs': [None] * world, 'barrier': threading.Barrier(world)}
        grams = [g[0].detach().contiguous() for g in lb.target_style_representation]
        results = [None] * world
        errors = []

        def run(rank):
            try:
                torch.cuda.set_device(dev())
                grp = ThreadGroup(rank, world, shared)
                sh = parallel.ShardedLevel(grp, net, cidx, sidx, lb.target_content_representation, grams, WEIGHTS, H, W,
                                           ops._prec(precision))
                x = nst.prepare_img(init, dev()).requires_grad_(True)
                t, cc, ss, vv = sh.build(x)
                t.backward()
                results[rank] = ([v.item() for v in (t, cc, ss, vv)], x.grad.clone())
            except Exception as e:   # pragma: no cover
                errors.append(e)
                shared['barrier'].abort()

        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        [t.start() for t in threads]
        [t.join() for t in threads]
        assert not errors, root_causes(errors)
    finally:
        nst.PRECISION = None
    tol = 2e-5 if precision == 'fp32' else 2e-4
    for r in range(world):
        np.testing.assert_allclose(results[r][0], ref, rtol=tol)
        assert results[r][0] == results[0][0]                 # every rank sees bit-identical losses
    gsum = sum(results[r][1] for r in range(world))
    gerr = float(torch.linalg.norm(gsum - ref_grad) / torch.linalg.norm(ref_grad))
    assert gerr < (1e-4 if precision == 'fp32' else 2e-3)


@pytest.mark.timeout(240)
@pytest.mark.parametrize('world,H,W', [(2, 128, 96), (4, 128, 64), (4, 256, 48), (8, 256, 32)])
def test_halo_exchange_level_matches_unsharded(seeded_vgg, world, H, W):
    """Channels-last path, one halo row exchanged per convolution: nothing is approximated, so the sharded level
    must reproduce the unsharded one up to summation order (fp32-exact convolutions here)."""
    from artstyletransfer_b200 import math_utils, neural_style_transfer as nst, parallel
    from artstyletransfer_b200.sharded_path import ShardedPathLevel
    content, style = O.synthetic_images(H, W, seed=7)
    init = np.clip(content * 0.5 + np.random.default_rng(8).uniform(0, 1, size=content.shape) * 0.5, 0, 1).astype(np.float32)
    net, cidx, sidx = math_utils.prepare_model('vgg19', dev())
    c_img, s_img = nst.prepare_img(content, dev()), nst.prepare_img(style, dev())
    lb = nst.LossBuilder(cidx, sidx, c_img, s_img, net, *WEIGHTS)
    img = nst.prepare_img(init, dev()).requires_grad_(True)
    total, c, s, tv = lb.build(img)
    total.backward()
    ref = [v.item() for v in (total, c, s, tv)]
    ref_grad = img.grad.clone()
    plan = lb.path_plan(img)
    assert plan is not None

    shared = {'bufs': [None] * world, 'barrier': threading.Barrier(world, timeout=60), 'mail': {}}
    results = [None] * world
    errors = []

    def run(rank):
        try:
            torch.cuda.set_device(dev())
            grp = ThreadGroup(rank, world, shared)
            sh = ShardedPathLevel(grp, plan, c_img, s_img, cidx, sidx, WEIGHTS, H, W)
            out = []
            # the schedules are called directly: torch runs every autograd backward of a device on ONE engine thread,
            # so R emulated ranks whose backward passes wait for one another cannot go through .backward() here
            # (the autograd wrapper itself is exercised by the multi-GPU run, one process per rank)
            from artstyletransfer_b200.sharded_path import sharded_forward, sharded_backward
            for _ in range(2):                                  # persistent buffers: a second closure must agree
                x = nst.prepare_img(init, dev())
                with torch.no_grad():
                    out4, state = sharded_forward(sh, x)
                    grad = sharded_backward(state, None)
                out.append(([v.item() for v in out4], grad.clone()))
            assert out[0][0] == out[1][0] and torch.equal(out[0][1], out[1][1])
            results[rank] = out[0]
        except Exception as e:   # pragma: no cover
            import traceback
            errors.append(traceback.format_exc())
            shared['barrier'].abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, root_causes(errors)
    for r in range(world):
        np.testing.assert_allclose(results[r][0], ref, rtol=1e-4)
        assert results[r][0] == results[0][0]                 # every rank sees bit-identical losses
    gsum = sum(results[r][1] for r in range(world))
    gerr = float(torch.linalg.norm(gsum - ref_grad) / torch.linalg.norm(ref_grad))
    assert gerr < 5e-4, gerr
    # each rank's gradient lives in its band rows +- 1 (TV gradient on rank 0 aside)
    hb = H // world
    for r in range(1, world):
        gr = results[r][1]
        lo, hi = max(r * hb - 1, 0), min((r + 1) * hb + 1, H)
        assert float(gr[:, :, :lo].abs().max() if lo > 0 else 0.0) == 0.0
        assert float(gr[:, :, hi:].abs().max() if hi < H else 0.0) == 0.0


def run_lockstep(world, bands, n_levels, weights=WEIGHTS, content_idx=None, hw=None, halo_depth=None):
    """The unsharded closure (bicubic chain + every level + backward) and the same closure on `world` emulated ranks
    in lock-step.  Returns a dict: ref_total, ref_grad, totals (per rank), grads (per rank), plan, inputs."""
    from artstyletransfer_b200 import math_utils, neural_style_transfer as nst, ops
    from artstyletransfer_b200.parallel import PyramidBands
    from artstyletransfer_b200.sharded_path import PyramidFn, ShardedPathLevel, ShardedPyramid
    H, W = hw or ((256, 96) if n_levels == 2 else (256, 128))
    content, style = O.synthetic_images(H, W, seed=11)
    init = np.clip(content * 0.5 + np.random.default_rng(12).uniform(0, 1, size=content.shape) * 0.5, 0, 1).astype(np.float32)
    net, cidx, sidx = math_utils.prepare_model('vgg19', dev())
    if content_idx is not None:
        cidx = content_idx
    c_lv = [content[::1 << i, ::1 << i].copy() for i in range(n_levels)]
    s_lv = [style[::1 << i, ::1 << i].copy() for i in range(n_levels)]
    c_img = [nst.prepare_img(c, dev()) for c in c_lv]
    s_img = [nst.prepare_img(s_, dev()) for s_ in s_lv]
    lbs = [nst.LossBuilder(cidx, sidx, c, s_, net, *weights) for c, s_ in zip(c_img, s_img)]
    img = nst.prepare_img(init, dev()).requires_grad_(True)
    lv, total = img, None
    for i in range(n_levels):
        if i:
            lv = ops.bicubic_half(lv)
        t = lbs[i].build(lv)[0]
        total = t if total is None else 1.0 * total + t
    total.backward()
    ref_total, ref_grad = total.item(), img.grad.clone()
    plan = lbs[0].path_plan(img)
    sizes = [(H >> i, W >> i) for i in range(n_levels)]
    pb = PyramidBands(sizes, world, uniform=bands == 'uniform')
    depth = pb.halo_depth(halo_depth)              # halo rows per exchange (parallel.halo_schedule)
    assert halo_depth is None or depth == halo_depth

    shared = {'bufs': [None] * world, 'barrier': threading.Barrier(world, timeout=60), 'mail': {}}
    results = [None] * world
    errors = []

    class Ctx:
        needs_input_grad = (False, True)

    def run(rank):
        try:
            torch.cuda.set_device(dev())
            grp = ThreadGroup(rank, world, shared)
            levels = [ShardedPathLevel(grp, plan, c_img[i], s_img[i], cidx, sidx, weights, *sizes[i],
                                       band=(*pb.band(i, rank), *pb.neighbours(i, rank)), halo_depth=depth)
                      for i in range(n_levels)]
            pyr = ShardedPyramid(levels)
            out = []
            for _ in range(2):                      # persistent buffers: a second closure must agree bit for bit
                x = nst.prepare_img(init, dev())
                ctx = Ctx()
                with torch.no_grad():
                    t = PyramidFn.forward(ctx, pyr, x)
                    _, g = PyramidFn.backward(ctx, None)
                out.append((t.item(), g.clone()))
            assert out[0][0] == out[1][0] and torch.equal(out[0][1], out[1][1])
            results[rank] = out[0]
        except Exception:   # pragma: no cover
            import traceback
            errors.append(traceback.format_exc())
            shared['barrier'].abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, root_causes(errors)
    return {'ref_total': ref_total, 'ref_grad': ref_grad, 'totals': [r[0] for r in results],
            'grads': [r[1] for r in results], 'plan': pb, 'sizes': sizes, 'init': init, 'c_lv': c_lv, 's_lv': s_lv,
            'cidx': cidx, 'sidx': sidx}


def fp64_closure_grad(run, weights):
    """The oracle's torch closure in float64 on the device: the 'true' gradient both TF32 paths approximate."""
    onet, _, _ = O.make_vgg19(1234)
    onet = onet.to(dev()).double()
    targets = [O.torch_targets(onet, run['cidx'], run['sidx'], torch.from_numpy(O.prepare_img(c)).to(dev()).double(),
                               torch.from_numpy(O.prepare_img(s)).to(dev()).double())
               for c, s in zip(run['c_lv'], run['s_lv'])]
    _, _, g = O.torch_closure(onet, run['cidx'], run['sidx'], targets,
                              torch.from_numpy(O.prepare_img(run['init'])).to(dev()).double(), weights)
    return g


def row_error_profile(diff, ref):
    """Per image row r of the top level: ||diff[..., r, :]|| / rms over rows of ||ref[..., r, :]||."""
    d = torch.linalg.norm(diff.double().permute(2, 0, 1, 3).reshape(diff.shape[2], -1), dim=1)
    n = torch.linalg.norm(ref.double().permute(2, 0, 1, 3).reshape(ref.shape[2], -1), dim=1)
    return (d / n.pow(2).mean().sqrt()).cpu().numpy()


LOCKSTEP_CASES = [(2, 'uniform', 2), (4, 'uniform', 2), (2, 'pyramid', 3), (3, 'pyramid', 3), (4, 'pyramid', 3),
                  (8, 'pyramid', 3)]


@pytest.fixture()
def correctly_rounded_convs(monkeypatch):
    """Replace the four cuDNN call sites of the feature path by float64 convolutions rounded once to float32.

    Why: cuDNN's fp32 result for one output pixel depends (at the 1e-7 level) on the problem shape — a row band and
    the whole level take different tilings — and VGG19's max-pools / ReLUs are discontinuous: a 1-ulp difference
    flips the arg-max of a near-tied 2x2 window (~1e6 windows per closure, ~1 flip expected) and re-routes that
    window's gradient.  Measured on a B200 (tests/tools/shard_error_probe.py, profiles/r02_shard_error_probe.log): the
    sharded-vs-unsharded gradient differs by 1e-7 everywhere EXCEPT around one or two pixels per closure, far from
    any band edge and at the SAME pixel whatever the band layout — that is what made the 5e-4 bound of round 1 fail
    on one box and pass on another.  With correctly rounded convolutions both paths see bit-identical activations,
    no window can flip, and everything that IS this repo's code (bands, halo exchange, lock-step schedule, pooling
    and ReLU glue, partial Grams + all-reduce + finalize, bicubic chain) can be held to fp32 summation order."""
    from artstyletransfer_b200 import feature_path as fp, sharded_path as sp
    CL = torch.channels_last

    def fwd(x, w, b):
        y = torch.nn.functional.conv2d(x.double().contiguous(), w.double().contiguous(), b.double(), padding=1)
        return y.relu_().float().contiguous(memory_format=CL)

    def dgrad(g, x, w):
        gi = torch.nn.grad.conv2d_input(x.shape, w.double().contiguous(), g.double().contiguous(), stride=1, padding=1)
        return gi.float().contiguous(memory_format=CL)

    monkeypatch.setattr(fp, '_conv_relu_fwd', fwd)
    monkeypatch.setattr(fp, '_conv_bwd_data', dgrad)
    monkeypatch.setattr(sp, '_conv_relu_fwd_padded', fwd)
    monkeypatch.setattr(sp, '_conv_bwd_data_padded', dgrad)


def summarise(tag, run, world):
    gsum = sum(run['grads'])
    gerr = float(torch.linalg.norm(gsum - run['ref_grad']) / torch.linalg.norm(run['ref_grad']))
    prof = row_error_profile(gsum - run['ref_grad'], run['ref_grad'])
    lerr = max(abs(t - run['ref_total']) / abs(run['ref_total']) for t in run['totals'])
    print(f'{tag} plan {run["plan"].describe()}: loss rel err {lerr:.2e}, gradient sharded-vs-unsharded {gerr:.3e}, '
          f'row error median {np.median(prof):.2e} max {prof.max():.2e} at row {int(prof.argmax())}')
    for r in range(world):
        assert run['totals'][r] == run['totals'][0]                     # every rank sees bit-identical losses
    return lerr, gerr


@pytest.mark.timeout(300)
@pytest.mark.parametrize('depth', [2, 1])
@pytest.mark.parametrize('world,bands,n_levels', LOCKSTEP_CASES)
def test_lockstep_pyramid_matches_unsharded_closure(seeded_vgg, correctly_rounded_convs, world, bands, n_levels, depth):
    """The pyramid levels evaluated in lock-step with grouped halo exchanges == the unsharded closure (bicubic chain
    + every level + backward), summed over the emulated ranks.  'uniform': every level cut into `world` equal bands;
    'pyramid': the level-aware plan (parallel.PyramidBands) — unequal bands, ranks that own rows of two levels, ranks
    that own nothing of a level and are skipped by their neighbours' exchange.

    Convolutions are correctly rounded (see the fixture), so the only legitimate differences are this library's own
    summation orders: the Gram is a sum over positions whose split-K order differs between one 148-CTA launch and
    per-band launches + all-reduce (1e-7 of G), D = G - A amplifies that by |G|/|D| ~ 1e2 and is then rounded to TF32
    for the backward's operand — a bounded, band-edge-free 1e-5-level effect.  Asserted: loss 1e-5, gradient 1e-4."""
    run = run_lockstep(world, bands, n_levels, halo_depth=depth)
    pb = run['plan']
    if bands == 'pyramid' and world > 2:           # the plan really is heterogeneous at these sizes
        assert any(pb.band(li, r)[0] == pb.band(li, r)[1] for li in range(n_levels) for r in range(world))
    lerr, gerr = summarise(f'lockstep[{world}-{bands}-{n_levels}-depth{depth}]', run, world)
    assert lerr <= 1e-5, lerr
    assert gerr <= 1e-4, gerr


@pytest.mark.timeout(300)
@pytest.mark.parametrize('content_idx,depth', [(5, 2), (5, 1), (3, 2)])
@pytest.mark.parametrize('world,bands,n_levels', [(2, 'pyramid', 3), (4, 'pyramid', 3), (4, 'uniform', 2)])
def test_lockstep_pyramid_is_exact_without_gram_noise(seeded_vgg, correctly_rounded_convs, world, bands, n_levels,
                                                      content_idx, depth):
    """Style weight 0 and the content term moved to the deepest tap (relu5_1): the gradient then flows through every
    convolution, pool, halo exchange and the bicubic chain but through no TF32 Gram, so with correctly rounded
    convolutions the sharded closure must reproduce the unsharded one to fp32 summation order.  With two halo rows
    per exchange (depth 2) every second convolution runs on redundantly computed edge rows; content_idx 3 (relu4_1)
    puts the content tap on a backward step that runs WITHOUT an exchange, i.e. on the owned rows +- 1."""
    weights = (1e3, 0.0, 1e2)
    run = run_lockstep(world, bands, n_levels, weights=weights, content_idx=content_idx, halo_depth=depth)
    lerr, gerr = summarise(f'exact[{world}-{bands}-{n_levels}-tap{content_idx}-depth{depth}]', run, world)
    assert lerr <= 2e-6, lerr
    assert gerr < 2e-5, gerr


@pytest.mark.timeout(300)
@pytest.mark.parametrize('world,bands,n_levels', [(2, 'pyramid', 3), (8, 'pyramid', 3)])
def test_lockstep_pyramid_with_cudnn_convs_is_within_pool_flip_noise(seeded_vgg, world, bands, n_levels):
    """The same comparison with the REAL (cuDNN fp32) convolutions.  Band and whole-level shapes round differently, a
    handful of near-tied max-pool windows flip, and each flip re-routes one window's gradient: the difference is a
    few isolated patches (the float64 oracle shows the same patches against EITHER path), not an edge effect.  Bound:
    loss 1e-4 (BASELINE), gradient 5e-3 overall — a broken halo row would be a 1e-1 effect along a whole band edge —
    and both paths inside the TF32 gradient budget (2e-3, tests/test_gpu_closure.py) of the float64 gradient unless a
    flip separates them from it too."""
    run = run_lockstep(world, bands, n_levels)
    lerr, gerr = summarise(f'cudnn[{world}-{bands}-{n_levels}]', run, world)
    g64 = fp64_closure_grad(run, WEIGHTS)
    n64 = torch.linalg.norm(g64)
    e_un = float(torch.linalg.norm(run['ref_grad'].double() - g64) / n64)
    e_sh = float(torch.linalg.norm(sum(run['grads']).double() - g64) / n64)
    print(f'   vs float64 oracle: unsharded {e_un:.3e}, sharded {e_sh:.3e}')
    assert lerr <= 1e-4, lerr
    assert gerr < 5e-3 and e_sh < 5e-3 and e_un < 5e-3, (gerr, e_sh, e_un)


def test_band_plan_and_pack_layout():
    from artstyletransfer_b200.parallel import BandPlan, pack_layout
    rows = []
    for r in range(8):
        p = BandPlan(2048, r, 8)
        assert (p.r1 - p.r0) == 256 and p.lo % 16 == 0 and p.hi % 16 == 0
        assert p.r0 - p.lo in (0, 80) and p.hi - p.r1 in (0, 80)
        a, b, n = p.feat_rows(16)
        assert b - a == 16 and n == (p.hi - p.lo) // 16
        rows += list(range(p.r0, p.r1))
    assert rows == list(range(2048))
    offs, slot, n = pack_layout([64, 128, 256, 512, 512])
    assert offs == [0, 4096, 20480, 86016, 348160] and slot == 610304 and n == 610305
    assert not BandPlan.shardable(383, 2) and BandPlan.shardable(256, 8) and not BandPlan.shardable(256, 16)
