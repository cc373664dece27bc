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
        sh['bufs'][self.rank] = t
        sh['barrier'].wait()
        total = sh['bufs'][0].clone()
        for r in range(1, self.world):
            total = total + sh['bufs'][r]
        sh['barrier'].wait()
        t.copy_(total)
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


def run_lockstep(world, bands, n_levels, weights=WEIGHTS, content_idx=None, hw=None):
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
                                       band=(*pb.band(i, rank), *pb.neighbours(i, rank))) for i in range(n_levels)]
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
# TF32 Gram operands: the same image-gradient budget as the unsharded closure (tests/test_gpu_closure.py)
GRAD_BUDGET_TF32 = 2e-3


@pytest.mark.timeout(300)
@pytest.mark.parametrize('world,bands,n_levels', LOCKSTEP_CASES)
def test_lockstep_pyramid_matches_unsharded_closure(seeded_vgg, world, bands, n_levels):
    """The pyramid levels evaluated in lock-step with grouped halo exchanges == the unsharded closure (bicubic chain
    + every level + backward), summed over the emulated ranks.  'uniform': every level cut into `world` equal bands;
    'pyramid': the level-aware plan (parallel.PyramidBands) — unequal bands, ranks that own rows of two levels, ranks
    that own nothing of a level and are skipped by their neighbours' exchange.

    Tolerance.  Nothing in the band scheme is approximate (test_lockstep_pyramid_is_exact_without_gram_noise below
    shows 1e-5 agreement once the TF32 Gram is out of the loss), but the two paths round DIFFERENT values: the Gram
    is a sum over positions, its split-K order differs between one 148-CTA launch and per-band launches + all-reduce,
    D = G - A is ~1e-2 of G (cancellation), and D is then rounded to TF32 (2^-11) as the backward's operand — a
    1-ulp fp32 difference in G flips TF32 roundings of D.  So sharded-vs-unsharded is bounded by the TF32 noise of
    EACH against the true (float64) gradient, not by a number tuned on one box:
      * both paths are within the TF32 gradient budget (2e-3) of the float64 oracle gradient;
      * the sharded path is no further from it than the unsharded one (x1.5 + 1e-4 slack);
      * the difference is not concentrated at band edges (a halo bug would be): rows within 2 of a band edge of the
        top level carry no more error than 3x the typical row."""
    from artstyletransfer_b200.parallel import PyramidBands  # noqa: F401
    run = run_lockstep(world, bands, n_levels)
    pb, ref_total, ref_grad = run['plan'], run['ref_total'], run['ref_grad']
    if bands == 'pyramid' and world > 2:           # the plan really is heterogeneous at these sizes
        assert any(pb.band(li, r)[0] == pb.band(li, r)[1] for li in range(n_levels) for r in range(world))
    for r in range(world):
        assert abs(run['totals'][r] - ref_total) <= 1e-4 * abs(ref_total)
        assert run['totals'][r] == run['totals'][0]
    gsum = sum(run['grads'])
    g64 = fp64_closure_grad(run, WEIGHTS)
    n64 = torch.linalg.norm(g64)
    e_un = float(torch.linalg.norm(ref_grad.double() - g64) / n64)
    e_sh = float(torch.linalg.norm(gsum.double() - g64) / n64)
    gerr = float(torch.linalg.norm(gsum - ref_grad) / torch.linalg.norm(ref_grad))
    prof = row_error_profile(gsum - ref_grad, ref_grad)
    edges = sorted({e for e in pb.bounds[0][1:-1] if 0 < e < run['sizes'][0][0]})
    near = sorted({r for e in edges for r in range(e - 2, e + 2)})
    typical = float(np.median(prof))
    worst_edge = float(prof[near].max()) if near else 0.0
    print(f'lockstep[{world}-{bands}-{n_levels}] sharded-vs-unsharded={gerr:.3e} unsharded-vs-fp64={e_un:.3e} '
          f'sharded-vs-fp64={e_sh:.3e} row error: median={typical:.3e} max={prof.max():.3e} at band edges {edges}: '
          f'{worst_edge:.3e}')
    assert e_un < GRAD_BUDGET_TF32 and e_sh < GRAD_BUDGET_TF32, (e_un, e_sh)
    assert e_sh <= 1.5 * e_un + 1e-4, (e_sh, e_un)
    assert gerr <= e_un + e_sh + 1e-6, (gerr, e_un, e_sh)           # triangle inequality: a sanity check of the probe
    assert worst_edge <= 3.0 * typical + 1e-6, (worst_edge, typical, edges)


@pytest.mark.timeout(300)
@pytest.mark.parametrize('world,bands,n_levels', [(2, 'pyramid', 3), (4, 'pyramid', 3), (4, 'uniform', 2)])
def test_lockstep_pyramid_is_exact_without_gram_noise(seeded_vgg, world, bands, n_levels):
    """Style weight 0 and the content term moved to the deepest tap (relu5_1): the gradient then flows through every
    convolution, pool, halo exchange and the bicubic chain but through no TF32 Gram, so the sharded closure must
    reproduce the unsharded one to fp32 summation order (convolutions are fp32-exact in this module)."""
    weights = (1e3, 0.0, 1e2)
    run = run_lockstep(world, bands, n_levels, weights=weights, content_idx=5)
    for r in range(world):
        assert abs(run['totals'][r] - run['ref_total']) <= 2e-6 * abs(run['ref_total'])
    gsum = sum(run['grads'])
    gerr = float(torch.linalg.norm(gsum - run['ref_grad']) / torch.linalg.norm(run['ref_grad']))
    print(f'exact[{world}-{bands}-{n_levels}] sharded-vs-unsharded={gerr:.3e}')
    assert gerr < 5e-5, gerr


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
