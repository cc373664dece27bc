"""GPU parity of the channels-last (HW, C) Gram kernels: against the fp64 oracle on seeded inputs, against the
NCHW kernel family (bit-for-bit: same accumulation order), and at BASELINE sizes through properties."""
import numpy as np
import pytest
import torch

from oracle import gatys_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda', 0)


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) /
                 max(np.linalg.norm(np.asarray(b, np.float64)), 1e-300))


SHAPES = [(64, 8), (64, 100), (64, 4096 + 37), (128, 31), (128, 5000), (256, 129), (256, 3000), (512, 77),
          (512, 1536), (512, 2500), (64, 24576), (128, 6144)]


def _feat(c, hw, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return (torch.relu(torch.randn((hw, c), generator=g, device=dev())) * 0.25).contiguous()


@pytest.mark.parametrize('c,hw', SHAPES)
def test_gram_fwd_nhwc_vs_oracle(c, hw):
    from artstyletransfer_b200 import ops
    f = _feat(c, hw, c * 7 + hw)
    a = torch.rand((c, c), device=dev()) * 1e-3
    a = (a + a.t()) / 2
    d = torch.empty((c, c), device=dev())
    loss = torch.empty((), device=dev())
    ws = ops.gram_workspace(c, hw, dev())
    ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), a, d, loss, ws)
    fn = f.cpu().numpy().astype(np.float64).T          # (C, HW)
    g_ref = fn @ fn.T / (c * hw)
    d_ref = g_ref - a.cpu().numpy().astype(np.float64)
    g_gpu = d.cpu().numpy().astype(np.float64) + a.cpu().numpy().astype(np.float64)
    assert rel(g_gpu, g_ref) < 5e-4                     # budget 1e-3 (north_star); TF32 operands
    assert abs(float(loss) - float((d_ref ** 2).mean())) <= 2e-2 * float((d_ref ** 2).mean()) + 1e-12
    assert torch.equal(d, d.t()) or rel(d.cpu().numpy(), d.t().cpu().numpy()) < 1e-6
    # plain Gram (no target), rerun is bit-identical (fixed-order split-K)
    g1 = torch.empty((c, c), device=dev()); g2 = torch.empty((c, c), device=dev())
    ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), None, g1, None, ws)
    ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), None, g2, None, ws)
    assert torch.equal(g1, g2)


@pytest.mark.parametrize('c,hw', [(64, 4096), (128, 5000), (256, 3000), (512, 1536)])
def test_gram_fwd_nhwc_matches_nchw_kernel(c, hw):
    """Both layouts feed the tensor core the same rounded operands in the same order per CTA -> equal results."""
    from artstyletransfer_b200 import ops
    f = _feat(c, hw, 3)
    fn = f.t().contiguous()
    ws = ops.gram_workspace(c, hw, dev())
    g1 = torch.empty((c, c), device=dev()); g2 = torch.empty((c, c), device=dev())
    ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), None, g1, None, ws)
    ops.gram_mse_fwd(fn, c, hw, 1.0 / (c * hw), None, g2, None, ws, 0)
    assert rel(g1.cpu().numpy(), g2.cpu().numpy()) < 1e-6


@pytest.mark.parametrize('hw', [24576, 98304])
def test_gram_fwd_c512_both_layouts_at_vgg_sizes_repeated(hw):
    """relu4_1 of the 1024x1536 / 2048x3072 levels (C = 512), what every LossBuilder.__init__ sends through
    math_utils.gram_matrix: NCHW and (HW, C) operands, 25 back-to-back launches each with the L2 flushed in between
    (the round-1 tuning sweep once died with a launch failure at the NCHW shape under a non-default pipeline shape)."""
    from artstyletransfer_b200 import ops
    c = 512
    f = _feat(c, hw, 11)
    fn = f.t().contiguous()
    ws = ops.gram_workspace(c, hw, dev())
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev())
    g_nhwc = torch.empty((c, c), device=dev()); g_nchw = torch.empty((c, c), device=dev())
    first = None
    for it in range(25):
        flush.zero_()
        ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), None, g_nhwc, None, ws)
        ops.gram_mse_fwd(fn, c, hw, 1.0 / (c * hw), None, g_nchw, None, ws, 0)
        if it % 8 == 0:
            torch.cuda.synchronize()
            if first is None:
                first = (g_nhwc.clone(), g_nchw.clone())
            assert torch.equal(g_nhwc, first[0]) and torch.equal(g_nchw, first[1])
    torch.cuda.synchronize()
    assert rel(g_nhwc.cpu().numpy(), g_nchw.cpu().numpy()) < 1e-6
    trace_ref = (f.double() ** 2).sum().item() / (c * hw)
    assert abs(g_nhwc.double().trace().item() - trace_ref) / trace_ref < 1e-4


@pytest.mark.parametrize('c,hw', SHAPES)
@pytest.mark.parametrize('accumulate', [False, True])
def test_gram_bwd_nhwc_vs_oracle(c, hw, accumulate):
    from artstyletransfer_b200 import ops
    f = _feat(c, hw, c + hw)
    d = torch.randn((c, c), device=dev()) * 1e-2
    d = ((d + d.t()) / 2).contiguous()
    gs = torch.tensor(0.5, device=dev())
    base = torch.randn((hw, c), device=dev()) * 1e-3
    df = base.clone() if accumulate else torch.full((hw, c), float('nan'), device=dev())
    ops.gram_bwd_nhwc(d, f, c, hw, 2.0, gs, df, accumulate)
    ref = 1.0 * (f.double() @ d.double())                 # dF[p, c] = s * sum_k F[p, k] D[k, c], s = 2 * 0.5
    if accumulate:
        ref = ref + base.double()
    assert not torch.isnan(df).any()
    assert rel(df.cpu().numpy(), ref.cpu().numpy()) < 2e-3


@pytest.mark.parametrize('c,hw', [(64, 6291456), (128, 1572864), (256, 393216), (512, 98304)])
def test_gram_nhwc_full_size_properties(c, hw):
    """BASELINE L=3 top-level shapes: trace = fp64 sum of squares, symmetry, exact 4x under 2x scaling; backward is
    linear in D and matches a sampled fp64 product."""
    from artstyletransfer_b200 import ops
    f = _feat(c, hw, c)
    ws = ops.gram_workspace(c, hw, dev())
    g1 = torch.empty((c, c), device=dev()); g2 = torch.empty((c, c), device=dev())
    ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), None, g1, None, ws)
    trace_ref = (f.double() ** 2).sum().item() / (c * hw)
    assert abs(g1.double().trace().item() - trace_ref) / trace_ref < 1e-4
    assert rel(g1.cpu().numpy(), g1.t().cpu().numpy()) < 1e-6
    ops.gram_mse_fwd_nhwc(f * 2.0, c, hw, 1.0 / (c * hw), None, g2, None, ws)
    assert rel(g2.cpu().numpy(), 4.0 * g1.cpu().numpy()) < 1e-6
    d = ((g1 + g1.t()) / 2).contiguous()
    df = torch.empty_like(f)
    ops.gram_bwd_nhwc(d, f, c, hw, 1.0, None, df, False)
    idx = torch.randint(0, hw, (512,), device=dev())
    ref = f[idx].double() @ d.double()
    assert rel(df[idx].cpu().numpy(), ref.cpu().numpy()) < 2e-3
    df2 = df.clone()
    ops.gram_bwd_nhwc(d, f, c, hw, 1.0, None, df2, True)   # accumulate: exactly doubles
    assert rel(df2[idx].cpu().numpy(), 2 * ref.cpu().numpy()) < 2e-3


@pytest.mark.parametrize('c,hw', [(64, 4096 + 37), (128, 5000), (256, 3000), (512, 2500)])
def test_prerounded_d_gives_bit_identical_backward(c, hw):
    """round_out stores D already rounded to TF32 and d_prerounded lets the backward's converters skip it: the
    tensor core must see exactly the same operand bits either way; the loss is computed from the exact D."""
    from artstyletransfer_b200 import ops
    f = _feat(c, hw, 17)
    a = torch.rand((c, c), device=dev()) * 1e-3
    a = (a + a.t()) / 2
    ws = ops.gram_workspace(c, hw, dev())
    d0 = torch.empty((c, c), device=dev()); d1 = torch.empty((c, c), device=dev())
    l0 = torch.empty((), device=dev()); l1 = torch.empty((), device=dev())
    ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), a, d0, l0, ws)
    ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), a, d1, l1, ws, round_out=True)
    assert l0.item() == l1.item()
    assert rel(d1.cpu().numpy(), d0.cpu().numpy()) < 1e-3 and not torch.equal(d0, d1)
    assert torch.equal(d1.view(torch.int32) & 0x1fff, torch.zeros_like(d1, dtype=torch.int32))
    g0 = torch.empty_like(f); g1 = torch.empty_like(f)
    ops.gram_bwd_nhwc(d0, f, c, hw, 3.0, None, g0, False)
    ops.gram_bwd_nhwc(d1, f, c, hw, 3.0, None, g1, False, d_prerounded=True)
    assert torch.equal(g0, g1)
    # ast_gram_finalize has the same switch (sharded path)
    raw = torch.empty((c, c), device=dev()); d2 = torch.empty((c, c), device=dev()); l2 = torch.empty((), device=dev())
    ops.gram_mse_fwd_nhwc(f, c, hw, 1.0, None, raw, None, ws)
    ops.gram_finalize(raw, c, 1.0 / (c * hw), a, d2, l2, ops.reduce_workspace(dev()), round_out=True)
    assert torch.equal(d2.view(torch.int32) & 0x1fff, torch.zeros_like(d2, dtype=torch.int32))
    assert rel(d2.cpu().numpy(), d0.cpu().numpy()) < 1e-3


@pytest.mark.parametrize('c,hw', [(64, 100), (64, 40), (64, 4096 + 37), (64, 57000), (128, 5000), (128, 40001), (256, 3000),
                                  (256, 39000), (512, 1536)])
@pytest.mark.parametrize('accumulate', [False, True])
def test_gram_bwd_nhwc_fused_relu_backward(c, hw, accumulate):
    """relu_mask: the kernel writes the gradient w.r.t. the ReLU's INPUT — bit-identical to the unfused sequence
    (gram_bwd accumulate, then relu_bwd in place) because both add the same fp32 values in the same order."""
    from artstyletransfer_b200 import ops
    f = _feat(c, hw, 5 * c + hw)
    d = torch.randn((c, c), device=dev()) * 1e-2
    d = ((d + d.t()) / 2).contiguous()
    base = torch.randn((hw, c), device=dev()) * 1e-3
    g_fused = base.clone() if accumulate else torch.full((hw, c), float('nan'), device=dev())
    ops.gram_bwd_nhwc(d, f, c, hw, 1.5, None, g_fused, accumulate, relu_mask=True)
    g_ref = base.clone() if accumulate else torch.empty((hw, c), device=dev())
    ops.gram_bwd_nhwc(d, f, c, hw, 1.5, None, g_ref, accumulate)
    ops.relu_bwd_(g_ref, f)
    assert not torch.isnan(g_fused).any()
    assert torch.equal(g_fused, g_ref)
    assert float(g_fused[f <= 0].abs().max()) == 0.0


def test_mse_bwd_fused_relu_backward():
    from artstyletransfer_b200 import ops
    n = 512 * 33 * 7 + 3
    x = torch.relu(torch.randn(n, device=dev())); t = torch.randn(n, device=dev())
    base = torch.randn(n, device=dev())
    for acc in (False, True):
        a = base.clone(); b = base.clone()
        ops.mse_bwd(x, t, 0.25, None, a, acc, True)
        ops.mse_bwd(x, t, 0.25, None, b, acc, False)
        b = torch.where(x > 0, b, torch.zeros_like(b))
        assert torch.equal(a, b)


def test_gram_finalize_batch_matches_single_finalize():
    """ast_gram_finalize_batch (all Grams of all pyramid levels in one launch, sharded path) == ast_gram_finalize item
    by item: same D bits (incl. TF32 rounding of the stored D), losses equal to fp32 round-off, reusable workspace."""
    from artstyletransfer_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(21)
    items, singles = [], []
    for n, c in enumerate([64, 128, 256, 512, 512, 64, 128]):
        raw = torch.rand((c, c), generator=g, device=dev()) * 1e4
        a = torch.rand((c, c), generator=g, device=dev()) * 1e-2 if n != 5 else None
        scale = 1.0 / (c * 1000.0 * (n + 1))
        rnd = n % 2 == 0
        d_b = torch.empty((c, c), device=dev()); l_b = torch.empty((), device=dev()) if n != 6 else None
        d_s = torch.empty((c, c), device=dev()); l_s = torch.empty((), device=dev())
        items.append((raw, c, scale, a, d_b, l_b, rnd))
        ops.gram_finalize(raw, c, scale, a, d_s, l_s, ops.reduce_workspace(dev()), round_out=rnd)
        singles.append((d_s, l_s))
    ws = ops.finalize_batch_workspace(len(items), dev())
    for _ in range(2):                                   # second launch: the workspace must have been left reusable
        ops.gram_finalize_batch(items, ws)
        for (raw, c, scale, a, d_b, l_b, rnd), (d_s, l_s) in zip(items, singles):
            assert torch.equal(d_b, d_s)
            if l_b is not None:
                assert abs(float(l_b) - float(l_s)) <= 1e-6 * abs(float(l_s))


@pytest.mark.parametrize('hw', [77, 1536, 2500, 24576])
@pytest.mark.parametrize('mode', ['store', 'accumulate', 'relu'])
def test_gram_bwd_bf16_c512_vs_oracle(hw, mode):
    """AST_PREC_BF16: the C = 512 backward with bfloat16 operands (D from the finalize kernel with round_out = 2, the
    feature tile rounded in shared memory), fp32 accumulation, against the float64 product of the EXACT operands
    (tolerance: bf16 has 8 mantissa bits, errors average over K = 512) and against the float64 product of the
    bf16-rounded operands (what the tensor core must compute: fp32-accumulation tight)."""
    from artstyletransfer_b200 import ops
    c = 512
    f = _feat(c, hw, 41 + hw)
    a = torch.rand((c, c), device=dev()) * 1e-3
    a = ((a + a.t()) / 2).contiguous()
    ws = ops.gram_workspace(c, hw, dev())
    d32 = torch.empty((c, c), device=dev()); l32 = torch.empty((), device=dev())
    dbf = torch.empty((c, c), dtype=torch.bfloat16, device=dev()); lbf = torch.empty((), device=dev())
    ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), a, d32, l32, ws)
    ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), a, dbf, lbf, ws, round_out=2)
    assert l32.item() == lbf.item()                                   # the loss never sees the rounding
    assert torch.equal(dbf, d32.to(torch.bfloat16))                   # round to nearest even, like torch
    base = torch.randn((hw, c), device=dev()) * 1e-3
    acc = mode != 'store'
    g = base.clone() if acc else torch.full((hw, c), float('nan'), device=dev())
    ops.gram_bwd_nhwc_auto(dbf, f, c, hw, 2.0, torch.tensor(0.5, device=dev()), g, acc, relu_mask=mode == 'relu')
    exact = f.double() @ d32.double()
    rounded = f.to(torch.bfloat16).double() @ dbf.double()
    if acc:
        exact, rounded = exact + base.double(), rounded + base.double()
    if mode == 'relu':
        exact, rounded = exact * (f > 0), rounded * (f > 0)
    assert not torch.isnan(g).any()
    assert rel(g.cpu().numpy(), rounded.cpu().numpy()) < 2e-5
    assert rel(g.cpu().numpy(), exact.cpu().numpy()) < 6e-3
