"""GPU parity: Gram matrix + fused MSE (K1) and its backward (K2), tcgen05 TF32 path and exact fp32 path,
called through the C ABI, vs the fp64 CPU oracle and the reference's goldens.

Tolerances (BASELINE.json north_star): per-layer Gram Frobenius relative error <= 1e-3.  Measured bounds are
much tighter and are asserted below: TF32 (operands rounded to nearest) <= 5e-4 (reached only by the K < 64
toy shapes, where rounding errors do not average out; ~1e-5 at VGG sizes), fp32 path <= 2e-6."""
import numpy as np
import pytest
import torch

from oracle import gatys_oracle as O

pytestmark = pytest.mark.gpu

FRO_TOL = {'tf32': 5e-4, 'fp32': 2e-6}
LOSS_TOL = {'tf32': 5e-3, 'fp32': 2e-5}     # per-layer MSE vs an INDEPENDENT target (no cancellation help)
GRAD_TOL = {'tf32': 2e-3, 'fp32': 1e-5}


def dev():
    return torch.device('cuda', 0)


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def features(c, hw, seed, sigma=0.25):
    """post-ReLU-like features relu(N(0,1))*sigma, ~50 % exact zeros."""
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn((c, hw), generator=g)) * sigma


SHAPES = [(64, 384), (64, 98304), (64, 4112), (128, 24576), (128, 6112), (256, 6144), (256, 1520), (512, 1536),
          (512, 368), (512, 1504), (512, 36), (64, 8), (256, 100)]


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
@pytest.mark.parametrize('c,hw', SHAPES)
def test_gram_matrix_vs_oracle(c, hw, precision):
    from artstyletransfer_b200 import math_utils
    f = features(c, hw, seed=c + hw)
    h = 4 if hw % 4 == 0 else 1
    x = f.view(1, c, h, hw // h).to(dev())
    g = math_utils.gram_matrix(x, precision=precision)
    ref = O.gram_matrix(f.view(1, c, h, hw // h).numpy())
    assert tuple(g.shape) == (1, c, c)
    assert rel(g.cpu().numpy(), ref) < FRO_TOL[precision]
    gn = math_utils.gram_matrix(x, should_normalize=False, precision=precision)
    assert rel(gn.cpu().numpy(), ref * (c * hw)) < FRO_TOL[precision]


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
def test_gram_matrix_golden(golden, precision):
    from artstyletransfer_b200 import math_utils
    gd = golden('small_ops.npz')
    x = torch.from_numpy(gd['gram_x_a']).to(dev())        # (1, 64, 8, 12)
    g = math_utils.gram_matrix(x, precision=precision)
    assert rel(g.cpu().numpy(), gd['gram_g_a']) < FRO_TOL[precision]
    x = torch.from_numpy(gd['gram_x_b']).to(dev())        # (1, 128, 6, 8)
    g = math_utils.gram_matrix(x, precision=precision)
    assert rel(g.cpu().numpy(), gd['gram_g_b']) < FRO_TOL[precision]


def test_gram_unsupported_shape_raises_for_tf32_but_fp32_works():
    from artstyletransfer_b200 import math_utils, ops, _lib
    x = torch.rand(1, 64, 3, 5, device=dev())     # HW = 15: not a multiple of 4 -> no TMA path
    out = torch.empty(64, 64, device=dev())
    assert _lib.load().ast_gram_tf32_supported(x.data_ptr(), 64, 15, 15) == 0
    with pytest.raises(RuntimeError, match='TF32 path'):      # the C ABI refuses ...
        _lib.call('ast_gram_mse_fwd', x.data_ptr(), 64, 15, 15, 1.0, None, out.data_ptr(), None,
                  ops.gram_workspace(64, 15, dev()).ptr, 1 << 20, _lib.AST_PREC_TF32, None)
    g = math_utils.gram_matrix(x, precision='tf32')            # ... the Python surface takes the exact CUDA kernel
    assert rel(g.cpu().numpy(), O.gram_matrix(x.cpu().numpy())) < 2e-6
    g = math_utils.gram_matrix(x, precision='fp32')
    assert rel(g.cpu().numpy(), O.gram_matrix(x.cpu().numpy())) < 2e-6
    with pytest.raises(RuntimeError, match='multiple of 64'):
        math_utils.gram_matrix(torch.rand(1, 16, 4, 4, device=dev()), precision='fp32')


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
@pytest.mark.parametrize('c,hw', [(64, 98304), (64, 4112), (128, 24576), (256, 6144), (256, 1520), (512, 1536),
                                  (512, 368), (64, 8)])
def test_style_loss_fwd_bwd_vs_oracle(c, hw, precision):
    from artstyletransfer_b200.neural_nets import StyleLoss
    f = features(c, hw, seed=1 + c + hw)
    a_src = features(c, hw, seed=2 + c + hw, sigma=0.27)
    a = O.gram_matrix(a_src.view(1, c, 1, hw).numpy())[0]
    x = f.view(1, c, 4, hw // 4).to(dev()).requires_grad_(True)
    mod = StyleLoss(torch.from_numpy(a.astype(np.float32)).to(dev()), precision=precision)
    loss = mod(x)
    (loss * 3.0).backward()
    ref_loss, _, d = O.style_layer_mse(f.numpy(), a.astype(np.float32))
    assert abs(loss.item() - ref_loss) / ref_loss < LOSS_TOL[precision]
    ref_grad = 3.0 * O.style_layer_grad(f.numpy(), d)
    assert rel(x.grad.cpu().numpy().reshape(c, hw), ref_grad) < GRAD_TOL[precision]
    # deterministic: a second evaluation is bit-identical
    loss2 = mod(x.detach())
    assert loss2.item() == loss.item()


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
def test_style_loss_golden(golden, precision):
    from artstyletransfer_b200.neural_nets import StyleLoss
    from artstyletransfer_b200 import math_utils
    gd = golden('small_ops.npz')
    if precision == 'tf32':
        pytest.skip('golden case has C=32 (not a VGG width): exact path only')
    # C = 32 is not supported by either kernel family -> the reference-shaped golden uses the closed form via
    # gram_matrix on a zero-padded 64-channel copy (zero rows change neither G[:32,:32] nor the MSE numerator)
    x = gd['style_x'][0]
    pad = np.zeros((1, 64) + x.shape[1:], np.float32); pad[0, :32] = x
    g = math_utils.gram_matrix(torch.from_numpy(pad).to(dev()), should_normalize=False, precision=precision)
    g32 = g[0, :32, :32].cpu().numpy() / (32 * x.shape[1] * x.shape[2])
    loss = np.mean((gd['style_a'][0] - g32) ** 2)
    assert abs(loss - float(gd['style_loss'])) / float(gd['style_loss']) < 1e-4


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
def test_gram_matrix_autograd_general(precision):
    """gram_matrix is differentiable for any upstream gradient, like the reference's bmm graph."""
    from artstyletransfer_b200 import math_utils
    f = features(64, 512, seed=5).view(1, 64, 16, 32)
    w = torch.randn(1, 64, 64, generator=torch.Generator().manual_seed(6))
    x = f.to(dev()).requires_grad_(True)
    (math_utils.gram_matrix(x, precision=precision) * w.to(dev())).sum().backward()
    xr = f.double().requires_grad_(True)
    fr = xr.view(1, 64, -1)
    ((fr.bmm(fr.transpose(1, 2)) / (64 * 512)) * w.double()).sum().backward()
    assert rel(x.grad.cpu().numpy(), xr.grad.numpy()) < GRAD_TOL[precision]


def test_gram_bwd_accumulate_flag():
    from artstyletransfer_b200 import ops
    for c, hw, prec in [(64, 640, 0), (128, 1000, 0), (64, 640, 1)]:
        f = features(c, hw, seed=9).to(dev())
        d = torch.randn(c, c, generator=torch.Generator().manual_seed(10)).to(dev())
        d = (d + d.t()).contiguous()
        fresh = torch.empty_like(f)
        ops.gram_bwd(d, f, c, hw, 0.5, None, fresh, False, prec)
        base = torch.randn(c, hw, generator=torch.Generator().manual_seed(11)).to(dev())
        acc = base.clone()
        gs = torch.tensor(2.0, device=dev())
        ops.gram_bwd(d, f, c, hw, 0.25, gs, acc, True, prec)        # 0.25 * 2.0 == 0.5
        assert torch.allclose(acc, base + fresh, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize('c,hw', [(64, 6291456), (128, 1572864), (256, 393216), (512, 98304)])
def test_gram_full_size_properties(c, hw):
    """BASELINE L=3 top-level shapes: properties that need no CPU Gram — symmetry, trace equals
    the fp64 sum of squares, 4x under 2x scaling, and agreement of the two kernel families."""
    from artstyletransfer_b200 import math_utils
    g = torch.Generator(device='cuda').manual_seed(c)
    x = torch.relu(torch.randn((1, c, 1, hw), generator=g, device=dev())) * 0.25
    gt = math_utils.gram_matrix(x, precision='tf32')[0]
    assert torch.equal(gt, gt.t()) or rel(gt.cpu().numpy(), gt.t().cpu().numpy()) < 1e-6
    trace_ref = (x.double() ** 2).sum().item() / (c * hw)
    assert abs(gt.double().trace().item() - trace_ref) / trace_ref < 1e-4
    g2 = math_utils.gram_matrix(x * 2.0, precision='tf32')[0]
    assert rel(g2.cpu().numpy(), 4.0 * gt.cpu().numpy()) < 1e-6      # power-of-two scaling is exact in TF32
    gf = math_utils.gram_matrix(x, precision='fp32')[0]
    assert rel(gt.cpu().numpy(), gf.cpu().numpy()) < 1e-4
