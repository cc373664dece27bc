"""GPU parity: content MSE, total variation, bicubic pyramid (+adjoint), general resize, noise init — CUDA
kernels called through the C ABI vs the CPU oracle and the committed goldens."""
import numpy as np
import pytest
import torch

from oracle import gatys_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda', 0)


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize('shape', [(512, 32, 48), (512, 7, 9), (3, 5, 7), (1, 1, 1)])
def test_content_mse_fwd_bwd(shape):
    from artstyletransfer_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.relu(torch.randn(shape, generator=g)) * 3
    t = torch.relu(torch.randn(shape, generator=g)) * 3
    xd = x.to(dev()).requires_grad_(True)
    loss = ops.ContentLossFn.apply(xd, t.to(dev()), ops.reduce_workspace(dev()))
    (loss * 7.0).backward()
    ref_loss, ref_grad = O.content_mse(t.numpy(), x.numpy())
    assert abs(loss.item() - ref_loss) <= 2e-6 * max(ref_loss, 1e-30)      # fp32 data, fp64 reduction
    assert rel(xd.grad.cpu().numpy(), 7.0 * ref_grad) < 1e-6


def test_content_mse_unaligned_and_reuse():
    from artstyletransfer_b200 import ops
    g = torch.Generator().manual_seed(1)
    base_x = torch.randn(4099, generator=g).to(dev()); base_t = torch.randn(4099, generator=g).to(dev())
    x, t = base_x[1:], base_t[1:]               # 4-byte aligned only -> scalar path
    ws = ops.reduce_workspace(dev())
    for _ in range(3):                          # workspace reusable without re-zeroing
        loss = ops.ContentLossFn.apply(x, t, ws)
        ref, _ = O.content_mse(t.cpu().numpy(), x.cpu().numpy())
        assert abs(loss.item() - ref) <= 2e-6 * ref


@pytest.mark.parametrize('shape', [(1, 3, 64, 96), (1, 3, 20, 30), (1, 3, 33, 47), (2, 3, 8, 12)])
def test_total_variation_fwd_bwd(shape):
    from artstyletransfer_b200 import math_utils
    g = torch.Generator().manual_seed(2)
    y = torch.randn(shape, generator=g) * 50
    yd = y.to(dev()).requires_grad_(True)
    tv = math_utils.total_variation(yd)
    (tv * 3.0).backward()
    ref, ref_grad = O.total_variation(y.numpy())
    assert abs(tv.item() - ref) <= 2e-6 * ref
    assert rel(yd.grad.cpu().numpy(), 3.0 * ref_grad) < 1e-6


@pytest.mark.parametrize('shape,rows', [((3, 64, 48), (16, 48)), ((3, 64, 48), (0, 64)), ((3, 33, 7), (1, 2)),
                                        ((3, 33, 7), (30, 33)), ((1, 8, 12), (0, 3))])
def test_tv_bwd_rows_is_the_full_gradient_on_those_rows(shape, rows):
    """ast_tv_bwd_rows (a sharded rank adds the TV gradient of ITS rows before the gather): bit-identical to the rows
    of ast_tv_bwd's result, every other row untouched; vectorised (W % 4 == 0) and scalar paths, accumulate or not."""
    from artstyletransfer_b200 import ops
    g = torch.Generator().manual_seed(3)
    y = torch.randn((1, *shape), generator=g).to(dev())
    sums2 = torch.empty(2, device=dev())
    tv = torch.empty(1, device=dev())
    ops.tv_fwd(y, sums2, tv, ops.reduce_workspace(dev()))
    gs = torch.tensor([0.7], device=dev())
    full = torch.empty_like(y)
    ops.tv_bwd(y, sums2, 100.0, gs, full, False)
    r0, r1 = rows
    for accumulate in (False, True):
        base = torch.randn(y.shape, generator=g).to(dev())
        out = base.clone()
        ops.tv_bwd(y, sums2, 100.0, gs, out, accumulate, rows=rows)
        want = base.clone()
        want[:, :, r0:r1] = (base[:, :, r0:r1] + full[:, :, r0:r1]) if accumulate else full[:, :, r0:r1]
        assert torch.equal(out, want)
    with pytest.raises(Exception, match='rows'):
        ops.tv_bwd(y, sums2, 1.0, None, full, False, rows=(r1, r1))


def test_total_variation_golden(golden):
    from artstyletransfer_b200 import math_utils
    gd = golden('small_ops.npz')
    yd = torch.from_numpy(gd['tv_y']).to(dev()).requires_grad_(True)
    tv = math_utils.total_variation(yd)
    tv.backward()
    assert abs(tv.item() - float(gd['tv'])) <= 1e-5 * float(gd['tv'])
    assert rel(yd.grad.cpu().numpy(), gd['tv_grad']) < 1e-5


@pytest.mark.parametrize('shape', [(1, 3, 64, 96), (1, 3, 256, 384), (1, 3, 34, 262), (1, 3, 50, 77), (1, 3, 6, 10),
                                   (1, 3, 2, 2), (1, 3, 31, 33)])
def test_bicubic_half_and_adjoint(shape):
    from artstyletransfer_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(shape, generator=g) * 60
    xd = x.to(dev()).requires_grad_(True)
    y = ops.bicubic_half(xd)
    ref = O.bicubic_down2x(x.numpy())
    assert tuple(y.shape) == ref.shape
    # |x| ~ 60..250; the general-ratio path forms its weights in fp32 like torch does: a few 1e-7 relative
    assert np.abs(y.detach().cpu().numpy() - ref).max() < 2e-4
    u = torch.randn(y.shape, generator=g)
    (y * u.to(dev())).sum().backward()
    ref_adj = O.bicubic_resize_chw_adjoint(u.numpy(), shape[2], shape[3])
    assert np.abs(xd.grad.cpu().numpy() - ref_adj).max() < 2e-6
    # adjoint identity <A x, u> == <x, A^T u> evaluated in fp64 from the kernel outputs
    lhs = (y.detach().double().cpu() * u.double()).sum().item()
    rhs = (x.double() * xd.grad.double().cpu()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)


def test_bicubic_pyramid_golden(golden):
    from artstyletransfer_b200 import ops
    gd = golden('small_ops.npz')
    for name in ('even', 'odd'):
        p0 = torch.from_numpy(gd[f'pyr_{name}_p0']).to(dev()).requires_grad_(True)
        p1 = ops.bicubic_half(p0)
        p2 = ops.bicubic_half(p1)
        assert np.abs(p1.detach().cpu().numpy() - gd[f'pyr_{name}_p1']).max() < 1e-4
        assert np.abs(p2.detach().cpu().numpy() - gd[f'pyr_{name}_p2']).max() < 1e-4
        ((p1 * torch.from_numpy(gd[f'pyr_{name}_u1']).to(dev())).sum() +
         (p2 * torch.from_numpy(gd[f'pyr_{name}_u2']).to(dev())).sum()).backward()
        assert np.abs(p0.grad.cpu().numpy() - gd[f'pyr_{name}_grad']).max() < 1e-5


def test_bicubic_adj_accumulate():
    from artstyletransfer_b200 import ops
    g = torch.Generator().manual_seed(4)
    for shape in [(3, 32, 48), (3, 25, 38)]:
        gy = torch.randn(shape, generator=g).to(dev())
        in_h, in_w = (64, 96) if shape[1] == 32 else (50, 77)
        base = torch.randn((3, in_h, in_w), generator=g).to(dev())
        fresh = ops.bicubic_down_adj_raw(gy, in_h, in_w)
        acc = ops.bicubic_down_adj_raw(gy, in_h, in_w, gx=base.clone(), accumulate=True)
        assert torch.allclose(acc, base + fresh, atol=1e-6, rtol=0)


def test_cv2_resize_golden(golden):
    from artstyletransfer_b200 import ops
    gd = golden('small_ops.npz')
    up = ops.bicubic_resize(torch.from_numpy(gd['cv_up_in']).to(dev()), 64, 96, layout='hwc', coord='cv2')
    assert np.abs(up.cpu().numpy() - gd['cv_up_out']).max() < 5e-6
    down = ops.bicubic_resize(torch.from_numpy(gd['cv_down_in']).to(dev()), 9, 13, layout='hwc', coord='cv2')
    assert np.abs(down.cpu().numpy() - gd['cv_down_out']).max() < 5e-6


def test_resize_level_api(golden):
    import asyncio
    from artstyletransfer_b200 import neural_style_transfer as nst
    gd = golden('small_ops.npz')
    r0 = asyncio.run(nst.resize(gd['resize_in'], 0))
    assert tuple(r0.shape) == tuple(gd['resize_l0_shape']) and r0.dtype == np.float32
    assert np.abs(r0[::4, ::4] - gd['resize_l0_sub']).max() < 5e-6


CFG = {
    'default': dict(noise_factor=0.95, noise_levels=(9, 18, 36, -1, 0),
                    central=(0.30, 0.20, 0.10, 0.20, 0.20), peripheral=(0.20, 0.30, 0.40, 0.10, 0.00),
                    dispersion=(0.20, 0.30, 0.40, 0.60, 0.30)),
    'pixel': dict(noise_factor=0.5, noise_levels=(-1,), central=(1.0,), peripheral=(1.0,), dispersion=(0.5,)),
}
CASES = {
    'default_L1': dict(init_method='content+noise', levels=1, normal=False, cfg='default'),
    'default_L2': dict(init_method='content+noise', levels=2, normal=False, cfg='default'),
    'random_L1': dict(init_method='random', levels=1, normal=False, cfg='default'),
    'pixel_normal_L1': dict(init_method='content+noise', levels=1, normal=True, cfg='pixel'),
}


@pytest.mark.parametrize('name', list(CASES))
def test_noise_init_golden_and_oracle(golden, name):
    """Init image through the product's own pyramid + K7 path vs (i) the reference's output (golden) and
    (ii) the CPU oracle on the full image."""
    import asyncio
    from artstyletransfer_b200 import neural_style_transfer as nst
    gd = golden('noise_init.npz')
    case = CASES[name]; cfg = CFG[case['cfg']]
    top = case['levels'] - 1
    content_top = asyncio.run(nst.resize(gd['content'], top))
    style_top = asyncio.run(nst.resize(gd['style'], top))
    pair = nst.ContentStylePair(('c', gd['content']), ('s', gd['style']))
    nst.USE_NORMAL_NOISE_JUST_FOR_DEMONSTRATION = case['normal']
    try:
        np.random.seed(0)
        init, _ = nst.build_init_image(pair, [content_top], [style_top], case['init_method'], cfg['noise_factor'],
                                       cfg['noise_levels'], cfg['central'], cfg['peripheral'], cfg['dispersion'], dev())
    finally:
        nst.USE_NORMAL_NOISE_JUST_FOR_DEMONSTRATION = False
    assert init.dtype == np.float32 and tuple(init.shape) == tuple(gd[f'{name}_shape'])
    step = int(gd[f'{name}_step'])
    assert np.abs(init[::step, ::step] - gd[f'{name}_sub']).max() < 2e-5
    s = float(gd[f'{name}_sum'])
    assert abs(init.astype(np.float64).sum() - s) / abs(s) < 1e-6
    np.random.seed(0)
    ref = O.structured_noise_init(content_top, style_top, init_method=case['init_method'],
                                  use_normal_noise=case['normal'], **cfg)
    assert np.abs(init - ref).max() < 2e-5


def test_gaussian_mask_host(golden):
    from artstyletransfer_b200 import neural_style_transfer as nst
    gd = golden('small_ops.npz')
    assert np.abs(nst.gaussian_mask((40, 60, 3), 0.3, 0.2, 0.2) - gd['gmask']).max() < 1e-14


def test_cpu_tensor_is_rejected():
    from artstyletransfer_b200 import math_utils
    with pytest.raises(RuntimeError, match='CUDA'):
        math_utils.total_variation(torch.zeros(1, 3, 8, 8))
    with pytest.raises(RuntimeError, match='CUDA'):
        math_utils.gram_matrix(torch.zeros(1, 64, 8, 8))


@pytest.mark.parametrize('c,h,w', [(3, 256, 384), (3, 50, 70), (3, 34, 516), (1, 2, 2), (3, 2048, 3072), (2, 130, 258)])
def test_fused_pyramid_step_and_tv(c, h, w):
    """ast_bicubic_down2x_tv == ast_bicubic_down2x (bit for bit) + ast_tv_fwd of the same image (the sums are added in
    a different order: 1e-6), including partial tiles, unaligned widths and a reused workspace."""
    from artstyletransfer_b200 import ops
    dev = torch.device('cuda', 0)
    g = torch.Generator(device='cuda').manual_seed(h * 1000 + w)
    wss = ops.LevelWorkspaces()
    for rep in range(2):
        x = (torch.rand((1, c, h, w), generator=g, device=dev) * 255 - 120).contiguous()
        y_ref = ops.bicubic_down_raw(x, h // 2, w // 2)
        sums_ref = torch.empty(2, device=dev); tv_ref = torch.empty((), device=dev)
        ops.tv_fwd(x, sums_ref, tv_ref, ops.reduce_workspace(dev))
        y, sums2, tv = ops.bicubic_down_tv_raw(x, wss)
        assert torch.equal(y, y_ref)
        ref64 = [(x[..., :-1] - x[..., 1:]).double().abs().sum().item(), (x[..., :-1, :] - x[..., 1:, :]).double().abs().sum().item()]
        np.testing.assert_allclose(sums2.cpu().numpy(), ref64, rtol=2e-6)
        np.testing.assert_allclose(sums2.cpu().numpy(), sums_ref.cpu().numpy(), rtol=2e-6)
        if h > 2 and w > 2:
            assert abs(float(tv) - float(tv_ref)) <= 4e-6 * abs(float(tv_ref))
    # autograd wrapper: same gradient as the unfused step
    x1 = x.clone().requires_grad_(True); x2 = x.clone().requires_grad_(True)
    ya, _, _ = ops.bicubic_half_tv(x1, wss)
    yb = ops.bicubic_half(x2)
    u = torch.randn_like(ya)
    (ya * u).sum().backward(); (yb * u).sum().backward()
    assert torch.equal(x1.grad, x2.grad)
