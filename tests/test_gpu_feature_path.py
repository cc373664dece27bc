"""GPU parity of the channels-last feature path: the glue kernels against torch's own ops (what the reference
runs), the taps against Vgg19.forward, and one level's loss + image gradient against the torch/autograd path."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
CL = torch.channels_last


def dev():
    return torch.device('cuda', 0)


def rel(a, b):
    a = a.detach().double().cpu().numpy(); b = b.detach().double().cpu().numpy()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def cl_randn(c, h, w, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.randn((1, c, h, w), generator=g, device=dev()).contiguous(memory_format=CL)


@pytest.mark.parametrize('c,h,w', [(64, 8, 12), (128, 33, 20), (512, 5, 7), (4, 2, 2)])
def test_bias_relu_and_relu_bwd_bit_exact(c, h, w):
    from artstyletransfer_b200 import ops
    y = cl_randn(c, h, w, 1)
    b = torch.randn(c, device=dev())
    ref = torch.relu(y + b.view(1, c, 1, 1))
    out = y.clone(memory_format=torch.preserve_format)
    ops.bias_relu_(out, b)
    assert torch.equal(out, ref)
    g = cl_randn(c, h, w, 2)
    gref = torch.where(ref > 0, g, torch.zeros_like(g))
    ops.relu_bwd_(g, out)
    assert torch.equal(g, gref)


@pytest.mark.parametrize('c,h,w', [(64, 8, 12), (128, 33, 21), (256, 6, 9), (4, 2, 2), (512, 16, 24)])
@pytest.mark.parametrize('relu_mask', [False, True])
def test_maxpool_fwd_bwd_matches_torch(c, h, w, relu_mask):
    """Forward bit-exact; backward equal to torch's max_pool2d backward (+ threshold backward when fused),
    including ties (post-ReLU zeros and duplicated maxima) and odd sizes (floor mode)."""
    from artstyletransfer_b200 import ops
    x = torch.relu(cl_randn(c, h, w, 3))
    x[:, :, ::3, ::2] = 0.5                                   # many exact ties
    if not relu_mask:
        x = x - 0.25                                          # negative maxima must still route the gradient
    xr = x.clone(memory_format=torch.preserve_format).requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    y = torch.empty((1, c, h // 2, w // 2), device=dev()).contiguous(memory_format=CL)
    ops.maxpool2x2(x, y)
    assert torch.equal(y, yr.detach())
    gy = cl_randn(c, h // 2, w // 2, 4)
    gx = torch.full_like(x, float('nan'), memory_format=torch.preserve_format)
    ops.maxpool2x2_bwd(gy, x, gx, relu_mask)
    (gref,) = torch.autograd.grad(yr, xr, gy)
    if relu_mask:
        gref = torch.where(x > 0, gref, torch.zeros_like(gref))
    assert not torch.isnan(gx).any()
    assert torch.equal(gx, gref)
    # and against the numpy oracle (H, W, C arrays)
    from oracle import gatys_oracle as O
    hwc = lambda t: t[0].permute(1, 2, 0).cpu().numpy()
    assert np.array_equal(hwc(y), O.maxpool2x2_hwc(hwc(x)))
    assert np.array_equal(hwc(gx), O.maxpool2x2_bwd_hwc(hwc(gy), hwc(x), relu_mask))


def test_image_layout_roundtrip():
    from artstyletransfer_b200 import ops
    img = torch.randn((1, 3, 37, 52), device=dev())
    x = torch.empty((1, 3, 37, 52), device=dev()).contiguous(memory_format=CL)
    ops.chw_to_hwc(img, x, 3, 37 * 52)
    assert torch.equal(x, img)                                # same logical tensor, channels_last strides
    back = torch.ones_like(img)
    ops.hwc_to_chw(x, back, 3, 37 * 52, True)
    assert torch.equal(back, img + 1)


@pytest.fixture()
def vgg(seeded_vgg):
    from artstyletransfer_b200 import math_utils
    return math_utils.prepare_model('vgg19', dev())


@pytest.mark.parametrize('h,w', [(64, 96), (80, 112), (50, 70)])
def test_taps_match_vgg19_forward(vgg, h, w):
    from artstyletransfer_b200 import feature_path
    net, cidx, sidx = vgg
    plan = feature_path.plan_for(net)
    assert plan is not None
    img = torch.randn((1, 3, h, w), device=dev()) * 50
    with torch.no_grad():
        ref = net(img)
        taps, _ = feature_path.features_forward(plan, img, keep=False)
    for k, (a, b) in enumerate(zip(taps, ref)):
        assert tuple(a.shape) == tuple(b.shape)
        assert rel(a, b) < 2e-3, k                             # both TF32 convolutions, different kernels


@pytest.mark.parametrize('h,w', [(64, 96), (48, 80)])
def test_level_loss_and_gradient_match_autograd_path(vgg, h, w):
    """The explicit schedule vs torch modules + autograd around the NCHW kernels, same network, fp32-exact convs."""
    from artstyletransfer_b200 import neural_style_transfer as nst
    net, cidx, sidx = vgg
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        g = torch.Generator(device='cuda').manual_seed(5)
        content = torch.rand((1, 3, h, w), generator=g, device=dev()) * 255 - 120
        style = torch.rand((1, 3, h + 16, w), generator=g, device=dev()) * 255 - 120
        init = torch.rand((1, 3, h, w), generator=g, device=dev()) * 255 - 120
        res = {}
        for flag in (False, True):
            nst.CHANNELS_LAST_PATH = flag
            lb = nst.LossBuilder(cidx, sidx, content, style, net, 1e3, 4e5, 1e2)
            img = init.clone().requires_grad_(True)
            t, c, s, v = lb.build(img)
            t.backward()
            res[flag] = (t.item(), c.item(), s.item(), v.item(), img.grad.clone())
    finally:
        nst.CHANNELS_LAST_PATH = True
        torch.backends.cudnn.allow_tf32 = old
    a, b = res[True], res[False]
    for i in range(4):
        assert abs(a[i] - b[i]) <= 1e-4 * abs(b[i]) + 1e-12, (i, a[i], b[i])
    assert rel(a[4], b[4]) < 2e-3


def test_level_without_grad_and_rerun_is_bit_identical(vgg):
    from artstyletransfer_b200 import neural_style_transfer as nst
    net, cidx, sidx = vgg
    g = torch.Generator(device='cuda').manual_seed(9)
    content = torch.rand((1, 3, 64, 96), generator=g, device=dev()) * 255 - 120
    style = torch.rand((1, 3, 64, 64), generator=g, device=dev()) * 255 - 120
    lb = nst.LossBuilder(cidx, sidx, content, style, net, 1e3, 4e5, 1e2)
    old = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    try:
        outs = []
        for _ in range(2):
            img = content.clone().requires_grad_(True)
            t, *_ = lb.build(img)
            t.backward()
            outs.append((t.item(), img.grad.clone()))
        with torch.no_grad():
            t0, *_ = lb.build(content.clone())
    finally:
        torch.backends.cudnn.deterministic = old
    assert outs[0][0] == outs[1][0] == t0.item()
    assert torch.equal(outs[0][1], outs[1][1])
