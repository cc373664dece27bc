"""How reproducible is the 50-step Adam trajectory of the ORACLE itself (BASELINE: final image PSNR >= 40 dB)?

Runs the oracle's torch closure (oracle/gatys_oracle.py, restating neural_style_transfer.py:152-202 of the reference)
under torch Adam for 50 steps on the 64x96 two-level job of tests/test_gpu_closure.py and compares end images of
  * float32 arithmetic (the reference's),
  * float64 arithmetic,
  * float32 from a start image perturbed by 1e-5 (relative to the [0,1] range),
  * float32 with every gradient perturbed by 1e-4 relative noise (an implementation that exactly meets the stated
    1e-4 loss / gradient tolerance),
at the reference's lr_start = 10 and at lr_start = 1.  CPU or GPU (--device cuda).  Output: one JSON line per lr.
The test computes the same floor on the device it runs on (tests/test_gpu_closure.py)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gatys_oracle as O  # noqa: E402

WEIGHTS = (1e3, 4e5, 1e2)


def oracle_adam(dev, lr, dtype=torch.float32, start_eps=0.0, grad_eps=0.0, steps=50, seed=1):
    content, style = O.synthetic_images(64, 96, seed=seed)
    c_lv = [content, O.bicubic_resize_hwc(content, 48, 32).astype(np.float32)]
    s_lv = [style, O.bicubic_resize_hwc(style, 48, 32).astype(np.float32)]
    init = np.clip(content * 0.6 + np.random.default_rng(4).uniform(0, 1, size=content.shape) * 0.4, 0, 1).astype(np.float32)
    if start_eps:
        init = (init + start_eps * np.random.default_rng(99).standard_normal(init.shape)).astype(np.float32)
    net, cidx, sidx = O.make_vgg19(1234)
    net = net.to(dev).to(dtype)
    targets = [O.torch_targets(net, cidx, sidx, torch.from_numpy(O.prepare_img(c)).to(dev).to(dtype),
                               torch.from_numpy(O.prepare_img(s)).to(dev).to(dtype)) for c, s in zip(c_lv, s_lv)]
    img = torch.from_numpy(O.prepare_img(init)).to(dev).to(dtype).requires_grad_(True)
    opt = torch.optim.Adam((img,), lr=lr)
    gen = torch.Generator(device='cpu').manual_seed(7)
    for _ in range(steps):
        for g in opt.param_groups:
            g['lr'] *= 0.999
        opt.zero_grad()
        _, _, grad = O.torch_closure(net, cidx, sidx, targets, img, WEIGHTS)
        if grad_eps:
            noise = torch.randn(grad.shape, generator=gen).to(grad)
            grad = grad + grad_eps * noise * (grad.norm() / noise.norm())
        img.grad = grad
        opt.step()
    return O.unprepare_img(img.detach().float().cpu().numpy())


def floors(dev, lr):
    base = oracle_adam(dev, lr)
    return base, {
        'fp32_vs_fp64_dB': round(O.psnr(base, oracle_adam(dev, lr, dtype=torch.float64)), 2),
        'start_perturbed_1e-5_dB': round(O.psnr(base, oracle_adam(dev, lr, start_eps=1e-5)), 2),
        'grad_perturbed_1e-4_dB': round(O.psnr(base, oracle_adam(dev, lr, grad_eps=1e-4)), 2),
    }


def product_adam(dev, lr, precision=None, steps=50, seed=1):
    """The product's NeuralStyleTransfer.process on the same job (GPU only)."""
    import asyncio
    import torchvision
    from artstyletransfer_b200 import neural_nets, neural_style_transfer as nst
    real = torchvision.models.vgg19
    neural_nets.models.vgg19 = lambda pretrained=False, progress=False, **kw: (torch.manual_seed(1234), real(weights=None))[1]
    content, style = O.synthetic_images(64, 96, seed=seed)
    c_lv = [content, O.bicubic_resize_hwc(content, 48, 32).astype(np.float32)]
    s_lv = [style, O.bicubic_resize_hwc(style, 48, 32).astype(np.float32)]
    init = np.clip(content * 0.6 + np.random.default_rng(4).uniform(0, 1, size=content.shape) * 0.4, 0, 1).astype(np.float32)
    nst.PRECISION = precision

    async def run():
        drv = nst.NeuralStyleTransfer(dev, 'vgg19', s_lv, 'adam')
        last = None
        async for img, step in drv.process(c_lv, init, lr, steps, *WEIGHTS, 'psnr'):
            last = img
        return np.array(last, copy=True)
    try:
        return asyncio.run(run())
    finally:
        nst.PRECISION = None
        neural_nets.models.vgg19 = real


def product_grad_error_along_trajectory(dev, lr, precision='tf32', autotune=False, steps=50, seed=1):
    """Relative error of the product's image gradient against the oracle's float64 gradient AT THE PRODUCT'S OWN
    iterates (steps 0, 10, 25, 49): how big is the per-step perturbation the chaotic trajectory amplifies?"""
    import torchvision
    from artstyletransfer_b200 import feature_path, neural_nets, neural_style_transfer as nst
    real = torchvision.models.vgg19
    neural_nets.models.vgg19 = lambda pretrained=False, progress=False, **kw: (torch.manual_seed(1234), real(weights=None))[1]
    content, style = O.synthetic_images(64, 96, seed=seed)
    c_lv = [content, O.bicubic_resize_hwc(content, 48, 32).astype(np.float32)]
    s_lv = [style, O.bicubic_resize_hwc(style, 48, 32).astype(np.float32)]
    init = np.clip(content * 0.6 + np.random.default_rng(4).uniform(0, 1, size=content.shape) * 0.4, 0, 1).astype(np.float32)
    nst.PRECISION = precision
    old_bm, feature_path.CUDNN_BENCHMARK = feature_path.CUDNN_BENCHMARK, autotune
    try:
        job = nst._Job(dev, 'vgg19', s_lv, 'adam', c_lv, init, lr, *WEIGHTS, 'graderr')
        net, cidx, sidx = O.make_vgg19(1234)
        net64 = net.to(dev).double()
        t64 = [O.torch_targets(net64, cidx, sidx, torch.from_numpy(O.prepare_img(c)).to(dev).double(),
                               torch.from_numpy(O.prepare_img(s)).to(dev).double()) for c, s in zip(c_lv, s_lv)]
        net32, _, _ = O.make_vgg19(1234)
        net32 = net32.to(dev)
        t32 = [O.torch_targets(net32, cidx, sidx, torch.from_numpy(O.prepare_img(c)).to(dev),
                               torch.from_numpy(O.prepare_img(s)).to(dev)) for c, s in zip(c_lv, s_lv)]
        out = {}
        for k in range(steps):
            if k in (0, 10, 25, 49):
                x = job.optimizing_img.detach().clone()
                l64, _, g64 = O.torch_closure(net64, cidx, sidx, t64, x.double(), WEIGHTS)
                l32, _, g32 = O.torch_closure(net32, cidx, sidx, t32, x, WEIGHTS)
                job.optimizer.zero_grad()
                lp = job._evaluate()
                gp = job.optimizing_img.grad.double()
                out[k] = {'product_grad_err': float((gp - g64).norm() / g64.norm()),
                          'oracle_fp32_grad_err': float((g32.double() - g64).norm() / g64.norm()),
                          'product_loss_err': abs(float(lp) - float(l64)) / float(l64),
                          'oracle_fp32_loss_err': abs(float(l32) - float(l64)) / float(l64), 'loss': float(l64)}
                job.optimizing_img.grad = None
            job.optimizer_step()
        final = O.unprepare_img(job.optimizing_img.detach().cpu().numpy())
        return out, final
    finally:
        nst.PRECISION = None
        feature_path.CUDNN_BENCHMARK = old_bm
        neural_nets.models.vgg19 = real


if __name__ == '__main__':
    dev = torch.device('cuda', 0) if '--device' in sys.argv and 'cuda' in sys.argv else torch.device('cpu')
    if dev.type == 'cuda':
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.deterministic = True
    for lr in (10.0, 1.0):
        base, f = floors(dev, lr)
        if dev.type == 'cuda' and '--product' in sys.argv:
            from artstyletransfer_b200 import feature_path
            for autotune in (False, True):
                feature_path.CUDNN_BENCHMARK = autotune
                for prec in ('tf32', 'fp32'):
                    f[f'product_{prec}_{"search" if autotune else "heur"}_dB'] = round(O.psnr(product_adam(dev, lr, prec), base), 2)
            feature_path.CUDNN_BENCHMARK = False
        if dev.type == 'cuda' and '--graderr' in sys.argv:
            for autotune in (False, True):
                traj, final = product_grad_error_along_trajectory(dev, lr, 'tf32', autotune)
                f[f'tf32_{"search" if autotune else "heur"}_trajectory'] = traj
                f[f'tf32_{"search" if autotune else "heur"}_direct_dB'] = round(O.psnr(final, base), 2)
        print(json.dumps({'device': str(dev), 'lr_start': lr, 'steps': 50, **f}), flush=True)
