"""GPU probe: the bandwidth-bound resample / elementwise kernels on footprints larger than L2 (126 MB) — the only
shapes that count toward an HBM-fraction claim: K4/K5 on a batch of 48 planes of 2048x3072 (the
kernels take any number of planes), K6 general-ratio resize, content MSE, TV, the glue kernels.  CUDA events, L2
flushed between iterations.  One JSON line per kernel."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from artstyletransfer_b200 import ops
from gram_sweep import timeit
HBM = 6546.2
dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
CL = torch.channels_last


def report(name, nbytes, fn, **kw):
    t = timeit(fn, flush, iters=7)
    print(json.dumps(dict(kernel=name, MB=round(nbytes / 1e6, 1), ms=round(t, 4), GBps=round(nbytes / t / 1e6, 1),
                          frac_of_hbm=round(nbytes / t / 1e6 / HBM, 3), **kw)), flush=True)


for planes in (3, 48):
    x = torch.randn((planes, 2048, 3072), device=dev)
    y = [None]
    report('down2x', 4.0 * planes * 2048 * 3072 * 1.25, lambda: y.__setitem__(0, ops.bicubic_down_raw(x, 1024, 1536)), planes=planes)
    gx = torch.zeros_like(x)
    report('down2x_adj', 4.0 * planes * 2048 * 3072 * 1.25, lambda: ops.bicubic_down_adj_raw(y[0], 2048, 3072, gx=gx, accumulate=False), planes=planes)
    report('down2x_adj_accumulate', 4.0 * planes * 2048 * 3072 * 2.25, lambda: ops.bicubic_down_adj_raw(y[0], 2048, 3072, gx=gx, accumulate=True), planes=planes)
    del x, gx
x = torch.randn((48, 2047, 3071), device=dev)       # odd size -> general-ratio kernel pair
report('resize_chw_torch_coords', 4.0 * 48 * (2047 * 3071 + 1023 * 1535), lambda: ops.bicubic_down_raw(x, 1023, 1535), planes=48)
del x
img = torch.rand((4096, 6144, 3), device=dev)
report('resize_hwc_cv2_down', 4.0 * 3 * (4096 * 6144 + 2048 * 3072), lambda: ops.bicubic_resize(img, 2048, 3072, 'hwc', 'cv2'))
small = torch.rand((1024, 1536, 3), device=dev)
report('resize_hwc_cv2_up', 4.0 * 3 * (1024 * 1536 + 4096 * 6144), lambda: ops.bicubic_resize(small, 4096, 6144, 'hwc', 'cv2'))
del img, small
for n in (50331648, 201326592):
    a = torch.randn(n, device=dev); b = torch.randn(n, device=dev); d = torch.zeros(n, device=dev)
    loss = torch.empty((), device=dev); rws = ops.reduce_workspace(dev)
    report('mse_fwd', 8.0 * n, lambda: ops.mse_fwd(a, b, 1.0 / n, loss, rws), n=n)
    report('mse_bwd', 12.0 * n, lambda: ops.mse_bwd(a, b, 2.0 / n, None, d, False), n=n)
    report('mse_bwd_accumulate', 16.0 * n, lambda: ops.mse_bwd(a, b, 2.0 / n, None, d, True), n=n)
    del a, b, d
for planes in (3, 48):
    img = torch.randn((1, planes, 2048, 3072), device=dev); dimg = torch.empty_like(img)
    s2 = torch.empty(2, device=dev); tv = torch.empty((), device=dev); rws = ops.reduce_workspace(dev)
    report('tv_fwd', 4.0 * img.numel(), lambda: ops.tv_fwd(img, s2, tv, rws), planes=planes)
    report('tv_bwd', 8.0 * img.numel(), lambda: ops.tv_bwd(img, s2, 1e2, None, dimg, False), planes=planes)
    del img, dimg
y = torch.randn((1, 64, 2048, 3072), device=dev).contiguous(memory_format=CL); bias = torch.randn(64, device=dev)
g = torch.randn((1, 64, 2048, 3072), device=dev).contiguous(memory_format=CL)
report('bias_relu', 8.0 * y.numel(), lambda: ops.bias_relu_(y, bias))
report('relu_bwd', 12.0 * y.numel(), lambda: ops.relu_bwd_(g, y))
p = torch.empty((1, 64, 1024, 1536), device=dev, memory_format=CL)
report('maxpool2x2', 5.0 * y.numel(), lambda: ops.maxpool2x2(y, p))
gp = torch.randn((1, 64, 1024, 1536), device=dev).contiguous(memory_format=CL)
report('maxpool2x2_relu_bwd', 9.0 * y.numel(), lambda: ops.maxpool2x2_bwd(gp, y, g, True))
