"""GPU probe: channels-last Gram kernels at the BASELINE shapes (CUDA events, L2 flushed), beside the NCHW family."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from artstyletransfer_b200 import ops
from gram_sweep import timeit
dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for c, hw in ((64, 6291456), (128, 1572864), (256, 393216), (512, 98304), (64, 1572864), (128, 393216), (512, 24576), (512, 1536)):
    gen = torch.Generator(device='cuda').manual_seed(c + hw)
    f = torch.relu(torch.randn((hw, c), generator=gen, device=dev)) * 0.25
    fn = f.t().contiguous()
    a = torch.rand((c, c), device=dev) * 1e-3
    d = torch.empty((c, c), device=dev); loss = torch.empty((), device=dev)
    df = torch.empty_like(f); dfn = torch.empty_like(fn)
    ws = ops.gram_workspace(c, hw, dev)
    byf, byb = 4.0 * c * hw + 8.0 * c * c, 8.0 * c * hw + 4.0 * c * c
    fl = 2.0 * c * c * hw
    t1 = timeit(lambda: ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), a, d, loss, ws), flush, iters=7)
    t2 = timeit(lambda: ops.gram_bwd_nhwc(d, f, c, hw, 1e-3, None, df, False), flush, iters=7)
    t3 = timeit(lambda: ops.gram_bwd_nhwc(d, f, c, hw, 1e-3, None, df, True), flush, iters=7)
    dr = torch.empty((c, c), device=dev)
    ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), a, dr, loss, ws, round_out=True)
    t6 = timeit(lambda: ops.gram_bwd_nhwc(dr, f, c, hw, 1e-3, None, df, False, d_prerounded=True), flush, iters=7)
    t7 = timeit(lambda: ops.gram_bwd_nhwc(dr, f, c, hw, 1e-3, None, df, True, d_prerounded=True), flush, iters=7)
    t8 = timeit(lambda: ops.gram_bwd_nhwc(dr, f, c, hw, 1e-3, None, df, True, d_prerounded=True, relu_mask=True), flush, iters=7)
    bf = {}
    if c == 512:      # AST_PREC_BF16: bfloat16 operands in the 512-channel backward (D from the finalize kernel as bf16)
        dbf = torch.empty((c, c), dtype=torch.bfloat16, device=dev)
        ops.gram_mse_fwd_nhwc(f, c, hw, 1.0 / (c * hw), a, dbf, loss, ws, round_out=2)
        t9 = timeit(lambda: ops.gram_bwd_nhwc_auto(dbf, f, c, hw, 1e-3, None, df, False), flush, iters=7)
        t10 = timeit(lambda: ops.gram_bwd_nhwc_auto(dbf, f, c, hw, 1e-3, None, df, True), flush, iters=7)
        bf = {'bf16_bwd_ms': round(t9, 4), 'bf16_bwd_TF': round(fl / t9 / 1e9), 'bf16_bwd_GBps': round((byb - 2.0 * c * c) / t9 / 1e6),
              'bf16_bwd_acc_ms': round(t10, 4), 'bf16_bwd_acc_GBps': round((byb - 2.0 * c * c + 4.0 * c * hw) / t10 / 1e6)}
    t4 = timeit(lambda: ops.gram_mse_fwd(fn, c, hw, 1.0 / (c * hw), a, d, loss, ws, 0), flush, iters=7)
    t5 = timeit(lambda: ops.gram_bwd(d, fn, c, hw, 1e-3, None, dfn, False, 0), flush, iters=7)
    print(json.dumps({'C': c, 'HW': hw, 'noround': os.environ.get('AST_GRAM_FWD_NOROUND'),
                      'nhwc_fwd_ms': round(t1, 4), 'nhwc_fwd_GBps': round(byf / t1 / 1e6), 'nhwc_fwd_TF': round(fl / t1 / 1e9),
                      'nhwc_bwd_ms': round(t2, 4), 'nhwc_bwd_GBps': round(byb / t2 / 1e6), 'nhwc_bwd_TF': round(fl / t2 / 1e9),
                      'nhwc_bwd_acc_ms': round(t3, 4), 'nhwc_bwd_acc_GBps': round((byb + 4.0 * c * hw) / t3 / 1e6),
                      'nhwc_bwd_prer_ms': round(t6, 4), 'nhwc_bwd_prer_GBps': round(byb / t6 / 1e6), 'nhwc_bwd_prer_TF': round(fl / t6 / 1e9),
                      'nhwc_bwd_prer_acc_ms': round(t7, 4), 'nhwc_bwd_prer_acc_GBps': round((byb + 4.0 * c * hw) / t7 / 1e6),
                      'nhwc_bwd_acc_relu_ms': round(t8, 4), 'nhwc_bwd_acc_relu_GBps': round((byb + 4.0 * c * hw) / t8 / 1e6),
                      'nchw_fwd_ms': round(t4, 4), 'nchw_fwd_GBps': round(byf / t4 / 1e6),
                      'nchw_bwd_ms': round(t5, 4), 'nchw_bwd_GBps': round(byb / t5 / 1e6), **bf}), flush=True)
