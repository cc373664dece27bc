"""BASELINE config 5: Gram fwd+bwd kernel sweep, C in {64,128,256,512} x HW of every VGG layer at every
pyramid level, on one B200.  CUDA events, L2 flushed between timed iterations (256 MB write).
Prints one JSON line per (C, HW): this library's kernels (TF32 tcgen05 and exact fp32) and the torch library
path the reference uses (bmm + MSELoss + autograd) as the bar to beat.  Usage: python tests/tools/gram_sweep.py [--quick]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from artstyletransfer_b200 import ops  # noqa: E402

HBM, TF32 = 6546.2, 744.9


def timeit(fn, flush, iters=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    quick = '--quick' in sys.argv
    dev = torch.device('cuda', 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = []
    for lvl in ([3, 1] if quick else [3, 2, 1, 0]):
        for c, base in ((64, 98304), (128, 24576), (256, 6144), (512, 1536), (512, 384)):
            shapes.append((c, base * 4 ** lvl))
    for c, hw in shapes:
        g = torch.Generator(device='cuda').manual_seed(c + hw)
        f = torch.relu(torch.randn((c, hw), generator=g, device=dev)) * 0.25
        a = torch.rand((c, c), device=dev) * 1e-3
        d = torch.empty((c, c), device=dev); loss = torch.empty((), device=dev)
        df = torch.empty_like(f)
        ws = ops.gram_workspace(c, hw, dev)
        row = {'C': c, 'HW': hw, 'F_MB': round(c * hw * 4 / 1e6, 1)}
        for name, prec in (('tf32', 0), ('fp32', 1)):
            if prec == 1 and c * hw > 64 * 6291456 // 4 and quick:
                continue
            tf = timeit(lambda: ops.gram_mse_fwd(f, c, hw, 1.0 / (c * hw), a, d, loss, ws, prec), flush)
            tb = timeit(lambda: ops.gram_bwd(d, f, c, hw, 1e-3, None, df, False, prec), flush)
            fl = 2.0 * c * c * hw
            bf, bb = 4.0 * c * hw + 8.0 * c * c, 8.0 * c * hw + 4.0 * c * c
            row[name] = {'fwd_ms': round(tf, 4), 'bwd_ms': round(tb, 4),
                         'fwd_GBps': round(bf / tf / 1e6, 1), 'bwd_GBps': round(bb / tb / 1e6, 1),
                         'fwd_TFLOPs': round(fl / tf / 1e9, 1), 'bwd_TFLOPs': round(fl / tb / 1e9, 1),
                         'fwd_frac_of_bound': round(max(bf / (HBM * 1e9), fl / (TF32 * 1e12)) / (tf * 1e-3), 3),
                         'bwd_frac_of_bound': round(max(bb / (HBM * 1e9), fl / (TF32 * 1e12)) / (tb * 1e-3), 3),
                         'fwdbwd_TFLOPs': round(2 * fl / (tf + tb) / 1e9, 1)}
        # torch library path (what the reference runs on a GPU): bmm + /= + MSELoss, autograd backward
        x = f.view(1, c, 1, hw).clone().requires_grad_(True)
        for tf32_flag in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32_flag

            def lib_fwd():
                feats = x.view(1, c, hw)
                gm = feats.bmm(feats.transpose(1, 2))
                gm = gm / (c * hw)
                return torch.nn.functional.mse_loss(a, gm[0])

            def lib_fwdbwd():
                x.grad = None
                lib_fwd().backward()
            t_f = timeit(lambda: lib_fwd(), flush, 3)
            t_fb = timeit(lib_fwdbwd, flush, 3)
            row['torch_tf32' if tf32_flag else 'torch_fp32'] = {'fwd_ms': round(t_f, 4), 'fwdbwd_ms': round(t_fb, 4)}
        torch.backends.cuda.matmul.allow_tf32 = False
        if 'tf32' in row:
            ours = row['tf32']['fwd_ms'] + row['tf32']['bwd_ms']
            row['speedup_vs_torch_fp32'] = round(row['torch_fp32']['fwdbwd_ms'] / ours, 2)
            row['speedup_vs_torch_tf32'] = round(row['torch_tf32']['fwdbwd_ms'] / ours, 2)
        print(json.dumps(row), flush=True)
        del f, df, x


if __name__ == '__main__':
    main()
