"""Measure this box's TF32 / fp32 matmul throughput and the torch-library baselines for the Gram (run on a B200)."""
import json
import sys
import time

import torch


def bench(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def main():
    dev = torch.device('cuda', 0)
    out = {'gpu': torch.cuda.get_device_name(0)}
    n = 8192
    a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    out['tf32_tflops'] = 2 * n ** 3 / bench(lambda: a @ b) / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    out['fp32_tflops'] = 2 * n ** 3 / bench(lambda: a @ b, 3) / 1e9
    ab = a.bfloat16(); bb = b.bfloat16()
    out['bf16_tflops'] = 2 * n ** 3 / bench(lambda: ab @ bb) / 1e9
    x = torch.empty(1 << 28, device=dev); y = torch.empty_like(x)
    out['copy_gbs'] = 2 * x.numel() * 4 / bench(lambda: y.copy_(x)) / 1e6
    print(json.dumps(out))


if __name__ == '__main__':
    main()
