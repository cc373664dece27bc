"""GPU probe: cuDNN conv + own bias/ReLU kernel vs torch.cudnn_convolution_relu (fused epilogue), channels_last TF32."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from artstyletransfer_b200 import ops
CL = torch.channels_last
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for (cin, cout, h, w) in ((3, 64, 2048, 3072), (64, 64, 2048, 3072), (64, 128, 1024, 1536), (128, 128, 1024, 1536), (128, 256, 512, 768),
                          (256, 256, 512, 768), (256, 512, 256, 384), (512, 512, 256, 384), (512, 512, 128, 192)):
    x = torch.randn((1, cin, h, w), device='cuda').contiguous(memory_format=CL)
    wt = (torch.randn((cout, cin, 3, 3), device='cuda') * 0.05).contiguous(memory_format=CL)
    b = torch.randn(cout, device='cuda')
    def ours():
        y = torch.ops.aten.cudnn_convolution(x, wt, [1, 1], [1, 1], [1, 1], 1, False, False, True)
        ops.bias_relu_(y, b)
        return y
    def conv_only():
        return torch.ops.aten.cudnn_convolution(x, wt, [1, 1], [1, 1], [1, 1], 1, False, False, True)
    def fused():
        return torch.cudnn_convolution_relu(x, wt, b, [1, 1], [1, 1], [1, 1], 1)
    t0, t1 = timeit(conv_only), timeit(ours)
    try:
        t2 = timeit(fused)
        y1, y2 = ours(), fused()
        err = float((y1 - y2).abs().max() / y1.abs().max())
        cl = y2.is_contiguous(memory_format=CL)
    except Exception as e:
        t2, err, cl = None, repr(e)[:80], None
    g = torch.randn((1, cout, h, w), device='cuda').contiguous(memory_format=CL)
    t3 = timeit(lambda: torch.ops.aten.convolution_backward(g, x, wt, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False]))
    fl = 2.0 * 9 * cin * cout * h * w
    print(json.dumps(dict(cin=cin, cout=cout, h=h, w=w, conv_ms=round(t0, 4), conv_TF=round(fl / t0 / 1e9), conv_plus_bias_relu_ms=round(t1, 4),
                          fused_ms=t2 and round(t2, 4), fused_cl=cl, err=err, dgrad_ms=round(t3, 4), dgrad_TF=round(fl / t3 / 1e9))), flush=True)
