"""GPU probe (not a test): where the torch/cuDNN share of a closure goes, NCHW vs channels_last, band sizes."""
import json, sys, time
import torch, torchvision

def vgg(seed=1234):
    torch.manual_seed(seed)
    m = torchvision.models.vgg19(weights=None).features[:30].cuda().eval()
    for p in m.parameters():
        p.requires_grad = False
    return m

def run(m, x, taps=(1, 6, 11, 20, 22, 29)):
    feats = []
    for i, l in enumerate(m):
        x = l(x)
        if i in taps:
            feats.append(x)
    return feats

def timeit(fn, n=5, w=2):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def closure(m, img, cl):
    def f():
        img.grad = None
        x = img.contiguous(memory_format=torch.channels_last) if cl else img
        fs = run(m, x)
        loss = sum((f_ * f_).mean() for f_ in fs)   # stand-in loss touching every tap
        loss.backward()
    return f

out = []
m = vgg()
for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    for (h, w) in ((2048, 3072), (1024, 1536), (512, 768), (256, 384), (416, 3072), (336, 3072), (288, 3072), (258, 3072)):
        for cl in (False, True):
            mm = m.to(memory_format=torch.channels_last) if cl else m.to(memory_format=torch.contiguous_format)
            img = torch.randn(1, 3, h, w, device='cuda', requires_grad=True)
            try:
                t = timeit(closure(mm, img, cl))
            except Exception as e:
                t = repr(e)[:100]
            rec = dict(cudnn_benchmark=bench, h=h, w=w, channels_last=cl, ms=t,
                       peak_gb=torch.cuda.max_memory_allocated() / 1e9)
            print(json.dumps(rec), flush=True)
            torch.cuda.reset_peak_memory_stats()
try:
    import torch.distributed._symmetric_memory as sm
    print('symm_mem available', [n for n in dir(sm) if not n.startswith('_')][:40])
except Exception as e:
    print('symm_mem import failed', e)
