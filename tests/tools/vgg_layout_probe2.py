"""GPU probe: wall (CUDA-event) time of VGG19[:30] fwd+bwd with a plain sum loss, NCHW vs channels_last."""
import json, torch, torchvision
torch.manual_seed(1234)
base = torchvision.models.vgg19(weights=None).features[:30].cuda().eval()
for p in base.parameters():
    p.requires_grad = False
for (h, w) in ((2048, 3072), (1024, 1536), (256, 3072)):
    for cl in (False, True):
        m = base.to(memory_format=torch.channels_last if cl else torch.contiguous_format)
        img = torch.randn(1, 3, h, w, device='cuda', requires_grad=True)
        def step():
            img.grad = None
            x = img.contiguous(memory_format=torch.channels_last) if cl else img
            m(x).sum().backward()
        for _ in range(2): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5): step()
        e1.record(); torch.cuda.synchronize()
        print(json.dumps(dict(h=h, w=w, channels_last=cl, ms=e0.elapsed_time(e1) / 5)), flush=True)
