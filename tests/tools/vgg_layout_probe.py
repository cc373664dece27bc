"""GPU probe (not a test): one VGG19[:30] forward+backward, NCHW or channels_last (argv[1] = nchw|cl), for an ncu launch list."""
import sys
import torch, torchvision
cl = sys.argv[1] == 'cl'
h, w = int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(1234)
m = torchvision.models.vgg19(weights=None).features[:30].cuda().eval()
for p in m.parameters():
    p.requires_grad = False
if cl:
    m = m.to(memory_format=torch.channels_last)
img = torch.randn(1, 3, h, w, device='cuda', requires_grad=True)
for it in range(3):
    img.grad = None
    x = img.contiguous(memory_format=torch.channels_last) if cl else img
    y = m(x)
    y.sum().backward()
torch.cuda.synchronize()
print('done')
