"""CPU probe (oracle + numpy, no GPU): what rounding the Gram OPERANDS to TF32 (nearest / truncated) or BF16 does to
the per-layer Gram and to the style loss on random-init VGG19 features of the synthetic 256x384 level — the parity
budget check behind the converter warps' round-to-nearest and behind the BF16-mode decision."""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import gatys_oracle as O
torch.set_num_threads(8)
net,cidx,sidx=O.make_vgg19(1234)
content,style=O.synthetic_images(256,384,seed=0)
init=np.clip(content*0.5+np.random.default_rng(1).uniform(0,1,content.shape)*0.5,0,1).astype(np.float32)
def feats(img):
    with torch.no_grad():
        return [f[0].reshape(f.shape[1],-1).double() for f in (net(torch.from_numpy(O.prepare_img(img)))[k] for k in sidx)]
fx, fs = feats(init), feats(style)
def rnd(t, kind):
    if kind=='exact': return t
    f=t.float()
    if kind=='bf16': return f.bfloat16().double()
    if kind=='tf32':   # round to nearest, 10 explicit mantissa bits
        i=f.view(torch.int32); i=(i+0x1000) & ~0x1FFF; return i.view(torch.float32).double()
    if kind=='tf32_trunc':
        i=f.view(torch.int32) & ~0x1FFF; return i.view(torch.float32).double()
res={}
for kind in ('exact','tf32','tf32_trunc','bf16'):
    losses=[]; fro=[]
    for x,s in zip(fx,fs):
        c,hw=x.shape
        A=(s@s.t())/(c*s.shape[1])          # target from exact style features
        xr=rnd(x,kind)
        G=(xr@xr.t())/(c*hw)
        Gex=(x@x.t())/(c*hw)
        fro.append(float((G-Gex).norm()/Gex.norm()))
        losses.append(float(((G-A)**2).mean()))
    res[kind]=(sum(losses)/5, fro)
ex=res['exact'][0]
for k,(l,fro) in res.items():
    print(f'{k:10s} style loss rel err {abs(l-ex)/ex:.3e}   Gram Frobenius rel err per layer', ' '.join(f'{e:.1e}' for e in fro))
