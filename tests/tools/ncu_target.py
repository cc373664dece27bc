"""Small target for `ncu --set full`: one forward + backward Gram launch per VGG width at the L=3 sizes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from artstyletransfer_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
shapes = [(64, 6291456), (128, 1572864), (256, 393216), (512, 98304)]
if len(sys.argv) > 1:
    shapes = [s for s in shapes if str(s[0]) in sys.argv[1:]]
for c, hw in shapes:
    f = torch.relu(torch.randn((c, hw), device=dev)) * 0.25
    a = torch.rand((c, c), device=dev) * 1e-3
    d = torch.empty((c, c), device=dev); loss = torch.empty((), device=dev); df = torch.empty_like(f)
    ws = ops.gram_workspace(c, hw, dev)
    for _ in range(2):
        ops.gram_mse_fwd(f, c, hw, 1.0 / (c * hw), a, d, loss, ws, 0)
        ops.gram_bwd(d, f, c, hw, 1e-3, None, df, False, 0)
    torch.cuda.synchronize()
    print(c, hw, float(loss))
