"""GPU tuning probe: forward Gram kernel pipeline shapes (AST_GRAM_FWD_CFG=NU,G is read once per process, so
each shape runs in its own process: python tests/tools/fwd_cfg_sweep.py NU G)."""
import json, os, sys
nu, g = sys.argv[1], sys.argv[2]
os.environ['AST_GRAM_FWD_CFG'] = f'{nu},{g}'
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from artstyletransfer_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gram_sweep import timeit
dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = {}
for c, hw in ((64, 6291456), (128, 1572864), (256, 393216), (512, 98304), (64, 1572864), (128, 393216), (512, 24576)):
    if c >= 256 and nu != '1':
        continue
    gen = torch.Generator(device='cuda').manual_seed(c + hw)
    f = torch.relu(torch.randn((c, hw), generator=gen, device=dev)) * 0.25
    a = torch.rand((c, c), device=dev) * 1e-3
    d = torch.empty((c, c), device=dev); loss = torch.empty((), device=dev)
    ws = ops.gram_workspace(c, hw, dev)
    t = timeit(lambda: ops.gram_mse_fwd(f, c, hw, 1.0 / (c * hw), a, d, loss, ws, 0), flush, iters=7)
    byt = 4.0 * c * hw + 8.0 * c * c
    print(json.dumps({'cfg': [nu, g], 'C': c, 'HW': hw, 'fwd_ms': round(t, 4), 'GBps': round(byt / t / 1e6, 1),
                      'TFLOPs': round(2.0 * c * c * hw / t / 1e9, 1), 'dsum': float(d.double().sum()),
                      'loss': float(loss)}), flush=True)
