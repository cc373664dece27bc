"""Bring-up probe for the tcgen05 Gram kernels (run on a B200): one (kind, C, HW) per subprocess so that a
trapped kernel cannot poison the next case.  Prints relative errors vs an fp64 torch Gram and a coarse map of
where the error sits (helps to tell a descriptor/swizzle mistake from a pipeline one)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def one(kind, c, hw):
    import torch
    from artstyletransfer_b200 import ops
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(c * 7 + hw)
    f = (torch.relu(torch.randn((c, hw), generator=g)) * 0.25).to(dev)
    if kind == 'fwd':
        out = torch.empty((c, c), device=dev)
        ws = ops.gram_workspace(c, hw, dev)
        ops.gram_mse_fwd(f, c, hw, 1.0 / (c * hw), None, out, None, ws, 0)
        torch.cuda.synchronize()
        ref = (f.double() @ f.double().t()) / (c * hw)
    else:
        d = torch.randn((c, c), generator=g).to(dev)
        d = (d + d.t()).contiguous()
        out = torch.empty_like(f)
        ops.gram_bwd(d, f, c, hw, 1.0, None, out, False, 0)
        torch.cuda.synchronize()
        ref = d.double() @ f.double()
    err = (out.double() - ref)
    relerr = (err.norm() / ref.norm()).item()
    print(f'{kind} C={c} HW={hw}: rel_fro={relerr:.3e} max_abs={err.abs().max().item():.3e} ref_max={ref.abs().max().item():.3e}')
    if relerr > 1e-3:
        rows = min(out.shape[0], 512)
        blk = 32
        e = err[:rows, :min(out.shape[1], 512)].abs()
        r = ref[:rows, :min(out.shape[1], 512)].abs() + 1e-30
        m = (e / r.max()).reshape(rows // blk, blk, -1, blk).amax(dim=(1, 3))
        torch.set_printoptions(linewidth=200, precision=2, sci_mode=True)
        print('blockwise max |err| / max|ref| (32x32 blocks):')
        print(m.cpu())
        print('out[0,:8]', out[0, :8].cpu().tolist())
        print('ref[0,:8]', ref[0, :8].cpu().tolist())


if __name__ == '__main__':
    if len(sys.argv) == 4:
        one(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    cases = [('fwd', 128, 256), ('fwd', 128, 4096), ('fwd', 64, 4096), ('fwd', 256, 2048), ('fwd', 512, 1536),
             ('fwd', 128, 24576), ('fwd', 64, 98304),
             ('bwd', 128, 256), ('bwd', 64, 4096), ('bwd', 128, 24576), ('bwd', 256, 6144), ('bwd', 512, 1536)]
    for kind, c, hw in cases:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), kind, str(c), str(hw)], capture_output=True,
                           text=True, timeout=120)
        sys.stdout.write(r.stdout)
        if r.returncode != 0:
            tail = (r.stderr or '').strip().splitlines()[-3:]
            print(f'{kind} C={c} HW={hw}: FAILED rc={r.returncode} :: ' + ' | '.join(tail))
        sys.stdout.flush()
