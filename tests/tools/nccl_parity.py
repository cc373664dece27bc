"""Real multi-process parity of the row-band sharded closure: N ranks over NCCL (one process per GPU) against the
unsharded closure computed in the same processes.  Launch:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      tests/tools/nccl_parity.py [--hw 256 128] [--levels 3] [--steps 6]

Checks (exit code 1 on failure, one JSON line on rank 0):
  * loss of the sharded closure == unsharded loss to 1e-4 (BASELINE tolerance), identical bits on every rank;
  * all-reduced image gradient == unsharded gradient within the TF32 gradient budget, identical bits on every rank;
  * `steps` Adam steps through _Job.optimizer_step (CUDA graph captured after two eager closures, collectives inside
    the graph): every rank ends with the bit-identical image, PSNR vs the unsharded run >= 60 dB at lr_start = 1.
AST_HALO=peer selects the NVLink peer-memory halo push (csrc/halo.cu) instead of grouped NCCL send/recv."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gatys_oracle as O  # noqa: E402  (test infrastructure: synthetic inputs + PSNR only)

WEIGHTS = (1e3, 4e5, 1e2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--hw', type=int, nargs=2, default=(256, 128))
    ap.add_argument('--levels', type=int, default=3)
    ap.add_argument('--steps', type=int, default=6)
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('LOCAL_RANK', 0), ('WORLD_SIZE', 1)))
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    import torchvision
    from artstyletransfer_b200 import feature_path, neural_nets, neural_style_transfer as nst, parallel
    feature_path.CUDNN_BENCHMARK = False
    nst.GRAPH_STRICT = True
    real = torchvision.models.vgg19
    neural_nets.models.vgg19 = lambda pretrained=False, progress=False, **kw: (torch.manual_seed(1234), real(weights=None))[1]

    H, W = args.hw
    content, style = O.synthetic_images(H, W, seed=11)
    init = np.clip(content * 0.5 + np.random.default_rng(12).uniform(0, 1, size=content.shape) * 0.5, 0, 1).astype(np.float32)
    c_lv = [content[::1 << i, ::1 << i].copy() for i in range(args.levels)]
    s_lv = [style[::1 << i, ::1 << i].copy() for i in range(args.levels)]

    def run(sharded):
        if sharded:
            parallel.init_sharding()
        else:
            parallel.disable_sharding()
        job = nst._Job(dev, 'vgg19', s_lv, 'adam', c_lv, init, 1.0, *WEIGHTS, 'nccl-parity')
        assert bool(job.sharded_levels) == sharded
        job.optimizer.zero_grad()
        total = job._evaluate()
        loss, grad = float(total.item()), job.optimizing_img.grad.clone()
        del total                                 # drop the autograd graph (and the leaf's AccumulateGrad node) before the capture
        job.optimizing_img.grad = None
        for _ in range(args.steps):
            job.optimizer_step()
        torch.cuda.synchronize()
        graphed = job._graph is not None
        img = job.optimizing_img.detach().clone()
        plan = parallel.PLAN.describe() if parallel.PLAN is not None else None
        del job
        return loss, grad, img, graphed, plan

    l0, g0, x0, graphed0, _ = run(False)
    l1, g1, x1, graphed1, plan = run(True)

    def same_on_all_ranks(t):
        mine = t.double().sum().reshape(1)
        bits = torch.stack([mine, t.double().abs().max().reshape(1)]).reshape(-1)
        got = [torch.empty_like(bits) for _ in range(world)]
        dist.all_gather(got, bits)
        return all(torch.equal(g, got[0]) for g in got)

    losses = [None] * world
    dist.all_gather_object(losses, l1)
    res = {
        'world': world, 'halo': os.environ.get('AST_HALO', parallel.DEFAULT_HALO), 'image': [H, W], 'levels': args.levels,
        'bands': plan, 'loss_unsharded': l0, 'loss_sharded': l1,
        'loss_rel_err': abs(l1 - l0) / abs(l0), 'loss_identical_on_ranks': all(v == losses[0] for v in losses),
        'grad_rel_err': float(torch.linalg.norm(g1 - g0) / torch.linalg.norm(g0)),
        'grad_identical_on_ranks': same_on_all_ranks(g1),
        'steps': args.steps, 'graph_unsharded': graphed0, 'graph_sharded': graphed1,
        'image_identical_on_ranks': same_on_all_ranks(x1),
        'psnr_vs_unsharded_dB': O.psnr(O.unprepare_img(x1.cpu().numpy()), O.unprepare_img(x0.cpu().numpy())),
    }
    ok = (res['loss_rel_err'] <= 1e-4 and res['loss_identical_on_ranks'] and res['grad_rel_err'] <= 2e-3
          and res['grad_identical_on_ranks'] and res['image_identical_on_ranks'] and res['graph_sharded']
          and res['psnr_vs_unsharded_dB'] >= 60.0)
    res['ok'] = bool(ok)
    flags = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps(res), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if int(flags.item()) == 1 else 1)     # skip NCCL teardown (captured collectives)


if __name__ == '__main__':
    main()
