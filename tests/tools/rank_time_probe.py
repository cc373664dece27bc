"""GPU diagnostic: how long does EACH RANK's share of a row-band sharded closure take when nobody waits for anybody?

On ONE GPU, for every rank r of a `world`-rank plan (parallel.PyramidBands), the product's lock-step closure
(sharded_path.PyramidFn forward + backward: bicubic chain, every band of every level this rank owns, partial Grams,
finalize, backward, adjoint chain) is captured in a CUDA graph with the communication stubbed out (halo rows are not
swapped, nothing is reduced: the numbers are wrong, the work is the same) and the replay is timed.  A lock-step
closure is as slow as its slowest rank: the table shows which rank that is and how well the plan's cost model
(rows x width + a per-band overhead) matches the hardware.  Usage:
  python tests/tools/rank_time_probe.py [--world 8] [--overheads 0.015,0.03,0.05] [--out gpurun_out/rank_times.jsonl]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gatys_oracle as O  # noqa: E402

WEIGHTS = (1e3, 4e5, 1e2)


class NullGroup:
    """A rank of a `world`-rank job whose neighbours never answer: every collective is a no-op."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def all_reduce_sum(self, t):
        pass

    def exchange(self, sends, recvs):
        pass


def time_rank(pyr, x, reps):
    from artstyletransfer_b200.sharded_path import PyramidFn

    class Ctx:
        needs_input_grad = (False, True)

    def closure():
        ctx = Ctx()
        with torch.no_grad():
            PyramidFn.forward(ctx, pyr, x)
            PyramidFn.backward(ctx, None)

    for _ in range(2):
        closure()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode='thread_local'):
        closure()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--world', type=int, default=8)
    ap.add_argument('--levels', type=int, default=4)
    ap.add_argument('--height', type=int, default=2048)
    ap.add_argument('--width', type=int, default=3072)
    ap.add_argument('--overheads', default='0.015,0.03,0.05,0.08')
    ap.add_argument('--bounds', default=None, help='JSON list of per-level row edges to time instead of planned ones')
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--sweep', action='store_true',
                    help='time ONE band of r rows of ONE level (an interior band with both neighbours), for every level '
                         'and r = 32, 32 + step, ...: the cost table a plan can be checked against')
    ap.add_argument('--sweep-max-rows', type=int, default=512, help='in rows of the TOP level equivalent cost')
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    import torchvision
    from artstyletransfer_b200 import math_utils, neural_nets, neural_style_transfer as nst
    from artstyletransfer_b200.parallel import PyramidBands
    from artstyletransfer_b200.sharded_path import ShardedPathLevel, ShardedPyramid
    real = torchvision.models.vgg19
    neural_nets.models.vgg19 = lambda pretrained=False, progress=False, **kw: (torch.manual_seed(1234), real(weights=None))[1]
    H, W, L = args.height, args.width, args.levels
    content, style = O.synthetic_images(H, W, seed=11)
    init = np.clip(content * 0.5 + np.random.default_rng(12).uniform(0, 1, size=content.shape) * 0.5, 0, 1).astype(np.float32)
    net, cidx, sidx = math_utils.prepare_model('vgg19', dev)
    c_img = [nst.prepare_img(content[::1 << i, ::1 << i].copy(), dev) for i in range(L)]
    s_img = [nst.prepare_img(style[::1 << i, ::1 << i].copy(), dev) for i in range(L)]
    lb = nst.LossBuilder(cidx, sidx, c_img[0], s_img[0], net, *WEIGHTS)
    x = nst.prepare_img(init, dev)
    plan = lb.path_plan(x)
    sizes = [(H >> i, W >> i) for i in range(L)]
    plans = []
    if args.sweep:
        out = open(args.out, 'a') if args.out else None
        for li, (h, w) in enumerate(sizes):
            step = 16 if li == 0 else 32
            top = min(h, args.sweep_max_rows << (2 * li))
            for rows in [0] + list(range(32, top + 1, step)) if li == 0 else range(32, top + 1, step):
                pb = PyramidBands(sizes, 3)
                a = min(max((h - rows) // 2 // 16 * 16, 0), h - rows)
                pb.bounds = [[0, a, a + rows, hh] if lj == li else [0, hh, hh, hh] for lj, (hh, _) in enumerate(sizes)]
                depth = 2
                grp = NullGroup(1, 3)
                levels = [ShardedPathLevel(grp, plan, c_img[i], s_img[i], cidx, sidx, WEIGHTS, *sizes[i],
                                           band=(*pb.band(i, 1), *pb.neighbours(i, 1)), halo_depth=depth)
                          for i in range(L)]
                ms = round(time_rank(ShardedPyramid(levels), x, args.reps), 4)
                rec = {'level': li, 'rows': rows, 'ms': ms}
                print(json.dumps(rec), flush=True)
                if out:
                    out.write(json.dumps(rec) + '\n')
                    out.flush()
                del levels
                torch.cuda.empty_cache()
        return
    if args.bounds:
        pb = PyramidBands(sizes, args.world)
        pb.bounds = json.loads(args.bounds)
        plans.append(('custom', pb))
    else:
        for ov in [float(v) for v in args.overheads.split(',')]:
            plans.append((f'overhead {ov}', PyramidBands(sizes, args.world, overhead=ov)))
    one = PyramidBands(sizes, 1)
    plans.append(('one rank', one))
    out = open(args.out, 'a') if args.out else None
    for tag, pb in plans:
        depth = pb.halo_depth()
        times = []
        for rank in range(pb.world):
            grp = NullGroup(rank, pb.world)
            levels = [ShardedPathLevel(grp, plan, c_img[i], s_img[i], cidx, sidx, WEIGHTS, *sizes[i],
                                       band=(*pb.band(i, rank), *pb.neighbours(i, rank)), halo_depth=depth)
                      for i in range(L)]
            pyr = ShardedPyramid(levels)
            times.append(round(time_rank(pyr, x, args.reps), 4))
            del pyr, levels
            torch.cuda.empty_cache()
        rec = {'plan': tag, 'bands': pb.describe(), 'model_loads': [round(v, 4) for v in pb.loads()],
               'ms_per_rank': times, 'max_ms': max(times), 'sum_ms': round(sum(times), 3)}
        print(json.dumps(rec), flush=True)
        if out:
            out.write(json.dumps(rec) + '\n')
            out.flush()


if __name__ == '__main__':
    main()
