"""GPU diagnostic: WHERE does the row-band sharded closure differ from the unsharded one?

For one (world, bands, n_levels) case of tests/test_gpu_sharding.py it compares, level by level and BEFORE the bicubic
adjoint chain mixes the levels, the image-gradient of every pyramid level summed over the emulated ranks with the
unsharded gradient of that level, and prints the worst rows next to the band edges.  Usage:
  python tests/tools/shard_error_probe.py WORLD pyramid|uniform N_LEVELS [exact]"""
import os
import sys
import threading

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import gatys_oracle as O  # noqa: E402
from test_gpu_sharding import ThreadGroup, WEIGHTS, dev  # noqa: E402


def main():
    world, bands, n_levels = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
    exact = len(sys.argv) > 4 and sys.argv[4] == 'exact'
    weights = (1e3, 0.0, 1e2) if exact else WEIGHTS
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    import torchvision
    from artstyletransfer_b200 import feature_path, math_utils, neural_nets, neural_style_transfer as nst, ops
    from artstyletransfer_b200.parallel import PyramidBands
    from artstyletransfer_b200 import sharded_path as sp
    feature_path.CUDNN_BENCHMARK = False
    real = torchvision.models.vgg19
    neural_nets.models.vgg19 = lambda pretrained=False, progress=False, **kw: (torch.manual_seed(1234), real(weights=None))[1]
    H, W = (256, 96) if n_levels == 2 else (256, 128)
    content, style = O.synthetic_images(H, W, seed=11)
    init = np.clip(content * 0.5 + np.random.default_rng(12).uniform(0, 1, size=content.shape) * 0.5, 0, 1).astype(np.float32)
    net, cidx, sidx = math_utils.prepare_model('vgg19', dev())
    if exact:
        cidx = 5
    c_img = [nst.prepare_img(content[::1 << i, ::1 << i].copy(), dev()) for i in range(n_levels)]
    s_img = [nst.prepare_img(style[::1 << i, ::1 << i].copy(), dev()) for i in range(n_levels)]
    lbs = [nst.LossBuilder(cidx, sidx, c, s_, net, *weights) for c, s_ in zip(c_img, s_img)]
    img = nst.prepare_img(init, dev())
    lv_imgs = [img]
    for i in range(1, n_levels):
        lv_imgs.append(ops.bicubic_down_raw(lv_imgs[-1], lv_imgs[-1].shape[-2] // 2, lv_imgs[-1].shape[-1] // 2))
    ref = []
    for i in range(n_levels):                       # unsharded gradient of every level w.r.t. ITS OWN image
        x = lv_imgs[i].clone().requires_grad_(True)
        t = lbs[i].build(x)[0]
        t.backward()
        ref.append((t.item(), x.grad.clone()))
    plan = lbs[0].path_plan(img)
    sizes = [(H >> i, W >> i) for i in range(n_levels)]
    pb = PyramidBands(sizes, world, uniform=bands == 'uniform')
    print('plan:', pb.describe())
    shared = {'bufs': [None] * world, 'barrier': threading.Barrier(world, timeout=60), 'mail': {}}
    results, errors = [None] * world, []

    def run(rank):
        try:
            torch.cuda.set_device(dev())
            grp = ThreadGroup(rank, world, shared)
            levels = [sp.ShardedPathLevel(grp, plan, c_img[i], s_img[i], cidx, sidx, weights, *sizes[i],
                                          band=(*pb.band(i, rank), *pb.neighbours(i, rank))) for i in range(n_levels)]
            lanes = sp.Lanes(dev(), n_levels) if os.environ.get('PROBE_LANES', '1') == '1' else sp._SERIAL
            with torch.no_grad():
                out4s, state = sp.pyramid_forward(levels, [t.clone() for t in lv_imgs], lanes)
                d_imgs = sp.pyramid_backward(state, None, lanes)
            results[rank] = ([o[0].item() for o in out4s], [d.clone() for d in d_imgs])
        except Exception:
            import traceback
            errors.append(traceback.format_exc())
            shared['barrier'].abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errors:
        print(errors[0])
        sys.exit(1)
    for li in range(n_levels):
        g = sum(results[r][1][li] for r in range(world))
        gr = ref[li][1]
        diff = (g - gr).double()
        rel = float(diff.norm() / gr.double().norm())
        rows = torch.linalg.norm(diff.permute(2, 0, 1, 3).reshape(diff.shape[2], -1), dim=1)
        scale = float(torch.linalg.norm(gr.double().permute(2, 0, 1, 3).reshape(gr.shape[2], -1), dim=1).pow(2).mean().sqrt())
        prof = (rows / scale).cpu().numpy()
        worst = np.argsort(-prof)[:8]
        print(f'level {li} {sizes[li]}: loss sharded {results[0][0][li]:.6e} vs {ref[li][0]:.6e}; grad rel err {rel:.3e}; '
              f'bands {pb.bounds[li]}; median row err {np.median(prof):.2e}; worst rows '
              + ', '.join(f'{int(r)}:{prof[r]:.1e}' for r in sorted(worst)))
        cols = torch.linalg.norm(diff.permute(3, 0, 1, 2).reshape(diff.shape[3], -1), dim=1)
        cprof = (cols / scale).cpu().numpy()
        cw = np.argsort(-cprof)[:6]
        print('        worst cols ' + ', '.join(f'{int(c)}:{cprof[c]:.1e}' for c in sorted(cw)))


if __name__ == '__main__':
    main()
