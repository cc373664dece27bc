#!/usr/bin/env python
"""bench.py — headline benchmark of the pyramid Gatys-loss hot path (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host cores

Metric: optimisation steps/s of the L=3 four-level pyramid job (2048x3072 top level; levels 1024x1536,
512x768, 256x384), random-init VGG19 (seed 1234), synthetic 2:3 content/style images, structured-noise init.
A "step" is one reference iteration = one closure (all four levels forward + backward + image gradient) plus
the Adam update that consumes it (neural_style_transfer.py:152-208 of the reference; `iters_num` counts exactly
these).  N GPUs shard every level into N row bands (artstyletransfer_b200/parallel.py) -> strong scaling.

  value : steps/s with everything resident in HBM, timed with CUDA events over K steps (max over ranks).
  e2e   : the same job driven through the public NeuralStyleTransfer.process() async generator, which yields
          the current image as a host numpy array after every optimizer step (device->host copy inside the
          timed region, like the reference's :207-208).  The job's inputs (images) are uploaded once at setup;
          a step has no per-step host input.
  roofline      : the dominant kernel of THIS library inside the timed region (CUDA events around its launches).
  cpu_baseline  : the oracle's restatement of the reference closure on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'pyramid_L3_optim_steps_per_s'
UNIT = 'steps/s'
LEVELS = 4
WEIGHTS = (1e3, 4e5, 1e2)
SRC_HW = (300, 450)           # synthetic source images; resize() maps them to 256*2^l x 384*2^l
L2_BYTES = 126e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--levels', type=int, default=LEVELS, help='pyramid levels (4 = the headline L=3 job)')
    ap.add_argument('--optimizer', default='adam', choices=['adam', 'lbfgs'])
    ap.add_argument('--precision', default=None, choices=[None, 'tf32', 'fp32'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cudnn-autotune', action='store_true',
                    help='leave cuDNN on its heuristics (default: time its engines once per convolution shape)')
    ap.add_argument('--profile', default=None, help='write a torch.profiler kernel table of 3 timed-mode steps (rank 0) to this path')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': float(d['hbm_gbs']), 'bf16_tflops': float(d['bf16_tflops']), 'source': 'MEASURED_PEAKS.json'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'source': 'fallback (B200_PROFILING.md)'}


# TF32 dense peak: not in MEASURED_PEAKS.json; cuBLAS TF32 8192^3 measured on this pool (profiles/r01_peaks.json)
TF32_TFLOPS_MEASURED = 744.9


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.idx}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        t0, t1 = getattr(self, 't0', 0.0), getattr(self, 't1', 1e30)
        inside = [ln for (t, ln) in self.lines if t0 <= t <= t1 + 0.25]
        for ln in (inside or [ln for _, ln in self.lines[-3:]]):
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def seeded_vgg_patch():
    import torch
    import torchvision
    from artstyletransfer_b200 import neural_nets
    real = torchvision.models.vgg19

    def seeded(pretrained=False, progress=False, **kw):
        torch.manual_seed(1234)
        return real(weights=None)

    neural_nets.models.vgg19 = seeded
    return real


def synthetic_sources():
    import numpy as np
    rng = np.random.default_rng(0)
    h, w = SRC_HW
    lo = rng.uniform(0, 1, size=(h // 10, w // 10, 3)).astype(np.float32)
    ys = (np.arange(h) * (lo.shape[0] - 1) / (h - 1)); xs = (np.arange(w) * (lo.shape[1] - 1) / (w - 1))
    y0 = np.floor(ys).astype(int).clip(0, lo.shape[0] - 2); x0 = np.floor(xs).astype(int).clip(0, lo.shape[1] - 2)
    fy = (ys - y0)[:, None, None]; fx = (xs - x0)[None, :, None]
    content = ((lo[y0][:, x0] * (1 - fy) + lo[y0 + 1][:, x0] * fy) * (1 - fx) +
               (lo[y0][:, x0 + 1] * (1 - fy) + lo[y0 + 1][:, x0 + 1] * fy) * fx).astype(np.float32)
    style = rng.uniform(0, 1, size=(h, w, 3)).astype(np.float32)
    return np.clip(content, 0, 1), style


# ---- algorithmic work per kernel ------------------------------------------------------
def kernel_work(key):
    """(bytes, flops) one launch must move / compute, from SURVEY §8(d)."""
    k = key[0]
    if k == 'gram_fwd':
        _, c, hw, _ = key
        return 4.0 * c * hw + 8.0 * c * c, 2.0 * c * c * hw
    if k == 'gram_bwd':
        _, c, hw, _ = key
        return 8.0 * c * hw + 4.0 * c * c, 2.0 * c * c * hw
    if k == 'mse_fwd':
        return 8.0 * key[1], 0.0
    if k == 'mse_bwd':
        return (16.0 if (len(key) > 2 and key[2]) else 12.0) * key[1], 0.0   # X, T (, dX) read; dX written
    if k == 'tv_fwd':
        return 4.0 * key[1], 0.0
    if k == 'tv_bwd':
        return 8.0 * key[1], 0.0
    if k in ('down2x', 'down2x_adj'):
        _, c, h, w = key
        return 4.0 * c * (h * w + h * w / 4), 0.0
    if k == 'gram_fwd_nhwc':
        _, c, hw = key
        return 4.0 * c * hw + 8.0 * c * c, 2.0 * c * c * hw
    if k == 'gram_bwd_nhwc':
        _, c, hw, mode = key                 # mode bit 0: accumulate (reads dF too); bit 1: fused ReLU backward
        return (12.0 if (mode & 1) else 8.0) * c * hw + 4.0 * c * c, 2.0 * c * c * hw
    if k == 'bias_relu':
        return 8.0 * key[2], 0.0            # read + write the activation in place
    if k == 'relu_bwd':
        return 12.0 * key[1], 0.0           # read g, read r, write g
    if k == 'maxpool':
        _, c, h, w = key
        return 4.0 * c * (h * w + (h // 2) * (w // 2)), 0.0
    if k == 'maxpool_bwd':
        _, c, h, w = key
        return 4.0 * c * (2 * h * w + (h // 2) * (w // 2)), 0.0
    if k in ('chw_to_hwc', 'hwc_to_chw'):
        return 8.0 * key[1] * key[2], 0.0
    return 0.0, 0.0


def ncu_traffic(kernel: str):
    """DRAM bytes (read + write) of one launch of `kernel` from the committed `ncu --set full` capture of
    tests/tools/ncu_target.py (profiles/r01_ncu_traffic.json; same shapes as the L=3 top level at N=1), or None."""
    p = os.path.join(ROOT, 'profiles', 'r01_ncu_traffic.json')
    try:
        table = json.load(open(p))
    except (OSError, ValueError):
        return None
    if kernel.startswith('mse_bwd/') and kernel.count('/') == 2:
        kernel = kernel.rsplit('/', 1)[0]            # captured in accumulate mode
    if kernel in table:
        return table[kernel]
    if kernel.startswith('gram_bwd_nhwc/'):          # the fused-ReLU modes (2, 3) move the same DRAM bytes as 0, 1
        head, mode = kernel.rsplit('/', 1)
        kernel = f'{head}/{int(mode) & 1}'
    return table.get(kernel)


def kernel_table(summary, pk):
    rows = []
    for key, st in summary.items():
        by, fl = kernel_work(key)
        if by == 0:
            continue
        t = st['ms_avg'] * 1e-3
        tensor_peak = TF32_TFLOPS_MEASURED * 1e12
        hbm_time, tensor_time = by / (pk['hbm_gbs'] * 1e9), (fl / tensor_peak if fl else 0.0)
        bound = 'tensor' if tensor_time > hbm_time else 'hbm'
        row = {'kernel': '/'.join(str(x) for x in key), 'calls': st['calls'], 'ms_avg': round(st['ms_avg'], 4),
               'ms_total': round(st['ms_total'], 3), 'GBps': round(by / t / 1e9, 1),
               'TFLOPs': round(fl / t / 1e12, 1) if fl else None, 'bound': bound,
               'frac_of_bound': round(max(hbm_time, tensor_time) / t, 3)}
        rows.append(row)
    rows.sort(key=lambda r: -r['ms_total'])
    return rows


_T0 = time.perf_counter()


def phase(msg):
    """Progress line on stderr (rank 0): where a bench run's wall time goes."""
    if int(os.environ.get('RANK', '0')) == 0:
        print(f'[bench +{time.perf_counter() - _T0:7.1f}s] {msg}', file=sys.stderr, flush=True)


def build_job(args, dev):
    """Images, init and the _Job (leaf image + optimizer + per-level LossBuilders), as process() builds them."""
    import numpy as np
    import torch
    from artstyletransfer_b200 import neural_style_transfer as nst
    content_src, style_src = synthetic_sources()
    content_levels, style_levels = [], []
    for lv in range(args.levels):
        content_levels.insert(0, asyncio.run(nst.resize(content_src, lv)))
        style_levels.insert(0, asyncio.run(nst.resize(style_src, lv)))
    pair = nst.ContentStylePair(('synthetic-content', content_src), ('synthetic-style', style_src))
    np.random.seed(0)
    t0 = time.perf_counter()
    init, name = nst.build_init_image(pair, content_levels, style_levels, 'content+noise', 0.95, (9, 18, 36, -1, 0),
                                      (0.30, 0.20, 0.10, 0.20, 0.20), (0.20, 0.30, 0.40, 0.10, 0.00),
                                      (0.20, 0.30, 0.40, 0.60, 0.30), dev)
    torch.cuda.synchronize()
    init_s = time.perf_counter() - t0
    return content_levels, style_levels, init, name, init_s


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    from artstyletransfer_b200 import _lib, neural_style_transfer as nst, ops, parallel
    assert _lib.load().ast_device_check() == 0, _lib.last_error()
    if args.precision:
        nst.PRECISION = args.precision
    from artstyletransfer_b200 import feature_path
    feature_path.CUDNN_BENCHMARK = not args.no_cudnn_autotune
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        parallel.init_sharding()
    seeded_vgg_patch()
    phase('library loaded')
    content_levels, style_levels, init, name, init_s = build_job(args, dev)
    H, W = init.shape[0], init.shape[1]
    phase('images + init built')
    job = nst._Job(dev, 'vgg19', style_levels, args.optimizer, content_levels, init, 10.0, *WEIGHTS, name)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    phase('job set up (targets, band plan)')
    # ---- device-resident timing ------------------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        job.optimizer_step()          # the first closures run eagerly, then the closure is captured as a CUDA graph
    barrier()
    phase('warm-up done (eager closures, cuDNN engine search, graph capture)')
    closures0 = job.step
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        job.optimizer_step()
    e1.record()
    barrier()
    clocks.window(w0, time.perf_counter())
    ms = e0.elapsed_time(e1)
    phase('timed steps done')
    closures = job.step - closures0
    graphed = job._graph is not None
    if args.profile:
        # device-side view of the graphed step (CUPTI sees the kernels a graph replay launches); not a timing source
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                job.optimizer_step()
            torch.cuda.synchronize()
        if rank == 0:
            with open(args.profile, 'w') as f:
                f.write(prof.key_averages().table(sort_by='cuda_time_total', row_limit=60, max_name_column_width=90))
        barrier()
    # ---- per-kernel pass: the same closure launched eagerly with CUDA events around every C-ABI call ---------
    ops.STATS.reset(enabled=True, timing=False)
    job.optimizer_step()              # first eager closure after the graph: allocator warm-up, not measured
    barrier()
    ops.STATS.reset(enabled=True, timing=True)
    probe_steps = 3
    c0 = job.step
    for _ in range(probe_steps):
        job.optimizer_step()
    barrier()
    probe_closures = job.step - c0
    phase('per-kernel eager pass done')
    launches = int(round(ops.STATS.launches / max(probe_closures, 1) * closures))
    summary = ops.STATS.summary()
    for v in summary.values():
        v['calls_per_closure'] = v['calls'] / max(probe_closures, 1)
    ops.STATS.reset(enabled=False)
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = closures / (ms * 1e-3)
    loss_now = float(job.closure().item()) if rank == 0 and world == 1 else None
    mem_gb = torch.cuda.max_memory_allocated(dev) / 1e9

    # ---- end to end through the public API ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        del job
        torch.cuda.empty_cache()

        first_yield_s = [None]

        async def drive():
            drv = nst.NeuralStyleTransfer(dev, 'vgg19', style_levels, args.optimizer)
            n, t_start, last = 0, None, None
            t_call = time.perf_counter()
            total = args.warmup + args.steps
            iters = total if args.optimizer == 'adam' else 2 * total
            async for img, step in drv.process(content_levels, init, 10.0, iters, *WEIGHTS, name):
                n += 1
                last = (img, step)
                if n == 1:
                    first_yield_s[0] = time.perf_counter() - t_call
                if n == args.warmup:
                    barrier()
                    t_start = (time.perf_counter(), step)
            barrier()
            return t_start, time.perf_counter(), last

        (t0, step0), t1, (img, step1) = asyncio.run(drive())
        phase('end-to-end pass through process() done')
        wall = t1 - t0
        if world > 1:
            t = torch.tensor([wall], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        assert img.shape == (H, W, 3) and img.dtype == np.float32
        e2e = {'value': (step1 - step0) / wall, 'unit': UNIT, 'h2d_bytes_per_step': 0,
               'd2h_bytes_per_step': int(img.nbytes), 'timed': 'wall clock around K optimizer steps incl. the per-step '
               'image yield (device->host); job inputs uploaded once at setup',
               'setup_h2d_bytes': int(sum(a.nbytes for a in content_levels + style_levels) + init.nbytes),
               'setup_plus_first_step_s': round(first_yield_s[0], 3) if first_yield_s[0] else None}

    if world > 1:
        # CUDA graphs hold captured NCCL kernels: release them before the communicator goes away
        try:
            del job
        except NameError:
            pass
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
    if rank != 0:
        sys.stdout.flush()
        os._exit(0)               # skip interpreter / NCCL teardown (it can wait forever on captured collectives)
    pk = peaks()
    table = kernel_table(summary, pk)
    top = table[0] if table else None
    top_key = next((k for k in summary if '/'.join(str(x) for x in k) == top['kernel']), None) if top else None
    roofline = None
    if top:
        bound = top['bound']
        achieved = top['GBps'] if bound == 'hbm' else top['TFLOPs']
        peak = pk['hbm_gbs'] if bound == 'hbm' else TF32_TFLOPS_MEASURED
        roofline = {'kernel': top['kernel'], 'bound': bound, 'achieved': achieved, 'peak': peak,
                    'unit': 'GB/s' if bound == 'hbm' else 'TFLOP/s', 'frac': round(achieved / peak, 3),
                    'traffic': ncu_traffic(top['kernel']), 'algorithmic_bytes': kernel_work(top_key)[0],
                    'peak_source': pk['source'] if bound == 'hbm' else 'cuBLAS TF32 8192^3 measured on this pool '
                    '(profiles/r01_peaks.json)', 'ms_avg': top['ms_avg'], 'calls_in_probe_pass': top['calls']}
    ours_ms = sum(r['ms_total'] for r in table)
    # everything else that was bracketed in the eager pass: cuDNN convolutions (torch ops) and collectives
    other = {}
    for key, st in summary.items():
        if kernel_work(key)[0] == 0:
            o = other.setdefault(str(key[0]), {'calls_per_closure': 0.0, 'ms_per_closure': 0.0})
            o['calls_per_closure'] += st['calls'] / max(probe_closures, 1)
            o['ms_per_closure'] += st['ms_total'] / max(probe_closures, 1)
    for o in other.values():
        o['calls_per_closure'] = round(o['calls_per_closure'], 1)
        o['ms_per_closure'] = round(o['ms_per_closure'], 3)
    line = {
        'metric': METRIC, 'value': round(value, 4), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': round(ms / max(closures, 1), 3), 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'tf32' if (args.precision or ops.DEFAULT_PRECISION) == 'tf32' else 'f32',
        'data': 'synthetic',
        'config': {'workload': f'L={args.levels - 1} {args.levels}-level pyramid {H}x{W}, random-init VGG19 seed 1234, '
                               f'{args.optimizer}, structured-noise init (BASELINE configs[3])',
                   'levels': args.levels, 'image': [H, W], 'optimizer': args.optimizer,
                   'parallelism': f'rowband{world}' if world > 1 else 'single',
                   'bands': parallel.PLAN.describe() if parallel.PLAN is not None else None,
                   'cache': f'working set {mem_gb:.1f} GB per step >> {L2_BYTES / 1e6:.0f} MB L2 (no flush needed)',
                   'closures_timed': closures, 'cuda_graph': graphed, 'cudnn_autotune': not args.no_cudnn_autotune, 'gram_operands': args.precision or ops.DEFAULT_PRECISION,
                   'vgg_convs': 'cuDNN via torch ops on channels_last tensors (out of scope)',
                   'kernel_table': 'separate eager pass of 3 steps, CUDA events around every launch of this library'},
        'clocks': clk, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline,
        'kernels': table[:20], 'own_kernels_ms_per_step': round(ours_ms / max(probe_closures, 1), 3),
        'other_bracketed_ms_per_step': other,
        'init_image_s': round(init_s, 4), 'loss_after': loss_now,
    }
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(steps=2, warmup=1, levels=args.levels)
        phase('cpu baseline done')
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---- CPU baseline: the oracle's restatement of the reference closure on the host cores ------------------------
def cpu_closure_setup(sample_levels):
    import numpy as np
    import torch
    from oracle import gatys_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    content_src, style_src = synthetic_sources()
    net, cidx, sidx = O.make_vgg19(1234)
    c_lv = [O.resize_to_level(content_src, lv) for lv in reversed(range(sample_levels))]
    s_lv = [O.resize_to_level(style_src, lv) for lv in reversed(range(sample_levels))]
    targets = [O.torch_targets(net, cidx, sidx, torch.from_numpy(O.prepare_img(c)), torch.from_numpy(O.prepare_img(s)))
               for c, s in zip(c_lv, s_lv)]
    img = torch.from_numpy(O.prepare_img(c_lv[0])).requires_grad_(True)
    opt = torch.optim.Adam((img,), lr=10.0)

    def step():
        for g in opt.param_groups:
            g['lr'] *= 0.999
        opt.zero_grad()
        _, _, grad = O.torch_closure(net, cidx, sidx, targets, img, WEIGHTS)
        img.grad = grad
        opt.step()
    return step


def pixel_ratio(levels_full, levels_sample):
    full = sum(4.0 ** l for l in range(levels_full))
    samp = sum(4.0 ** l for l in range(levels_sample))
    return full / samp


def cpu_baseline(steps, warmup, levels, sample_levels=2):
    sample_levels = min(sample_levels, levels)
    step = cpu_closure_setup(sample_levels)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    ratio = pixel_ratio(levels, sample_levels)
    return {'value': round(1.0 / (dt * ratio), 5), 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
            'sample': f'{steps} closures of the {sample_levels}-level pyramid (top {256 * 2 ** (sample_levels - 1)}x'
                      f'{384 * 2 ** (sample_levels - 1)}) at {dt:.3f} s each on {os.cpu_count()} threads, scaled to the '
                      f'{levels}-level job by the pixel ratio {ratio:.2f} (VGG conv cost is linear in pixels)',
            'sample_s_per_step': round(dt, 4)}


def run_reference(args):
    """The reference's own algorithm for this path (oracle port: torch-CPU closure restating
    neural_style_transfer.py:84-112, :152-193 + the Adam update) on the host cores; /root/reference itself does
    not exist on the GPU box.  Each step is a bounded sample (2-level pyramid), scaled by pixel count."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample_levels = min(2, args.levels)
    step = cpu_closure_setup(sample_levels)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    ratio = pixel_ratio(args.levels, sample_levels)
    value = 1.0 / (dt * ratio)
    sample = (f'{args.steps} closures of the {sample_levels}-level pyramid at {dt:.3f} s each on {os.cpu_count()} '
              f'threads, scaled to the {args.levels}-level job by the pixel ratio {ratio:.2f}')
    line = {'impl': 'reference', 'metric': METRIC, 'value': round(value, 5), 'unit': UNIT,
            'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': round(dt * ratio * 1e3, 1), 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'L={args.levels - 1} {args.levels}-level pyramid 2048x3072, random-init VGG19 seed 1234, '
                                   'adam (CPU, oracle port of the reference closure)', 'levels': args.levels,
                       'optimizer': 'adam'},
            'cpu_baseline': {'value': round(value, 5), 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                             'sample': sample},
            'e2e': {'value': round(value, 5), 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
        if int(os.environ.get('WORLD_SIZE', '1')) > 1:
            sys.stdout.flush()
            os._exit(0)


if __name__ == '__main__':
    main()
