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


def metric_name(levels):
    return METRIC if levels == LEVELS else f'pyramid_L{levels - 1}_optim_steps_per_s'
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
    ap.add_argument('--precision', default=None, choices=[None, 'tf32', 'fp32', 'bf16'])
    ap.add_argument('--init', default='structured', choices=['structured', 'pixel'],
                    help="structured: Config() noise levels (9,18,36,-1,0), style-permutation noise (BASELINE configs[2],[3]); "
                         "pixel: lab.py's PIXEL_WIDE_NOISE_CONFIG with clipped normal noise per pixel (BASELINE configs[1])")
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cudnn-autotune', action='store_true',
                    help='leave cuDNN on its heuristics (default: time its engines once per convolution shape)')
    ap.add_argument('--profile', default=None, help='write a torch.profiler kernel table of 3 timed-mode steps (rank 0) to this path')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ncu-closure', action='store_true',
                    help='for `ncu --profile-from-start off`: set the job up, run 2 eager closures, then ONE eager closure + '
                         'optimizer update between cudaProfilerStart/Stop, and exit (no timing, no JSON line)')
    ap.add_argument('--no-library-baseline', action='store_true',
                    help='skip the torch/cuDNN/cuBLAS "library bar" (the reference closure with device=cuda) at N=1')
    ap.add_argument('--cpu-budget-s', type=float, default=480.0,
                    help='--impl reference: wall-clock budget for warmup+steps of the REAL config; above it the '
                         'steps fall back to a bounded sample and the line says so')
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
    if k in ('tv_bwd', 'tv_bwd_rows'):     # (name, elements touched): the rows variant counts its own rows only
        return 8.0 * key[1], 0.0
    if k in ('down2x', 'down2x_adj'):
        _, c, h, w = key
        return 4.0 * c * (h * w + h * w / 4), 0.0
    if k == 'gram_fwd_nhwc':
        _, c, hw = key
        return 4.0 * c * hw + 8.0 * c * c, 2.0 * c * c * hw
    if k in ('gram_bwd_nhwc', 'gram_bwd_nhwc_bf16'):
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
    if k == 'unprepare':
        return 24.0 * key[1], 0.0           # 3 planes read, 3 interleaved channels written
    if k == 'noise_init':
        return 24.0 * key[1] * key[2], 0.0  # K7: content read once, init written once
    if k == 'resize':
        _, c, h, w, oh, ow = key
        return 4.0 * c * (h * w + oh * ow), 0.0
    if k in ('down2x_tv', 'down2x_adj_tv'):
        _, c, h, w = key
        return 4.0 * c * (h * w + h * w / 4), 0.0
    return 0.0, 0.0


def _gap_summary(times, warmup):
    """Wall-clock gaps between consecutive yields of the timed region: median, max and where the max was."""
    gaps = [1e3 * (b - a) for a, b in zip(times[warmup - 1:-1], times[warmup:])]
    if not gaps:
        return None
    srt = sorted(gaps)
    worst = max(range(len(gaps)), key=gaps.__getitem__)
    return {'median': round(srt[len(srt) // 2], 3), 'max': round(gaps[worst], 3), 'max_at_timed_step': worst + 1,
            'over_3x_median': sum(1 for g in gaps if g > 3 * srt[len(srt) // 2])}


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


INIT_ARGS = ('content+noise', 0.95, (9, 18, 36, -1, 0), (0.30, 0.20, 0.10, 0.20, 0.20), (0.20, 0.30, 0.40, 0.10, 0.00),
             (0.20, 0.30, 0.40, 0.60, 0.30))          # config.Config() defaults (config.py:10-18)
# lab.py:26-32 PIXEL_WIDE_NOISE_CONFIG, used with USE_NORMAL_NOISE_JUST_FOR_DEMONSTRATION (neural_style_transfer.py:296-298)
PIXEL_INIT_ARGS = ('content+noise', 0.5, (-1,), (1.0,), (1.0,), (0.5,))


def init_args(args):
    return PIXEL_INIT_ARGS if args.init == 'pixel' else INIT_ARGS


def workload_config(args):
    """The workload, stated identically by both arms (--impl ours / reference): same images, weights, optimizer,
    init and step definition.  Arm-specific facts (sharding, graphs, sampling) live in 'impl_config'."""
    H, W = 256 * 2 ** (args.levels - 1), 384 * 2 ** (args.levels - 1)
    cfg_no = {1: 0, 2: 1, 3: 2, 4: 3}.get(args.levels)
    return {'workload': f'L={args.levels - 1} {args.levels}-level pyramid {H}x{W}, random-init VGG19 seed 1234, '
                        f'{args.optimizer}, structured-noise init (BASELINE configs[{cfg_no}])',
            'levels': args.levels, 'image': [H, W], 'optimizer': args.optimizer,
            'weights': list(WEIGHTS),
            'init': ('content+noise, PIXEL_WIDE_NOISE_CONFIG, clipped normal noise per pixel, np.random.seed(0)' if args.init == 'pixel'
                     else 'content+noise, Config() noise levels (9,18,36,-1,0), np.random.seed(0)'),
            'step': 'one optimizer.step = Adam: 1 closure (all levels fwd+bwd) + update; LBFGS: 2 closures',
            'cache': 'working set of a step (>= 1 GB of activations at every level set) >> 126 MB L2: no flush needed'}


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
    nst.USE_NORMAL_NOISE_JUST_FOR_DEMONSTRATION = args.init == 'pixel'
    init, name = nst.build_init_image(pair, content_levels, style_levels, *init_args(args), dev)
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
    if args.warmup < 3:
        # the contract asks for W >= 3; with fewer the closure's CUDA-graph capture (closure GRAPH_WARMUP + 1 = 3: a
        # once-per-job 0.3 s) would land inside the timed region.  The line reports the warm-up that was run.
        if rank == 0:
            print(f'[bench] --warmup {args.warmup} raised to 3 (graph capture happens in closure 3)', file=sys.stderr)
        args.warmup = 3
    if args.precision:
        nst.PRECISION = args.precision
    from artstyletransfer_b200 import feature_path
    if args.no_cudnn_autotune:        # product default: on (feature_path.CUDNN_BENCHMARK); off for runs under ncu
        feature_path.CUDNN_BENCHMARK = False
    nst.GRAPH_STRICT = True           # a failed capture must fail the bench, not show up as a slower number
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        parallel.init_sharding()
    seeded_vgg_patch()
    phase('library loaded')
    build_job(args, dev)              # untimed first pass: allocator / module warm-up of the set-up path
    ops.STATS.reset(enabled=True, timing=True)
    content_levels, style_levels, init, name, init_s = build_job(args, dev)
    torch.cuda.synchronize()
    setup_kernels = ops.STATS.summary()
    ops.STATS.reset(enabled=False)
    H, W = init.shape[0], init.shape[1]
    phase('images + init built')
    job = nst._Job(dev, 'vgg19', style_levels, args.optimizer, content_levels, init, 10.0, *WEIGHTS, name)
    # parity across N: the first closure sees the same image on every rank count, so its loss and gradient must agree
    parity = first_closure_parity(job, rank, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    phase('job set up (targets, band plan)')
    if args.ncu_closure:
        nst.GRAPH_CLOSURE = False
        for _ in range(2):
            job.optimizer_step()
        barrier()
        torch.cuda.profiler.start()
        job.optimizer_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        phase('one eager closure profiled')
        return
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
    # per-rank busy time of the eager pass (ms per closure): where a sharded step's critical path and imbalance sit
    busy = {'cudnn': 0.0, 'own': 0.0, 'halo': 0.0, 'allreduce': 0.0}
    for key, st in summary.items():
        name = str(key[0])
        per = st['ms_total'] / max(probe_closures, 1)
        if name.startswith('cudnn'):
            busy['cudnn'] += per
        elif name.startswith('halo_exchange'):
            busy['halo'] += per
        elif name.startswith('allreduce'):
            busy['allreduce'] += per
        elif kernel_work(key)[0] > 0:
            busy['own'] += per
    busy = {k: round(v, 3) for k, v in busy.items()}
    per_rank_busy = [busy]
    if world > 1:
        per_rank_busy = [None] * world
        dist.all_gather_object(per_rank_busy, busy)
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = closures / (ms * 1e-3)
    loss_closures = job.step                        # warm-up + timed + per-kernel pass (+ 3 with --profile)
    loss_now = float(job.closure().item())          # collective under sharding: every rank evaluates
    mem_gb = torch.cuda.max_memory_allocated(dev) / 1e9

    # ---- end to end through the public API ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        del job
        torch.cuda.empty_cache()

        first_yield_s = [None]
        yield_times = []
        host_allocs = [None, None]        # page-locked allocations (cudaHostAlloc) at the start / end of the timed region

        def _host_allocs():
            try:
                st = torch.cuda.host_memory_stats()
                return int(st.get('num_host_alloc', st.get('host_alloc_count', -1)))
            except Exception:
                return -1

        import gc
        gc_log, gc_t0 = [], [0.0]       # (generation, pause ms, wall time) of every collection during the pass

        def gc_cb(phase, info):
            if phase == 'start':
                gc_t0[0] = time.perf_counter()
            else:
                now = time.perf_counter()
                gc_log.append((info.get('generation'), 1e3 * (now - gc_t0[0]), now))

        gc.callbacks.append(gc_cb)

        async def drive():
            drv = nst.NeuralStyleTransfer(dev, 'vgg19', style_levels, args.optimizer)
            n, t_start, last = 0, None, None
            t_call = time.perf_counter()
            total = args.warmup + args.steps
            iters = total if args.optimizer == 'adam' else 2 * total
            async for img, step in drv.process(content_levels, init, 10.0, iters, *WEIGHTS, name):
                n += 1
                last = (img, step)
                yield_times.append(time.perf_counter())
                if n == 1:
                    first_yield_s[0] = time.perf_counter() - t_call
                if n == args.warmup:
                    barrier()
                    host_allocs[0] = _host_allocs()
                    t_start = (time.perf_counter(), step)
            # the clock stops at the LAST YIELD: what follows is the generator's tear-down (CUDA graph, NCCL-capturing
            # buffers, executor), a once-per-job cost of 0.05-1 s that is not part of a step
            t_end = yield_times[-1]
            barrier()
            host_allocs[1] = _host_allocs()
            return t_start, t_end, last

        (t0, step0), t1, (img, step1) = asyncio.run(drive())
        gc.callbacks.remove(gc_cb)
        gc_timed = [(g, round(ms, 2)) for g, ms, t in gc_log if t0 <= t <= t1 and ms >= 1.0]
        phase('end-to-end pass through process() done')
        wall = t1 - t0
        if world > 1:
            t = torch.tensor([wall], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        if rank == 0:
            assert img.shape == (H, W, 3) and img.dtype == np.float32
        e2e = {'value': (step1 - step0) / wall, 'unit': UNIT, 'h2d_bytes_per_step': 0,
               'd2h_bytes_per_step': int(H * W * 3 * 4), 'timed': 'wall clock from yield W to yield W+K of '
               'NeuralStyleTransfer.process(), max over ranks, incl. the per-step image yield (device->host into page-locked memory, '
               'overlapped with the next step; rank 0 copies under sharding); job inputs uploaded once at setup',
               'setup_h2d_bytes': int(sum(a.nbytes for a in content_levels + style_levels) + init.nbytes),
               'setup_plus_first_step_s': round(first_yield_s[0], 3) if first_yield_s[0] else None,
               'yield_gaps_ms': _gap_summary(yield_times, args.warmup),
               'python_gc_pauses_over_1ms_in_timed_region': gc_timed,
               'host_trace_over_5ms': ([(w, st, round(ms, 2)) for w, st, ms in nst.YIELD_TRACE if ms > 5.0]
                                       if nst.YIELD_TRACE is not None else None),
               'page_locked_allocs_in_timed_region': (host_allocs[1] - host_allocs[0]
                                                      if None not in host_allocs and min(host_allocs) >= 0 else None)}

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
        'metric': metric_name(args.levels), 'value': round(value, 4), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': round(ms / max(closures, 1), 3), 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None,
        'dtype': {'tf32': 'tf32', 'fp32': 'f32', 'bf16': 'tf32 (bf16 operands in the 512-channel backward)'}[args.precision or ops.DEFAULT_PRECISION],
        'data': 'synthetic',
        'config': workload_config(args),
        'impl_config': {'parallelism': f'rowband{world}' if world > 1 else 'single',
                        'bands': parallel.PLAN.describe() if parallel.PLAN is not None else None,
                        'peak_memory_gb': round(mem_gb, 1),
                        'closures_timed': closures, 'cuda_graph': graphed,
                        'cudnn_autotune': feature_path.CUDNN_BENCHMARK, 'gram_operands': args.precision or ops.DEFAULT_PRECISION,
                        'halo': os.environ.get('AST_HALO', parallel.DEFAULT_HALO) if world > 1 else None,
                        'vgg_convs': 'cuDNN via torch ops on channels_last tensors (out of scope)',
                        'kernel_table': 'separate eager pass of 3 steps, CUDA events around every launch of this library'},
        'clocks': clk, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline,
        'kernels': table[:20], 'own_kernels_ms_per_step': round(ours_ms / max(probe_closures, 1), 3),
        'other_bracketed_ms_per_step': other,
        'per_rank_busy_ms_eager_pass': per_rank_busy,
        'setup_kernels': kernel_table(setup_kernels, pk)[:8],
        'init_image_s': round(init_s, 4), 'loss_after': loss_now, 'loss_after_closures': loss_closures, 'parity': parity,
        # `value` counts reference iterations (closures: what iters_num counts); LBFGS takes two per optimizer.step
        'closures_per_s': round(value, 4),
        'optimizer_steps_per_s': round(value / (2 if args.optimizer == 'lbfgs' else 1), 4),
    }
    if world == 1 and not args.no_library_baseline:
        line['library_baseline'] = library_baseline(args, dev, content_levels, style_levels, init)
        phase('library baseline (reference torch path on this GPU) done')
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(args)
        phase('cpu baseline done')
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def first_closure_parity(job, rank, dev):
    """Loss and image gradient of the FIRST closure (the init image: identical whatever the rank count), evaluated
    eagerly on every rank (collective under sharding) and summarised so that runs at different N can be compared:
    total loss, gradient L2 norm and its projection on a fixed pseudo-random direction."""
    import torch
    job.optimizer.zero_grad()
    total = job._evaluate()
    g = job.optimizing_img.grad
    gen = torch.Generator(device=dev).manual_seed(2024)
    direction = torch.randn(g.shape, generator=gen, device=dev)
    out = {'loss_first': float(total.item()), 'grad_first_l2': float(g.double().norm().item()),
           'grad_first_proj': float((g.double() * direction.double()).sum().item() / direction.double().norm().item())}
    job.optimizing_img.grad = None
    return out if rank == 0 else None


# ---- the library bar: the reference's torch path on THIS GPU ----------------------------
def library_baseline(args, dev, content_levels, style_levels, init):
    """What a user of the reference gets on the same B200 by setting device='cuda': the oracle's restatement of the
    reference closure (torch modules + autograd, neural_style_transfer.py:152-202 without the dead CPU randn, anomaly
    mode and prints — i.e. the reference's BEST case) under torch Adam, plus gram_matrix + MSELoss and the
    F.interpolate chain alone.  CUDA events; these are baselines, not the product path."""
    import torch
    import torch.nn.functional as F
    from oracle import gatys_oracle as O
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, iters=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    # B3: gram_matrix + MSELoss forward + backward alone, the five tap shapes of the top level
    top = 4 ** (args.levels - 1)
    gram = {}
    for c, base in ((64, 98304), (128, 24576), (256, 6144), (512, 1536), (512, 384)):
        hw = base * top
        g = torch.Generator(device=dev).manual_seed(c + hw)
        x = (torch.relu(torch.randn((1, c, 1, hw), generator=g, device=dev)) * 0.25).requires_grad_(True)
        a = torch.rand((c, c), device=dev) * 1e-3

        def fwdbwd():
            x.grad = None
            f = x.view(1, c, hw)
            gm = f.bmm(f.transpose(1, 2))
            gm = gm / (c * hw)
            F.mse_loss(a, gm[0]).backward()
        ms = timeit(fwdbwd, iters=3)
        gram[f'{c}x{hw}'] = {'fwdbwd_ms': round(ms, 4), 'TFLOPs': round(4.0 * c * c * hw / ms / 1e9, 1)}
        del x
    out['gram_mse_fwdbwd_torch_fp32'] = gram
    # B4: the bicubic chain (levels-1 transitions) forward + backward
    H, W = init.shape[0], init.shape[1]
    img = torch.randn((1, 3, H, W), device=dev, requires_grad=True)

    def chain():
        img.grad = None
        t, acc = img, 0.0
        for _ in range(args.levels - 1):
            t = F.interpolate(t, size=(t.shape[2] // 2, t.shape[3] // 2), mode='bicubic')
            acc = acc + t.sum()
        if args.levels > 1:
            acc.backward()
    if args.levels > 1:
        ms = timeit(chain, iters=5)
        byt = sum(2 * 15.0 * (H >> i) * (W >> i) for i in range(args.levels - 1))
        out['interpolate_chain_fwdbwd'] = {'ms': round(ms, 4), 'GBps': round(byt / ms / 1e6, 1)}
    del img
    # B6: the closure + Adam, end to end on the device
    net, cidx, sidx = O.make_vgg19(1234)
    net = net.to(dev)
    targets = [O.torch_targets(net, cidx, sidx, torch.from_numpy(O.prepare_img(c)).to(dev),
                               torch.from_numpy(O.prepare_img(s_)).to(dev)) for c, s_ in zip(content_levels, style_levels)]
    x = torch.from_numpy(O.prepare_img(init)).to(dev).requires_grad_(True)
    opt = torch.optim.Adam((x,), lr=10.0)

    def step():
        for g in opt.param_groups:
            g['lr'] *= 0.999
        opt.zero_grad()
        _, _, grad = O.torch_closure(net, cidx, sidx, targets, x, WEIGHTS)
        x.grad = grad
        opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    out['closure_adam'] = {'value': round(1e3 / ms, 3), 'unit': UNIT, 'ms_per_step': round(ms, 3), 'steps': n,
                           'what': 'oracle port of the reference closure (torch modules + autograd, NCHW, cuDNN TF32 '
                                   'convs as torch defaults, fp32 bmm) + torch Adam, device=cuda, CUDA events'}
    del net, targets, x, opt
    torch.cuda.empty_cache()
    return out


# ---- CPU baseline: the oracle's restatement of the reference closure on the host cores ------------------------
def cpu_closure_setup(levels, optimizer='adam', init='structured'):
    """The stated job on the host: same synthetic images, same pyramid (the oracle's cv2-equivalent resize), same
    structured-noise init (np.random.seed(0)), random-init VGG19 seed 1234, torch Adam lr 10 x 0.999 per closure."""
    import numpy as np
    import torch
    from oracle import gatys_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    content_src, style_src = synthetic_sources()
    net, cidx, sidx = O.make_vgg19(1234)
    c_lv = [O.resize_to_level(content_src, lv) for lv in reversed(range(levels))]
    s_lv = [O.resize_to_level(style_src, lv) for lv in reversed(range(levels))]
    targets = [O.torch_targets(net, cidx, sidx, torch.from_numpy(O.prepare_img(c)), torch.from_numpy(O.prepare_img(s)))
               for c, s in zip(c_lv, s_lv)]
    np.random.seed(0)
    method, nf, nl, central, peripheral, dispersion = PIXEL_INIT_ARGS if init == 'pixel' else INIT_ARGS
    init_img = O.structured_noise_init(c_lv[0], s_lv[0], init_method=method, noise_factor=nf, noise_levels=nl,
                                       central=central, peripheral=peripheral, dispersion=dispersion,
                                       use_normal_noise=init == 'pixel')
    img = torch.from_numpy(O.prepare_img(init_img)).requires_grad_(True)
    if optimizer == 'adam':
        opt = torch.optim.Adam((img,), lr=10.0)
    else:   # neural_style_transfer.py:136 — one LBFGS step = 2 closures under torch >= 2.x
        opt = torch.optim.LBFGS((img,), max_iter=1, line_search_fn='strong_wolfe', lr=10.0)
    last = [None]

    def closure():
        for g in opt.param_groups:
            g['lr'] *= 0.999
        opt.zero_grad()
        total, _, grad = O.torch_closure(net, cidx, sidx, targets, img, WEIGHTS)
        img.grad = grad
        last[0] = float(total)
        return total

    def step():
        opt.step(closure)
        return last[0]
    return step


def pixel_ratio(levels_full, levels_sample):
    full = sum(4.0 ** l for l in range(levels_full))
    samp = sum(4.0 ** l for l in range(levels_sample))
    return full / samp


def cpu_baseline(args):
    """Bounded sample for the default run: ONE real closure + Adam update of the stated job (all `levels` levels, the
    real image sizes) after one untimed warm-up step — ~10-20 s on a GPU box's host."""
    step = cpu_closure_setup(args.levels, args.optimizer, args.init)      # builds the targets: 2 x levels VGG forwards warm the conv path up
    t0 = time.perf_counter()
    step()
    dt = time.perf_counter() - t0
    return {'value': round(1.0 / dt, 5), 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
            'sample': f'1 real step (1 closure of the stated {args.levels}-level job, top level {256 * 2 ** (args.levels - 1)}x'
                      f'{384 * 2 ** (args.levels - 1)}, + Adam), {dt:.2f} s on {os.cpu_count()} threads; oracle port of the '
                      'reference closure (no dead randn, anomaly mode or prints); `--impl reference` times K such steps',
            'sample_s_per_step': round(dt, 4)}


def run_reference(args):
    """The reference's own algorithm for this path (oracle port: torch-CPU closure restating
    neural_style_transfer.py:84-112, :152-193 + the Adam update) on the host cores; /root/reference itself does
    not exist on the GPU box.  Every step is a REAL step of the stated job (all levels, full image sizes) when
    warmup + steps of it fit --cpu-budget-s (they do on a GPU box: ~8 s per step at L=3); only otherwise the steps
    are a bounded sample (fewer levels, scaled by pixel count) and the line says so."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t_setup = time.perf_counter()
    step = cpu_closure_setup(args.levels, args.optimizer, args.init)
    t0 = time.perf_counter()
    step()                                     # first warm-up step doubles as the probe of a real step's duration
    probe = time.perf_counter() - t0
    real = probe * (args.warmup + args.steps) <= args.cpu_budget_s or args.levels == 1
    sample_levels, ratio = args.levels, 1.0
    if not real:
        sample_levels = args.levels
        while sample_levels > 1 and probe / pixel_ratio(args.levels, sample_levels) * (args.warmup + args.steps) > args.cpu_budget_s:
            sample_levels -= 1
        ratio = pixel_ratio(args.levels, sample_levels)
        del step
        step = cpu_closure_setup(sample_levels, args.optimizer, args.init)
        step()
    for _ in range(max(args.warmup - 1, 0)):
        step()
    t0 = time.perf_counter()
    loss = None
    for _ in range(args.steps):
        loss = step()
    dt = (time.perf_counter() - t0) / args.steps
    value = 1.0 / (dt * ratio)
    if real:
        sample = (f'{args.steps} real steps of the stated {args.levels}-level job at {dt:.3f} s each on {os.cpu_count()} '
                  'threads (no sampling, no scaling)')
    else:
        sample = (f'{args.steps} closures of the {sample_levels}-level pyramid at {dt:.3f} s each on {os.cpu_count()} '
                  f'threads, scaled to the {args.levels}-level job by the pixel ratio {ratio:.2f} (a real step took '
                  f'{probe:.1f} s: {args.warmup + args.steps} of them exceed the {args.cpu_budget_s:.0f} s budget)')
    line = {'impl': 'reference', 'metric': metric_name(args.levels), 'value': round(value, 5), 'unit': UNIT,
            'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': round(dt * ratio * 1e3, 1), 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args),
            'impl_config': {'what': f'CPU, oracle port of the reference closure (torch modules + autograd) + torch {args.optimizer}',
                            'sampled': not real, 'sample_levels': sample_levels, 'threads': os.cpu_count(),
                            'setup_s': round(t0 - t_setup, 1)},
            'loss_after': loss,
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
