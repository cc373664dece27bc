"""Drop-in for the reference's neural_style_transfer.py: same classes, functions, signatures, generator
protocol and error behaviour, with the pyramid Gatys-loss hot path running in libast_sm100.so.

Kept from the reference (file:line into its neural_style_transfer.py):
  ContentStylePair :32-36 · RepresentationBuilder :39-63 · LossBuilder :66-112 · NeuralStyleTransfer.process
  :115-208 (async generator, torch Adam/LBFGS driver, lr*0.999 per closure, `step` counts closures,
  optimizer.step runs on the asyncio default executor) · resize :211-226 · neural_style_transfer :229-372 ·
  prepare_img/unprepare_img :375-393 · gaussian_mask :396-418 · make_style_noise :422-439 · the constants and
  the four demo flags :22-29.

Dropped because they are exactly zero or pure overhead: the `0 * clip(randn)` tensor built on the
CPU every closure (:91-93), set_detect_anomaly(True) (:150), the per-level `.item()` prints (:159-196; enable
with VERBOSE = True), the deepcopy before the per-step device->host copy (:207).
"""
from __future__ import annotations

import asyncio
import concurrent.futures
import os
import time
import threading
import traceback

import numpy as np
import torch
from torch import Tensor
from torch.autograd import Variable
from torch.optim import Adam, LBFGS

from . import feature_path, math_utils, ops
from . import parallel as _parallel

# ImageNet statistics (:22-23)
IMAGENET_MEAN_255 = [123.675, 116.28, 103.53]
IMAGENET_STD_NEUTRAL = [1, 1, 1]

# Flags for debug/demonstration (:26-29)
USE_NORMAL_NOISE_JUST_FOR_DEMONSTRATION = False
WITHOUT_GAUSSIAN_MASK_JUST_FOR_DEMONSTRATION = False
SHOW_TEST_IMGS = False          # accepted for compatibility; the debug JPEG dumps are not produced
IGNORE_GRADIENT_MAP_JUST_FOR_DEMONSTRATION = False

VERBOSE = False                 # True restores the reference's per-closure prints (adds host syncs)
PRECISION = None                # None -> ops.DEFAULT_PRECISION ('tf32'); 'fp32' for the exact path; 'bf16': TF32 + bfloat16
                                # operands in the backward of the 512-channel layers (AST_PREC_BF16)
# True: a level runs as the explicit channels-last schedule of feature_path.py (cuDNN convs without layout
# transposes + this library's glue and (HW, C) Gram kernels).  False: torch modules + autograd around the NCHW
# kernels (also what 'fp32' precision and non-Vgg19 feature nets use).
CHANNELS_LAST_PATH = os.environ.get('AST_CHANNELS_LAST', '1') != '0'
# True: after GRAPH_WARMUP eager closures the whole closure (every level forward + backward, the pyramid
# resampling, the partial-Gram / halo / image-gradient collectives) is captured ONCE into a CUDA graph and
# replayed: a closure is ~1 000 launches, and at the small pyramid levels (and on row bands) the host cannot
# issue them as fast as the GPU retires them.
GRAPH_CLOSURE = os.environ.get('AST_CUDA_GRAPH', '1') != '0'
# A failed capture normally falls back to eager launches (a feature net that synchronises cannot be captured).
# When the graph was asked for explicitly (AST_CUDA_GRAPH=1 in the environment, or AST_CUDA_GRAPH_STRICT=1, or
# bench.py) the failure is raised instead: a silently eager run is only visible as a slower number.
GRAPH_STRICT = os.environ.get('AST_CUDA_GRAPH') == '1' or os.environ.get('AST_CUDA_GRAPH_STRICT', '0') == '1'
GRAPH_WARMUP = 2
FUSED_ADAM = os.environ.get('AST_FUSED_ADAM', '1') != '0'
# True: the bicubic pyramid step that reads a level's image also computes that level's total variation
# (ast_bicubic_down2x_tv): one pass over every image but the smallest instead of two.
FUSED_PYRAMID_TV = os.environ.get('AST_FUSED_PYRAMID_TV', '1') != '0'
# True: process() takes the per-step image snapshot (:207-208) off the critical path — one fused unprepare kernel into
# a device staging buffer, the device->host copy on a side stream into page-locked memory, and the NEXT optimizer
# step is enqueued before the copy is awaited, so the GPU never idles on the yield.  The sequence of yielded
# (image, step) pairs is the reference's.  False: snapshot, copy and wait before the next step starts.
ASYNC_YIELD = os.environ.get('AST_ASYNC_YIELD', '1') != '0'
# Row-band sharding: every rank runs the same loop and holds the same image; only rank 0 copies it to the host
# (the other ranks yield None for the image) unless this is set.
YIELD_ON_ALL_RANKS = os.environ.get('AST_YIELD_ALL_RANKS', '0') == '1'
YIELD_TRACE = [] if os.environ.get('AST_YIELD_TRACE', '0') == '1' else None   # (what, step, host ms) per step, for bench.py
YIELD_PREWARM = int(os.environ.get('AST_YIELD_PREWARM', '4'))     # page-locked blocks cached at job start (_ImageYielder)

# The reference runs up to simultaneous_tasks_count = 2 jobs in one process (config.py:1, task_executor.py:9): their
# closures come from different executor threads, all on the device's default (legacy) stream.  While one job captures
# its closure into a CUDA graph NO other thread may touch the device: an allocation breaks a global-mode capture, and
# any launch on the legacy stream is an implicit dependency on the capturing streams (cudaErrorStreamCaptureImplicit,
# measured on a B200: it invalidates the capture and poisons the autograd engine thread the two jobs share).  So
# every host-side GPU section of a job — set-up, init image, resize, one optimizer step (closure + update), the
# yield's snapshot — holds this process-wide lock.  The GPU executes the two jobs' kernels back to back on one stream
# anyway; the lock only serialises their host-side enqueueing.
_GPU_SETUP_LOCK = threading.RLock()


class ContentStylePair:
    """ Pairs content image - style image """
    def __init__(self, content, style):
        self.content = content      # (content_img_name, content_img)
        self.style = style          # (style_img_name, style_img)


class RepresentationBuilder:
    """Content / style representations of an image from the network's feature maps (:39-63)."""
    def __init__(self, image, neural_net):
        self.__image = image
        self.__neural_net = neural_net
        self.__features = neural_net(image)

    def build_content(self, feature_map_indices: int | list[int]):
        list_taken = isinstance(feature_map_indices, list)
        indices = feature_map_indices if list_taken else [feature_map_indices]
        rep = [x.squeeze(axis=0) for index, x in enumerate(self.__features) if index in indices]
        return rep if list_taken else rep[0]

    def build_style(self, feature_map_indices: int | list[int]):
        list_taken = isinstance(feature_map_indices, list)
        indices = feature_map_indices if list_taken else [feature_map_indices]
        rep = [math_utils.gram_matrix(x, precision=PRECISION) for index, x in enumerate(self.__features)
               if index in indices]
        return rep if list_taken else rep[0]


class LossBuilder:
    """One pyramid level's loss (:66-112).  Targets are built once at construction (:78-82); build() runs the
    VGG forward on torch/cuDNN and then ONE fused autograd node: 5 x (split-K tcgen05 Gram + fused MSE),
    content MSE, TV, weighted sum — and, backward, (G-A)F per layer, 2(X-T)/n and the TV gradient."""
    def __init__(self, content_feature_maps_index, style_feature_maps_indices, target_content_image,
                 target_style_image, neural_net, content_weight, style_weight, tv_weight):
        self.__content_feature_maps_index = content_feature_maps_index
        self.__style_feature_maps_indices = style_feature_maps_indices
        self.__neural_net = neural_net
        self.__content_weight = content_weight
        self.__style_weight = style_weight
        self.__tv_weight = tv_weight
        with torch.no_grad():
            content_rep_builder = RepresentationBuilder(image=target_content_image, neural_net=neural_net)
            self.__target_content_representation = \
                content_rep_builder.build_content(content_feature_maps_index).detach().clone().contiguous()
            del content_rep_builder
            style_rep_builder = RepresentationBuilder(image=target_style_image, neural_net=neural_net)
            self.__target_style_representation = style_rep_builder.build_style(style_feature_maps_indices)
            del style_rep_builder
        self.__target_grams = [g[0].detach().contiguous() for g in self.__target_style_representation]
        self.__wss = ops.LevelWorkspaces()
        # channels-last schedule: targets are rebuilt through the same path on first use
        self.__target_images = (target_content_image, target_style_image)
        self.__path_targets = None
        self.shard = None                   # set by parallel.maybe_shard for row-band sharded levels
        self.replicated_rank0_only = False  # sharding on, level not shardable: ranks != 0 evaluate without grad

    @property
    def target_content_representation(self):
        return self.__target_content_representation

    @property
    def target_style_representation(self):
        return self.__target_style_representation

    @property
    def target_images(self):
        return self.__target_images

    def path_plan(self, optimizing_img):
        return self.__path_plan(optimizing_img)

    def __path_plan(self, optimizing_img):
        """The channels-last plan when this level can use it: TF32 Gram operands, a frozen Vgg19 on CUDA, batch 1."""
        if not CHANNELS_LAST_PATH or ops._prec(PRECISION) not in (ops.L.AST_PREC_TF32, ops.L.AST_PREC_BF16):
            return None
        if not (torch.is_tensor(optimizing_img) and optimizing_img.is_cuda and optimizing_img.dim() == 4
                and optimizing_img.shape[0] == 1 and optimizing_img.dtype == torch.float32):
            return None
        if isinstance(self.__style_feature_maps_indices, int) or isinstance(self.__content_feature_maps_index, list):
            return None
        return feature_path.plan_for(self.__neural_net)

    @property
    def workspaces(self):
        return self.__wss

    def build(self, optimizing_img, _tv=None):
        """_tv (internal): (sums2, tv) of optimizing_img already computed by the fused pyramid step
        (ops.bicubic_half_tv) — the level then skips its own total-variation pass."""
        if self.shard is not None:
            return self.shard.build(optimizing_img)
        if self.replicated_rank0_only and torch.is_grad_enabled():
            with torch.no_grad():
                return self.build(optimizing_img)
        plan = self.__path_plan(optimizing_img)
        if plan is not None:
            if self.__path_targets is None:
                self.__path_targets = feature_path.build_targets(
                    plan, self.__target_images[0], self.__target_images[1], self.__content_feature_maps_index,
                    self.__style_feature_maps_indices, self.__wss)
            cfg = (plan, self.__path_targets, self.__content_feature_maps_index,
                   tuple(self.__style_feature_maps_indices),
                   (self.__content_weight, self.__style_weight, self.__tv_weight), self.__wss,
                   ops._prec(PRECISION) == ops.L.AST_PREC_BF16, _tv)
            return feature_path.LevelPathFn.apply(cfg, optimizing_img)
        feats = self.__neural_net(optimizing_img)
        cfg = (self.__target_content_representation, self.__target_grams,
               (self.__content_weight, self.__style_weight, self.__tv_weight), self.__wss, ops._prec(PRECISION))
        total_loss, content_loss, style_loss, tv_loss = ops.LevelLossFn.apply(
            cfg, optimizing_img, feats[self.__content_feature_maps_index],
            *[feats[k] for k in self.__style_feature_maps_indices])
        return total_loss, content_loss, style_loss, tv_loss


class _Job:
    """What NeuralStyleTransfer.process sets up (:124-147) and its closure (:152-202): the leaf image, the torch
    optimizer, one LossBuilder per pyramid level.  Split out so that bench.py can time closures directly."""

    def __init__(self, device, model_name, style_imgs, optimizer_name, content_imgs, init_img, lr_start,
                 content_weight, style_weight, tv_weight, init_img_name):
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('artstyletransfer_b200 runs the Gatys-loss path on sm_100a CUDA only; '
                               f'got device {device} (no CPU fallback)')
        if optimizer_name not in ('adam', 'lbfgs'):
            raise RuntimeError("Unknown optimizer")
        with _GPU_SETUP_LOCK:
            self._setup(device, model_name, style_imgs, optimizer_name, content_imgs, init_img, lr_start,
                        content_weight, style_weight, tv_weight, init_img_name)

    def _setup(self, device, model_name, style_imgs, optimizer_name, content_imgs, init_img, lr_start,
               content_weight, style_weight, tv_weight, init_img_name):
        neural_net, content_feature_maps_index, style_feature_maps_indices = \
            math_utils.prepare_model(model_name, device)
        if VERBOSE:
            print(f'Using {model_name} in the optimization procedure.')
        init_img = prepare_img(init_img, device)
        # we are tuning optimizing_img's pixels! (that's why requires_grad=True)
        self.optimizing_img = Variable(init_img, requires_grad=True)
        if optimizer_name == 'adam':
            # torch's own Adam as in the reference (:133-134); fused=True is the same update rule as ONE kernel
            # instead of ~10 foreach launches over the 75 MB image (0.36 -> 0.09 ms per step at 2048x3072)
            self.optimizer = Adam((self.optimizing_img,), lr=lr_start, fused=FUSED_ADAM)
        else:
            self.optimizer = LBFGS((self.optimizing_img,), max_iter=1, line_search_fn='strong_wolfe', lr=lr_start)
        self.loss_builders = []
        for content_img, style_img in zip(content_imgs, style_imgs):
            content_img = prepare_img(content_img, device)
            style_img = prepare_img(style_img, device)
            self.loss_builders.append(LossBuilder(content_feature_maps_index, style_feature_maps_indices, content_img,
                                                  style_img, neural_net, content_weight, style_weight, tv_weight))
        self.sharded_levels = _parallel.maybe_shard(self.loss_builders, self.optimizing_img, neural_net,
                                                    content_feature_maps_index, style_feature_maps_indices,
                                                    (content_weight, style_weight, tv_weight))
        self.pyramid = _parallel.lockstep_pyramid(self.loss_builders) if self.sharded_levels else None
        self.step = 0
        self.optimizer_name = optimizer_name
        self.init_img_name = init_img_name
        self.weights = (content_weight, style_weight, tv_weight)
        self._graph = None            # (CUDAGraph, static total loss, static image gradient) once captured
        self._graph_failed = False
        self._eager_closures = 0

    # ---- the closure (:152-202) --------------------------------------------------------------------------
    def _evaluate(self):
        """All levels forward, summed loss, backward, image-gradient sync: everything a closure launches."""
        optimizing_img, loss_builders = self.optimizing_img, self.loss_builders
        content_weight, style_weight, tv_weight = self.weights
        optimizing_img_levels = None
        total_loss = None
        if self.pyramid is not None and not VERBOSE:
            # every level is row-band sharded: all levels in lock-step, grouped halo exchanges (sharded_path.py)
            total_loss = self.pyramid.evaluate(optimizing_img)
            loss_builders = []
        # lower resolutions of optimizing_img: chained bicubic 2x down (:168-176).  The whole chain first: the step that
        # reads level i-1 to produce level i also hands back total_variation(level i-1), which level i-1's loss needs
        tvs = [None] * len(loss_builders)
        for i in range(len(loss_builders)):
            if i == 0:
                optimizing_img_levels = [optimizing_img]
                continue
            prev = optimizing_img_levels[i - 1]
            fused = (FUSED_PYRAMID_TV and prev.shape[-2] % 2 == 0 and prev.shape[-1] % 2 == 0
                     and loss_builders[i - 1].path_plan(prev) is not None and loss_builders[i - 1].shard is None)
            if fused:
                nxt, sums2, tv = ops.bicubic_half_tv(prev, loss_builders[i - 1].workspaces)
                tvs[i - 1] = (sums2, tv)
            else:
                nxt = ops.bicubic_half(prev)
            optimizing_img_levels.append(nxt)
        for i in range(len(loss_builders)):
            total_loss_l, content_loss, style_loss, tv_loss = loss_builders[i].build(optimizing_img_levels[i], _tv=tvs[i])
            if total_loss is None:
                total_loss = total_loss_l
            else:
                previous_loss_importance = 1.0
                total_loss = previous_loss_importance * total_loss + total_loss_l
            if VERBOSE:
                with torch.no_grad():
                    print(f' - level {i} | level loss={total_loss_l.item():.3e}, '
                          f'content_loss={content_weight * content_loss.item():.3e}, '
                          f'style loss={style_weight * style_loss:.3e}, '
                          f'tv loss={tv_weight * tv_loss.item():.3e}')
        if total_loss.requires_grad:
            total_loss.backward()
        if torch.is_grad_enabled():
            gathered = self.pyramid is not None and not VERBOSE and getattr(self.pyramid, 'gather', None) is not None
            _parallel.sync_image_grad(optimizing_img, already_global=gathered)
        return total_loss

    def _graph_eligible(self):
        return (GRAPH_CLOSURE and not VERBOSE and not self._graph_failed and torch.is_grad_enabled()
                and not ops.STATS.enabled and self._eager_closures >= GRAPH_WARMUP)

    def _capture(self):
        """Capture one closure into a CUDA graph.  The image leaf is updated in place by Adam / LBFGS, so its
        storage is the graph's static input; the summed loss and the image gradient are its static outputs."""
        dev = self.optimizing_img.device
        with _GPU_SETUP_LOCK:                    # no other job sets up / captures while this capture is open
            self.optimizing_img.grad = None      # backward() inside the capture creates .grad in the graph's pool
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            # thread_local: another job's executor thread may still allocate (its eager closures) during the capture
            with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                total = self._evaluate()
            static_total = total.detach()
            static_grad = self.optimizing_img.grad
            if static_grad is None:
                raise RuntimeError('closure capture produced no image gradient')
            self._graph = (graph, static_total, static_grad)

    def closure(self):
        try:
            optimizer, optimizing_img = self.optimizer, self.optimizing_img
            # learning rate schedule (:155-159)
            lr = 0
            for g in optimizer.param_groups:
                g['lr'] *= 0.999
                lr = g['lr']
            if VERBOSE:
                print(f"new lr = {lr}")
                print(f'{self.optimizer_name} | processing image: {self.init_img_name} | iteration: {self.step:03} :')
            if self._graph is None and self._graph_eligible():
                try:
                    self._capture()
                except Exception:
                    # leave capture mode cleanly and keep working eagerly (e.g. a feature net that syncs)
                    self._graph_failed = True
                    self._graph = None
                    optimizing_img.grad = None
                    if GRAPH_STRICT:
                        raise
                    traceback.print_exc()
            if self._graph is not None and torch.is_grad_enabled() and not ops.STATS.enabled and not VERBOSE:
                graph, static_total, static_grad = self._graph
                graph.replay()
                optimizing_img.grad = static_grad      # the replay refilled it (assignment, not accumulation)
                total_loss = static_total
            else:
                if torch.is_grad_enabled():
                    optimizer.zero_grad()
                # detached: a caller that keeps the returned loss must not keep the autograd graph — and with it the
                # leaf's AccumulateGrad node, created on the default stream — alive into the CUDA-graph capture, where
                # a node of the legacy stream is an illegal dependency on the capturing stream
                total_loss = self._evaluate().detach()
                self._eager_closures += 1
            if VERBOSE:
                with torch.no_grad():
                    print(f'{self.optimizer_name} | total loss={total_loss.item():.3e}')
            self.step += 1
            return total_loss
        except:
            traceback.print_exc()
            raise

    def optimizer_step(self):
        with _GPU_SETUP_LOCK:
            return self.optimizer.step(self.closure)


class NeuralStyleTransfer:
    """ The main class for calculating of artistic style transfer (:115-208) """
    def __init__(self, device, model_name, style_imgs, optimizer_name):
        self.__device = device
        self.__model_name = model_name
        self.__style_imgs = style_imgs
        self.__optimizer_name = optimizer_name

    async def process(self, content_imgs, init_img, lr_start, iters_num, content_weight, style_weight, tv_weight,
                      init_img_name):
        job = _Job(self.__device, self.__model_name, self.__style_imgs, self.__optimizer_name, content_imgs, init_img,
                   lr_start, content_weight, style_weight, tv_weight, init_img_name)
        loop = asyncio.get_running_loop()
        # The reference hands optimizer.step to the loop's default executor (:206).  Here every step of a job runs on
        # ONE worker thread of its own: cuDNN handles (and their device workspaces) are per thread in torch, and a
        # step that landed on a fresh pool thread inside the CUDA-graph capture would make cuDNN allocate there
        # (CUDNN_STATUS_INTERNAL_ERROR_DEVICE_ALLOCATION_FAILED, seen on a B200 once the pool had grown).
        worker = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix='ast-job')
        steps = self.__steps(job, loop, worker, iters_num)
        try:
            async for item in steps:
                yield item
        finally:
            await steps.aclose()             # lets a look-ahead step in flight finish before its thread goes away
            worker.shutdown(wait=False)

    @staticmethod
    async def __steps(job, loop, worker, iters_num):
        if not ASYNC_YIELD:
            # the main optimization loop (:205-208)
            while job.step < iters_num:
                await loop.run_in_executor(worker, job.optimizer_step)
                with _GPU_SETUP_LOCK:
                    img = unprepare_img(job.optimizing_img)
                yield img, job.step
            return
        # Same loop, same (image, step) sequence, with the yield overlapped: after step k has been enqueued its image
        # is snapshotted on the device (in stream order, before step k+1 can touch it), step k+1 is handed to the
        # executor, and only then is the snapshot's device->host copy awaited and yielded.
        yielder = _ImageYielder(job.optimizing_img)

        def step_and_snapshot():
            """Executor thread: one optimizer step, then the snapshot of its image — enqueued back to back under the
            GPU lock, so the next step (this job's or another job's) cannot slip in between."""
            with _GPU_SETUP_LOCK:
                t_step = time.perf_counter()
                had_graph = job._graph is not None
                job.optimizer_step()
                if job._graph is not None and not had_graph:
                    yielder.prewarm()        # the capture emptied torch's page-locked cache
                t_snap = time.perf_counter()
                ticket = yielder.begin()
                if YIELD_TRACE is not None:
                    now = time.perf_counter()
                    YIELD_TRACE.append(('optimizer_step_host', job.step, 1e3 * (t_snap - t_step)))
                    YIELD_TRACE.append(('snapshot_host', job.step, 1e3 * (now - t_snap)))
                return ticket, job.step

        pending = loop.run_in_executor(worker, step_and_snapshot) if job.step < iters_num else None
        try:
            while pending is not None:
                fut, pending = pending, None
                ticket, step = await fut
                if step < iters_num:
                    pending = loop.run_in_executor(worker, step_and_snapshot)    # step k+1 runs while image k travels
                img = await loop.run_in_executor(None, yielder.finish, ticket)
                yield img, step
        finally:
            if pending is not None:          # the consumer stopped early: let the step in flight finish quietly
                try:
                    await pending
                except Exception:
                    pass


class _ImageYielder:
    """The per-step image yield of process() (:207-208: deepcopy, unprepare_img, device->host) off the critical path.
    begin() — called right after an optimizer step, under the GPU lock — enqueues ONE kernel on the current stream that writes the
    unprepared (H, W, 3) [0,1] image into one of two device staging buffers (that is the snapshot: the next step may
    overwrite the leaf as soon as it has run), and the device->host copy into a fresh page-locked block on a side
    stream.  finish() waits for that copy only.  Under row-band sharding every rank holds the same image; rank 0
    alone copies (YIELD_ON_ALL_RANKS restores the copy everywhere), the others yield None."""

    def __init__(self, optimizing_img: Tensor):
        self.img = optimizing_img
        dev = optimizing_img.device
        _, c, h, w = optimizing_img.shape
        rank, world = _parallel.world()
        self.active = YIELD_ON_ALL_RANKS or world == 1 or rank == 0
        self.fused = c == 3 and (h * w) % 4 == 0 and optimizing_img.is_contiguous()
        self.shape = (h, w, c)
        self.n = 0
        if self.active:
            with _GPU_SETUP_LOCK:
                self.stage = [torch.empty(self.shape, dtype=torch.float32, device=dev) for _ in range(2)]
                self.side = torch.cuda.Stream(dev)
            self.drained = [None, None]      # event: the copy that last read stage[i] has finished
            self.prewarm(min(YIELD_PREWARM, 3))

    def prewarm(self, n: int = None):
        """Every yield hands out a FRESH page-locked block (the consumer may keep the array).  Torch's host allocator
        caches freed blocks, but page-locking a new 75 MB block costs ~50 ms (measured on the B200 boxes) — ten 8-GPU
        steps.  Fill the cache with as many blocks as are ever alive at once (the consumer's current and previous
        image, the look-ahead copy in flight, blocks whose copy event is still pending).  Called at job start, and
        AGAIN right after the closure's CUDA graph has been captured: torch.cuda.graph() empties the page-locked cache
        (torch._C._host_emptyCache) on entry, which used to put three fresh allocations into the steps after it."""
        if not self.active:
            return
        with _GPU_SETUP_LOCK:
            warm = [torch.empty(self.shape, dtype=torch.float32, pin_memory=True)
                    for _ in range(YIELD_PREWARM if n is None else n)]
            del warm

    def begin(self):
        if not self.active:
            return None
        dev = self.img.device
        i = self.n & 1
        self.n += 1
        main = torch.cuda.current_stream(dev)
        if self.drained[i] is not None:
            main.wait_event(self.drained[i])
        if self.fused:
            ops.unprepare_hwc(self.img.detach(), self.stage[i], IMAGENET_MEAN_255)
        else:
            mean = torch.tensor(IMAGENET_MEAN_255, dtype=torch.float32, device=dev).view(1, 3, 1, 1)
            self.stage[i].copy_(((self.img.detach() + mean) / 255).permute([0, 2, 3, 1]).squeeze(0))
        ready = torch.cuda.Event()
        ready.record(main)
        t_alloc = time.perf_counter()
        with _GPU_SETUP_LOCK:                # a pinned block may be a fresh cudaHostAlloc: not during a capture
            host = torch.empty(self.shape, dtype=torch.float32, pin_memory=True)
        if YIELD_TRACE is not None:
            YIELD_TRACE.append(('page_locked_block', self.n, 1e3 * (time.perf_counter() - t_alloc)))
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            host.copy_(self.stage[i], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.side)
        self.drained[i] = done
        return host, done

    @staticmethod
    def finish(ticket):
        if ticket is None:
            return None
        host, done = ticket
        done.synchronize()
        return host.numpy()


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('artstyletransfer_b200 needs a CUDA device (B200, sm_100a); no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def level_size(img, level):
    """(new_height, new_width) of pyramid `level` (:213-224; int() truncation of the long side)."""
    base_diameter = 256
    current_height, current_width = img.shape[:2]
    if current_height >= current_width:
        base_width = base_diameter
        base_height = int(base_width * (current_height / current_width))
    else:
        base_height = base_diameter
        base_width = int(base_height * (current_width / current_height))
    return base_height * pow(2, level), base_width * pow(2, level)


def _resize_hwc_device(img_dev, new_h, new_w):
    return ops.bicubic_resize(img_dev, new_h, new_w, layout='hwc', coord='cv2')


async def resize(img, level):
    """ A function for proper resizing of an image according to the level of pyramid (:211-226).
    Same result as cv2.resize(..., INTER_CUBIC) on float32 images, computed by the K6 kernel. """
    new_height, new_width = level_size(img, level)
    dev = _device()
    with _GPU_SETUP_LOCK:
        src = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)).to(dev)
        if src.dim() == 2:
            src = src.unsqueeze(-1)
        out = _resize_hwc_device(src, new_height, new_width).cpu().numpy()
    return out if img.ndim == 3 else out[:, :, 0]


async def neural_style_transfer(content_n_style: ContentStylePair,
                                content_weight, style_weight, tv_weight,
                                optimizer, model, init_method,
                                iters_num, levels_num, noise_factor, noise_levels, noise_levels_central_amplitude,
                                noise_levels_peripheral_amplitude, noise_levels_dispersion):
    """ The main function (:229-372) """
    device = _device()
    model_name = model
    optimizer_name = optimizer

    # pyramids of the content and style images, high -> low resolution (:250-263)
    level = 0
    content_img_levels = [await resize(content_n_style.content[1], level=level)]
    style_img_levels = [await resize(content_n_style.style[1], level=level)]
    for level in range(1, levels_num):
        content_img_levels.insert(0, await resize(content_n_style.content[1], level=level))
        style_img_levels.insert(0, await resize(content_n_style.style[1], level=level))

    with _GPU_SETUP_LOCK:
        init_img_next, init_img_name = build_init_image(
            content_n_style, content_img_levels, style_img_levels, init_method, noise_factor, noise_levels,
            noise_levels_central_amplitude, noise_levels_peripheral_amplitude, noise_levels_dispersion, device)

    nst = NeuralStyleTransfer(device, model_name, style_img_levels, optimizer_name)
    lr_start = 10.0
    async for img, cur_iter in nst.process(content_img_levels, init_img_next, lr_start, iters_num, content_weight,
                                           style_weight, tv_weight, init_img_name):
        percent = cur_iter / iters_num * 100.0
        cur_iter += 1
        yield percent, img


def gaussian_kernel_1d(n, sigma):
    """cv2.getGaussianKernel(n, sigma) (CV_64F) as OpenCV >= 4 computes it: exp(-x^2/(2 sigma^2)) normalised,
    with the two centre taps of an EVEN-length kernel set to the x = 0 value before normalisation."""
    if sigma <= 0:
        sigma = ((n - 1) * 0.5 - 1) * 0.3 + 0.8
    i = np.arange(n, dtype=np.float64) - (n - 1) * 0.5
    g = np.exp(-(i * i) / (2.0 * sigma * sigma))
    if n % 2 == 0:
        g[n // 2 - 1] = g[n // 2] = 1.0
    return g / g.sum()


def gaussian_mask(shape, central_amplitude, peripheral_amplitude, dispersion_scale=0.5):
    """ Gaussian envelope for the noise map (:396-418): (rows, cols, 3) float64.  Host helper kept for API parity;
    the init kernel evaluates the same separable envelope on the fly from the two 1-D vectors. """
    rows, cols = shape[:2]
    gx = gaussian_kernel_1d(cols, cols * dispersion_scale)
    gy = gaussian_kernel_1d(rows, rows * dispersion_scale)
    resultant_kernel = np.outer(gy, gx)
    gauss_norm = resultant_kernel / resultant_kernel[rows // 2, cols // 2]
    mask = peripheral_amplitude + gauss_norm * (central_amplitude - peripheral_amplitude)
    return np.repeat(np.expand_dims(mask, 2), 3, axis=2)


def make_style_noise(style_img_np, targ_shape):
    """ Noise map made by randomly permuting the pixels of the (resized) style image (:422-439).
    The resize runs on the device; the permutation uses numpy's legacy global RNG exactly like the reference. """
    nw = targ_shape[1]
    nh = targ_shape[0]
    dev = _device()
    src = torch.from_numpy(np.ascontiguousarray(style_img_np, dtype=np.float32)).to(dev)
    style_img_np_resized = _resize_hwc_device(src, nh, nw).cpu().numpy()
    style_vect = style_img_np_resized.reshape(nh * nw, -1)
    style_noise_vect = np.random.permutation(style_vect)
    return style_noise_vect.reshape(targ_shape)


def build_init_image(content_n_style, content_img_levels, style_img_levels, init_method, noise_factor,
                     noise_levels, noise_levels_central_amplitude, noise_levels_peripheral_amplitude,
                     noise_levels_dispersion, device):
    """Structured-noise initial image (:265-362) -> (HxWx3 float32 numpy, name).  Host: numpy RNG draws, low-res
    grids, 1-D Gaussian vectors.  Device: one fused K7 pass (upsample x envelope accumulation, Sobel map, blend)."""
    if init_method not in ('random', 'content+noise'):
        # init image has same dimension as content image - this is a hard constraint (:358-362)
        return style_img_levels[0], content_n_style.style[0]

    noise_shape = content_img_levels[0].shape
    nw = noise_shape[1]
    nh = noise_shape[0]
    levels = []
    for noise_granularity, central_amplitude, peripheral_amplitude, dispersion_scale in zip(
            noise_levels, noise_levels_central_amplitude, noise_levels_peripheral_amplitude, noise_levels_dispersion):
        entry = {'central': central_amplitude, 'peripheral': peripheral_amplitude}
        with_mask = noise_granularity == 0 or not WITHOUT_GAUSSIAN_MASK_JUST_FOR_DEMONSTRATION
        if with_mask:
            gx = gaussian_kernel_1d(nw, nw * dispersion_scale)
            gy = gaussian_kernel_1d(nh, nh * dispersion_scale)
            entry.update(gy=torch.from_numpy(gy).to(device), gx=torch.from_numpy(gx).to(device),
                         center=float(gy[nh // 2] * gx[nw // 2]))
        if noise_granularity == 0:
            entry['kind'] = 0           # constant level: envelope only (:272-275)
        else:
            if noise_granularity > 0:   # number of noise spots along the shortest axis (:277-287)
                if nh <= nw:
                    noise_shape_div_h = noise_granularity
                    noise_shape_div_w = nw * noise_granularity // nh
                else:
                    noise_shape_div_w = noise_granularity
                    noise_shape_div_h = nh * noise_granularity // nw
            else:                       # spot size in pixels (:288-291)
                noise_shape_div_w = nw // (-noise_granularity)
                noise_shape_div_h = nh // (-noise_granularity)
            noise_shape_div = (noise_shape_div_h, noise_shape_div_w, noise_shape[2])
            if USE_NORMAL_NOISE_JUST_FOR_DEMONSTRATION:
                lowres = np.clip(np.random.normal(loc=0, scale=255, size=noise_shape_div).astype(np.float32) / 255,
                                 0.0, 1.0)
            else:
                lowres = make_style_noise(style_img_levels[0], noise_shape_div)
            entry['lowres'] = torch.from_numpy(np.ascontiguousarray(lowres, dtype=np.float32)).to(device)
            entry['kind'] = 1 if with_mask else 2
        levels.append(entry)

    g101 = gaussian_kernel_1d(101, 0.2)     # GaussianBlur((101,101), sigmaX=0.2) (:340): taps beyond +-1 < 2e-22
    content_dev = None
    if init_method == 'content+noise':
        content_dev = torch.from_numpy(np.ascontiguousarray(content_img_levels[0], dtype=np.float32)).to(device)
    out = ops.noise_init(content_dev, nh, nw, levels, noise_factor, init_method,
                         not IGNORE_GRADIENT_MAP_JUST_FOR_DEMONSTRATION, g101[50], g101[49], device)
    name = 'random' if init_method == 'random' else content_n_style.content[0]
    return out.cpu().numpy(), name


def prepare_img(img, device):
    """ HWC float [0,1] -> (1,3,H,W): x*255 - ImageNet mean, std 1 (:375-383) """
    if isinstance(img, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(img.transpose((2, 0, 1))))
        if t.dtype == torch.uint8:
            t = t.to(torch.float32).div(255)   # torchvision ToTensor semantics for uint8
        t = t.to(torch.float32)
    else:
        t = img.to(torch.float32)
    t = t.to(device)
    mean = torch.tensor(IMAGENET_MEAN_255, dtype=torch.float32, device=t.device).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD_NEUTRAL, dtype=torch.float32, device=t.device).view(3, 1, 1)
    return t.mul(255).sub_(mean).div_(std).unsqueeze(0)


def unprepare_img(img: Tensor):
    """ Reverse of prepare_img (:388-393): (1,3,H,W) tensor -> HxWx3 float32 numpy on the host.
    The mean add and the /255 run on the device, then ONE contiguous device->host copy of the HWC image. """
    t = img.detach()
    if t.is_cuda and t.dim() == 4 and t.shape[0] == 1 and t.shape[1] == 3 and (t.shape[2] * t.shape[3]) % 4 == 0 \
            and t.dtype == torch.float32 and t.is_contiguous():
        hwc = torch.empty((t.shape[2], t.shape[3], 3), dtype=torch.float32, device=t.device)
        ops.unprepare_hwc(t, hwc, IMAGENET_MEAN_255)
    else:
        mean = torch.tensor(IMAGENET_MEAN_255, dtype=torch.float32, device=t.device).view(1, 3, 1, 1)
        hwc = ((t + mean) / 255).permute([0, 2, 3, 1]).squeeze(0).contiguous()
    if not hwc.is_cuda:
        return hwc.numpy()
    # fresh page-locked block from torch's caching host allocator (recycled once the caller drops the array)
    with _GPU_SETUP_LOCK:
        host = torch.empty(hwc.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(hwc, non_blocking=True)
    torch.cuda.current_stream(hwc.device).synchronize()
    return host.numpy()
