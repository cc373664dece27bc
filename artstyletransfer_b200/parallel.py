"""Row-band sharding of the pyramid levels across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Every rank holds the full optimizing image and
an identical optimizer.  The default scheme (sharded_path.ShardedPathLevel on the channels-last feature path):

  * PyramidBands deals the rows of the WHOLE pyramid out to the ranks (a rank owns a band of one or two levels; at 8
    ranks the 2048x3072 level sits on six of them and the three lower levels on the other two);
  * a rank computes its rows plus two halo rows per side and exchanges those with each neighbour before every
    SECOND 3x3 convolution (halo_schedule; the mirrored gradient rows in the backward): 13 neighbour synchronisations
    per closure; the rows travel through NVLink peer memory (PeerHaloGroup) or a grouped NCCL send/recv;
  * the RAW partial Grams F_r F_r^T of the bands (the Gram's K dimension is the spatial index, so the full Gram is
    the plain sum over ranks) and the partial content SSE of all levels are all-reduced (sum) tap by tap under the
    convolutions above the tap; every rank finalises identically: 1/(C*HW), minus target, MSE, weighted total; the
    backward dF_r = s (G - A) F_r is rank-local;
  * the rows of the image-gradient pyramid owned by different ranks are disjoint: they are gathered through peer
    memory (PeerGradGather; an all-reduce over NCCL without it) so the replicated optimizers stay bit-identical.
    The total variation VALUE is evaluated on the whole level image by every rank; its gradient is added by the
    owner of each row before the gather (by rank 0 on the all-reduce path).

The older scheme (ShardedLevel below: torch modules + autograd around the NCHW kernels, used for fp32 precision)
cuts every level into `world` equal bands, pads each with an 80-row halo (>= the 78-row receptive-field radius of
relu5_1) and crops before the Gram — exact, but with redundant convolution work.

The reference has no counterpart (a commented-out two-GPU round-robin, neural_style_transfer.py:238-243).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

HALO = 80            # rows; >= 78 = receptive-field radius of relu5_1, multiple of 16
ALIGN = 16           # VGG19 down-samples by 16 up to relu5_1
LAYER_STRIDE = {0: 1, 1: 2, 2: 4, 3: 8, 4: 8, 5: 16}     # feature index -> spatial stride (neural_nets.py:26-29)


class BandPlan:
    """Pure host logic: which image rows a rank owns and feeds to the network at one level."""

    def __init__(self, height: int, rank: int, world: int, halo: int = HALO):
        if not self.shardable(height, world):
            raise ValueError(f'height {height} cannot be cut into {world} bands of a multiple of {ALIGN} rows')
        band = height // world
        self.height, self.rank, self.world = height, rank, world
        self.r0, self.r1 = rank * band, (rank + 1) * band          # owned rows
        self.lo, self.hi = max(0, self.r0 - halo), min(height, self.r1 + halo)   # rows fed to VGG

    @staticmethod
    def shardable(height: int, world: int) -> bool:
        return world >= 1 and height % world == 0 and (height // world) % ALIGN == 0 and height // world >= 2 * ALIGN

    def feat_rows(self, stride: int):
        """(first owned row, one-past-last owned row, rows present) in a feature map of the given stride."""
        return (self.r0 - self.lo) // stride, (self.r1 - self.lo) // stride, (self.hi - self.lo) // stride

    def global_feat_rows(self, stride: int):
        return self.r0 // stride, self.r1 // stride


# how the halo rows of a lock-step step (and, with them, the image gradient) travel: 'peer' = one ast_halo_exchange /
# ast_band_gather launch over NVLink peer memory (PeerHaloGroup, PeerGradGather; validated between processes at 2 and 8
# GPUs in round 2: bit-identical losses, 8 % faster at 8 GPUs), 'nccl' = grouped send/recv + all-reduce.  'peer' falls
# back to 'nccl' — on every rank together — when torch symmetric memory cannot be set up on the box.
DEFAULT_HALO = 'peer'


def all_ranks_ok(ok: bool, device: torch.device) -> bool:
    """Collective: True only when `ok` holds on every rank (so that all ranks take the same transport)."""
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(int(flag.item()))


BAND_OVERHEAD = float(os.environ.get('AST_BAND_OVERHEAD', '0.015'))
MIN_BAND_ROWS = 2 * ALIGN
# Halo rows a band swaps per exchange on the channels-last path (halo_schedule): 2 = every second convolution runs on
# redundantly computed edge rows instead of waiting for the neighbours (13 synchronisation points per closure instead
# of 25, for 2 extra convolution rows per band and layer); 1 = an exchange before every convolution (round 1).
HALO_DEPTH = min(max(int(os.environ.get('AST_HALO_DEPTH', '2')), 1), 2)


class PyramidBands:
    """Pure host logic: which rows of WHICH LEVEL every rank owns (the level-aware successor of BandPlan).

    Cutting every level into `world` equal bands makes the small levels tiny (a 32-row band of the 256x384 level is
    a handful of CTAs per launch) and multiplies the number of latency-sized launches per rank by the number of
    levels.  The levels' costs are 64 : 16 : 4 : 1, so instead the ranks are filled like a snake: rank 0 takes rows
    of the top level until its share of the whole pyramid's cost is reached, the next rank continues where it
    stopped, and whoever finishes a level carries on with the level below.  At 8 ranks and 4 levels that puts the
    2048x3072 level on ranks 0-5 and the three lower levels on the remaining ranks; most ranks run ONE band.

    Cost model (units: the whole top level = 1): rows * width / (H0 * W0) per band plus `overhead` per band (the
    ~120 launches of a band cost their latency however few rows it has).  The per-rank budget T is scanned and the
    plan with the smallest maximum load wins.  Band edges are multiples of 16 rows (the four 2x2 max-pools never
    straddle ranks), bands are at least 32 rows; bounds[level] is the non-decreasing list of world + 1 row edges,
    an empty band (two equal edges) means the rank does not work on that level."""

    def __init__(self, sizes: Sequence[Sequence[int]], world: int, overhead: Optional[float] = None,
                 uniform: bool = False):
        self.sizes = [(int(h), int(w)) for h, w in sizes]
        self.world = int(world)
        self.overhead = BAND_OVERHEAD if overhead is None else float(overhead)
        for h, w in self.sizes:
            if h % ALIGN or h < MIN_BAND_ROWS:
                raise ValueError(f'level height {h} is not a multiple of {ALIGN} rows >= {MIN_BAND_ROWS}')
        if uniform:
            if not all(BandPlan.shardable(h, world) for h, _ in self.sizes):
                raise ValueError(f'levels {self.sizes} cannot all be cut into {world} equal bands')
            self.bounds = [[r * (h // world) for r in range(world + 1)] for h, _ in self.sizes]
        else:
            self.bounds = self._search()

    # -- planning -------------------------------------------------------------------------------------------
    def _unit(self):
        return float(self.sizes[0][0] * self.sizes[0][1])

    def _snake(self, budget: float):
        world, ov = self.world, self.overhead
        rank, load = 0, 0.0
        bounds = []
        for h, w in self.sizes:
            per_row = w / self._unit()
            rows_of = [0] * world
            row = 0
            while row < h:
                remaining = h - row
                if rank == world - 1:
                    take = remaining
                else:
                    cap = int(max(budget - load - ov, 0.0) / per_row) // ALIGN * ALIGN
                    if cap >= remaining:
                        take = remaining
                    elif cap < MIN_BAND_ROWS:
                        if load > 0.0:                 # no room left for a worthwhile band: the next rank goes on
                            rank, load = rank + 1, 0.0
                            continue
                        take = min(remaining, MIN_BAND_ROWS)
                    else:
                        take = cap
                    if 0 < remaining - take < MIN_BAND_ROWS:      # never leave a sliver behind
                        take = remaining - MIN_BAND_ROWS if remaining - MIN_BAND_ROWS >= MIN_BAND_ROWS else remaining
                rows_of[rank] += take
                load += take * per_row + ov
                row += take
                if row < h:                            # this rank is full; the level continues on the next one
                    rank, load = rank + 1, 0.0
            edges = [0]
            for r in range(world):
                edges.append(edges[-1] + rows_of[r])
            bounds.append(edges)
        return bounds

    def loads(self, bounds=None):
        """Modelled cost per rank (top level = 1)."""
        bounds = self.bounds if bounds is None else bounds
        out = [0.0] * self.world
        for (h, w), edges in zip(self.sizes, bounds):
            for r in range(self.world):
                rows = edges[r + 1] - edges[r]
                if rows:
                    out[r] += rows * w / self._unit() + self.overhead
        return out

    def _search(self):
        work = sum(h * w for h, w in self.sizes) / self._unit()
        lo = work / self.world
        hi = work + self.overhead * len(self.sizes)
        best = None
        steps = 400
        for i in range(steps + 1):
            budget = lo * (hi / lo) ** (i / steps) if hi > lo else lo
            bounds = self._snake(budget)
            n_bands = sum(1 for edges in bounds for r in range(self.world) if edges[r + 1] > edges[r])
            key = (round(max(self.loads(bounds)), 9), n_bands)
            if best is None or key < best[0]:
                best = (key, bounds)
        if all(BandPlan.shardable(h, self.world) for h, _ in self.sizes):
            # tiny pyramids, where the 16-row granularity dominates: equal bands of every level can still win
            bounds = [[r * (h // self.world) for r in range(self.world + 1)] for h, _ in self.sizes]
            key = (round(max(self.loads(bounds)), 9), len(self.sizes) * self.world)
            if key < best[0]:
                best = (key, bounds)
        return best[1]

    # -- queries --------------------------------------------------------------------------------------------
    def band(self, level: int, rank: int):
        """(first owned row, one past the last owned row) of `rank` at `level`; equal when it owns nothing."""
        return self.bounds[level][rank], self.bounds[level][rank + 1]

    def neighbours(self, level: int, rank: int):
        """(rank owning the rows above this rank's band, rank owning the rows below), None at the image border or
        when this rank has no band at the level."""
        edges = self.bounds[level]
        if edges[rank + 1] == edges[rank]:
            return None, None
        up = next((r for r in range(rank - 1, -1, -1) if edges[r + 1] > edges[r]), None)
        dn = next((r for r in range(rank + 1, self.world) if edges[r + 1] > edges[r]), None)
        return up, dn

    def halo_depth(self, wanted: Optional[int] = None) -> int:
        """Rows of halo every band can swap per exchange: `wanted` (default HALO_DEPTH) when every band owns at least
        that many rows at the deepest layer of the feature path (stride ALIGN), else 1.  The same on every rank."""
        wanted = HALO_DEPTH if wanted is None else int(wanted)
        rows = [e[r + 1] - e[r] for e in self.bounds for r in range(self.world) if e[r + 1] > e[r]]
        return wanted if wanted <= 1 or min(rows) // ALIGN >= wanted else 1

    def describe(self) -> str:
        parts = []
        for li, ((h, w), edges) in enumerate(zip(self.sizes, self.bounds)):
            owners = [f'r{r}:{edges[r + 1] - edges[r]}' for r in range(self.world) if edges[r + 1] > edges[r]]
            parts.append(f'{h}x{w}[' + ' '.join(owners) + ']')
        return ' '.join(parts)


def pack_layout(channels: Sequence[int]):
    """Offsets of the raw Grams in the all-reduced buffer; the content SSE sits in the last slot."""
    offs, o = [], 0
    for c in channels:
        offs.append(o)
        o += c * c
    return offs, o, o + 1     # gram offsets, index of the content slot, total floats


class TorchDistGroup:
    """The collective used in production: torch.distributed (NCCL) on the current stream."""

    def __init__(self):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def all_reduce_sum(self, t: torch.Tensor) -> None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def exchange(self, sends, recvs) -> None:
        """One grouped point-to-point step: sends / recvs are lists of (contiguous tensor, peer rank).  Over NCCL
        this is a single ncclGroup (one kernel moving every row over NVLink); the wait orders the current stream
        after it without blocking the host."""
        p2p = [dist.P2POp(dist.isend, t, peer) for t, peer in sends] + \
              [dist.P2POp(dist.irecv, t, peer) for t, peer in recvs]
        if not p2p:
            return
        for work in dist.batch_isend_irecv(p2p):
            work.wait()


class PeerHaloGroup(TorchDistGroup):
    """EXPERIMENTAL (AST_HALO=peer; the kernel passes its one-GPU loop-back test, this class has NOT yet run between
    two processes — round 2 validates it at 2 GPUs under a timeout before it may become the default).  The halo rows of one lock-step step travel through
    NVLink peer memory in ONE launch of ast_halo_exchange (csrc/halo.cu) instead of a grouped NCCL send/recv
    (~35 us per step, 25 steps per closure = 0.9 ms of a 5.4 ms step at 8 GPUs): every rank owns a symmetric
    buffer (torch.distributed._symmetric_memory) with a pair of staging slots and an arrival counter per
    (pyramid level, direction); a neighbour pushes its edge row into my slot, bumps my counter, and the same kernel
    copies the staged row into my halo.  The all-reduces stay on NCCL."""

    MAX_LEVELS = 8

    def __init__(self):
        super().__init__()
        self._hdl = None

    def prepare(self, device: torch.device, row_bytes: int) -> None:
        """Collective: (re)allocate the symmetric buffer for rows of up to row_bytes and reset every counter."""
        import torch.distributed._symmetric_memory as symm_mem
        slot = (int(row_bytes) + 255) // 256 * 256
        n_entries = 2 * self.MAX_LEVELS                       # (level, from-above / from-below)
        flags_bytes = 256 * n_entries                         # one counter per 256-byte line
        need = flags_bytes + n_entries * 2 * slot
        if self._hdl is None or self._slot < slot:
            enable = getattr(symm_mem, 'enable_symm_mem_for_group', None)
            if enable is not None:
                try:
                    enable(dist.group.WORLD.group_name)
                except Exception:
                    pass
            self._buf = symm_mem.empty(need, dtype=torch.uint8, device=device)
            self._hdl = symm_mem.rendezvous(self._buf, dist.group.WORLD)
            self._ptrs = [int(p) for p in self._hdl.buffer_ptrs]
            self._slot, self._flags_bytes = slot, flags_bytes
            self._state = torch.zeros(n_entries * 64, dtype=torch.int32, device=device)   # 256 B per entry
        torch.cuda.synchronize(device)
        dist.barrier()                                        # nobody is still exchanging with the old counters
        self._buf.zero_()
        self._state.zero_()
        torch.cuda.synchronize(device)
        dist.barrier()

    def _entry(self, level: int, direction: int) -> int:
        if not 0 <= level < self.MAX_LEVELS:
            raise ValueError(f'peer halo exchange supports {self.MAX_LEVELS} pyramid levels; got level {level}')
        return 2 * level + direction

    def exchange_rows(self, entries, zero_border: bool, depth: int = 1) -> None:
        from . import _lib as L, ops
        if self._hdl is None:
            raise RuntimeError('PeerHaloGroup.prepare() has not run (parallel.maybe_shard calls it)')
        rows = []

        def add(src, halo, peer, mine, theirs):
            """mine: my entry (the slot pair the neighbour fills); theirs: the neighbour's entry that I fill."""
            nbytes = halo.numel() * 4
            if nbytes > self._slot:
                raise ValueError(f'halo row of {nbytes} bytes exceeds the staging slot ({self._slot})')
            r = L.HaloRow()
            r.halo, r.bytes, r.slot_stride = halo.data_ptr(), nbytes, self._slot
            if src is not None:
                r.src = src.data_ptr()
                r.dst_remote = self._ptrs[peer] + self._flags_bytes + theirs * 2 * self._slot
                r.flag_remote = self._ptrs[peer] + 256 * theirs
                r.stage = self._ptrs[self.rank] + self._flags_bytes + mine * 2 * self._slot
                r.flag_local = self._ptrs[self.rank] + 256 * mine
                r.state = self._state.data_ptr() + 256 * mine
            rows.append(r)

        dev = None
        d = int(depth)
        for r, up, dn, level in entries:
            dev = r.device
            h = r.shape[0] - 2 * d
            if h < d:
                raise ValueError(f'a band of {h} rows cannot hand {d} rows to its neighbours')
            # my "from above" entry (direction 0) pairs with the upper neighbour's "from below" entry (1); the `d` edge
            # rows are one contiguous piece
            if up is not None:
                add(r[d:2 * d], r[0:d], up, self._entry(level, 0), self._entry(level, 1))
            elif zero_border:
                add(None, r[0:d], None, 0, 0)
            if dn is not None:
                add(r[h:h + d], r[h + d:h + 2 * d], dn, self._entry(level, 1), self._entry(level, 0))
            elif zero_border:
                add(None, r[h + d:h + 2 * d], None, 0, 0)
        if not rows:
            return
        if len(rows) > L.AST_HALO_MAX_ROWS:
            # never split a step over two launches: neighbours could then wait on each other's second launch
            raise ValueError(f'{len(rows)} halo rows in one step; ast_halo_exchange takes {L.AST_HALO_MAX_ROWS}')
        arr = (L.HaloRow * len(rows))(*rows)
        ops._launch(dev, ('halo_exchange', len(rows)), 'ast_halo_exchange', arr, len(rows))


def grad_gather_layout(sizes, bands):
    """Pure host logic of PeerGradGather: byte offsets of the levels' (3, H, W) fp32 gradients in the symmetric buffer
    (256-byte aligned), the offset of the flag area, and this rank's segments [(offset, bytes)] — one per plane of every
    level it owns rows [r0, r1) of."""
    offs, o = [], 0
    for h, w in sizes:
        offs.append(o)
        o += (3 * h * w * 4 + 255) // 256 * 256
    segs = []
    for i, ((h, w), (r0, r1)) in enumerate(zip(sizes, bands)):
        if r1 <= r0:
            continue
        for c in range(3):
            off, nbytes = offs[i] + (c * h + r0) * w * 4, (r1 - r0) * w * 4
            if off % 16 or nbytes % 16:
                raise ValueError('gradient rows must be 16-byte aligned (level widths multiples of 4)')
            segs.append((off, nbytes))
    return offs, o, segs


class PeerGradGather:
    """The image-gradient pyramid of a row-band sharded job in NVLink peer memory (ast_band_gather, csrc/halo.cu).

    Every rank keeps one gradient tensor per pyramid level in a SYMMETRIC buffer (torch symmetric memory: same layout
    on every rank, peer-mapped), writes the rows it owns there, and `gather()` stores them into the same place of every
    peer's buffer and waits for theirs: the rows of different ranks are disjoint, so no reduction — and no 75 MB NCCL
    all-reduce — is needed; afterwards all ranks hold bit-identical gradients.  `announce()` at the start of a closure
    tells the peers that this rank's buffer may be overwritten (its previous gradient has been consumed)."""

    def __init__(self, group, device: torch.device, sizes, bands):
        """sizes: [(H, W)] per level; bands: [(r0, r1)] rows this rank owns per level.  Collective."""
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        self.rank, self.world = group.rank, group.world
        if self.world - 1 > L.AST_GATHER_MAX_PEERS:
            raise ValueError(f'peer-memory gradient gather supports {L.AST_GATHER_MAX_PEERS + 1} ranks')
        offs, flags_off, segs = grad_gather_layout(sizes, bands)
        if len(segs) > L.AST_GATHER_MAX_SEGS:
            raise ValueError('too many gradient segments for ast_band_gather')
        total = flags_off + 256 * 2 * self.world
        enable = getattr(symm_mem, 'enable_symm_mem_for_group', None)
        if enable is not None:
            try:
                enable(dist.group.WORLD.group_name)
            except Exception:
                pass
        self._buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self._hdl = symm_mem.rendezvous(self._buf, dist.group.WORLD)
        ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        self._buf.zero_()
        self._state = torch.zeros(32, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier()
        self.views = [self._buf[offs[i]:offs[i] + 3 * h * w * 4].view(torch.float32).view(1, 3, h, w)
                      for i, (h, w) in enumerate(sizes)]
        d = L.BandGatherDesc()
        d.local_base = ptrs[self.rank]
        peers = [r for r in range(self.world) if r != self.rank]
        for k, r in enumerate(peers):
            d.peer_base[k] = ptrs[r]
            # ready[src] at flags + 256 * src, arrive[src] at flags + 256 * (world + src) of the OWNER's buffer
            d.ready_remote[k] = ptrs[r] + flags_off + 256 * self.rank
            d.ready_local[k] = ptrs[self.rank] + flags_off + 256 * r
            d.arrive_remote[k] = ptrs[r] + flags_off + 256 * (self.world + self.rank)
            d.arrive_local[k] = ptrs[self.rank] + flags_off + 256 * (self.world + r)
        d.state = self._state.data_ptr()
        d.n_peers = len(peers)
        for n, (off, nbytes) in enumerate(segs):
            d.seg_off[n], d.seg_bytes[n] = off, nbytes
        d.n_segs = len(segs)
        self._desc = d
        self._device = device

    def announce(self) -> None:
        from . import ops
        ops._launch(self._device, ('band_announce',), 'ast_band_announce', self._desc)

    def gather(self) -> None:
        from . import ops
        ops._launch(self._device, ('band_gather', self._desc.n_segs), 'ast_band_gather', self._desc)


def halo_schedule(kinds: Sequence[str], taps_at, depth: int):
    """Pure host logic: WHERE the lock-step feature path swaps halo rows when a band carries `depth` halo rows.

    kinds: 'conv' / 'pool' per step of the feature path; taps_at: the steps that have a loss tap; depth: 1 or 2.
    A 3x3 convolution over a band whose `v` innermost halo rows per side are valid leaves v - 1 valid halo rows
    (the band's outer output rows saw its zero padding instead of real rows), a 2x2 max-pool of the owned rows
    leaves none; the image band itself comes from the replicated image with all `depth` rows.  An exchange restores
    `depth`.  So with depth 2 only every second convolution of a block needs its neighbours — 6 forward exchanges
    for VGG19 up to conv5_1 instead of 12.  The backward is the mirror image on the gradient w.r.t. each
    convolution's output: backward-data over a band with u valid gradient halo rows leaves u - 1; the steps that run
    without an exchange apply their loss-tap gradients and ReLU masks on the owned rows +- 1 (`ext`), reading the
    forward band's halo row, which is valid there because that band fed a convolution.  7 instead of 13.

    Returns (forward: set of steps with an exchange BEFORE them, backward: {conv step: ('exchange', 0) |
    ('local', ext)} for every convolution at or below the deepest tap)."""
    depth = int(depth)
    if depth not in (1, 2):
        raise ValueError(f'halo depth {depth}: 1 or 2 rows')
    fwd, v = set(), depth
    for s, kind in enumerate(kinds):
        if kind == 'conv':
            if v < 1:
                fwd.add(s)
                v = depth
            v -= 1
        else:
            v = 0
    bwd, u, flowing = {}, 0, False
    for s in range(len(kinds) - 1, -1, -1):
        if not (flowing or s in taps_at):
            continue
        flowing = True
        if kinds[s] == 'conv':
            if u < 1:
                bwd[s] = ('exchange', 0)
                u = depth
            else:
                bwd[s] = ('local', u)
            u -= 1
        else:
            u = 0
    return fwd, bwd


def halo_exchange(group, rows, zero_border: bool = False, depth: int = 1) -> None:
    """rows: one (h + 2 depth, w, C) contiguous view of a padded band, or a list of them (the first and last `depth`
    rows are the halos; depth rows are contiguous in NHWC, so they travel as one piece); a list entry may also be (view, up, dn[, level]) naming the ranks that own the rows above / below this
    band (None at the image border) — the default is rank - 1 / rank + 1, the plan of equal bands — and the index
    of the pyramid level (only the peer-memory group needs it: one staging slot pair per level and direction).  Sends the first /
    last owned row of every band to its neighbours and receives their edge rows into the halos, all in ONE grouped
    exchange (called with an empty list by a rank that has no band at this step, so emulated collectives stay in
    step; over NCCL that is a no-op).

    Forward (activations, zero_border=False): halos at the image border are left alone — they hold the
    convolution's zero padding and nothing ever writes them.
    Backward (gradients w.r.t. a convolution's output, zero_border=True): the same exchange makes every owned row
    of the following backward-data convolution complete (it needs gradient rows i-1..i+1); there is no gradient
    row outside the image, so a border halo is zeroed (gradient buffers are recycled, unlike activation bands)."""
    if torch.is_tensor(rows):
        rows = [rows]
    entries = []
    for n, entry in enumerate(rows):
        if torch.is_tensor(entry):
            entries.append((entry, group.rank - 1 if group.rank > 0 else None,
                            group.rank + 1 if group.rank + 1 < group.world else None, n))
        else:
            entries.append(tuple(entry) if len(entry) == 4 else (*entry, n))
    d = int(depth)
    if hasattr(group, 'exchange_rows'):          # one peer-memory kernel for the whole step (PeerHaloGroup)
        group.exchange_rows(entries, zero_border, d)
        return
    sends, recvs = [], []
    for r, up, dn, _ in entries:
        h = r.shape[0] - 2 * d
        if h < d:
            raise ValueError(f'a band of {h} rows cannot hand {d} rows to its neighbours')
        if up is not None:
            sends.append((r[d:2 * d] if d > 1 else r[1], up))
            recvs.append((r[0:d] if d > 1 else r[0], up))
        elif zero_border:
            r[0:d].zero_()
        if dn is not None:
            sends.append((r[h:h + d] if d > 1 else r[h], dn))
            recvs.append((r[h + d:h + 2 * d] if d > 1 else r[h + 1], dn))
        elif zero_border:
            r[h + d:].zero_()
    group.exchange(sends, recvs)


_GROUP = None     # set by init_sharding(); None -> single-process path


def world():
    if _GROUP is not None:
        return _GROUP.rank, _GROUP.world
    return 0, 1


def init_sharding(group=None) -> None:
    """Enable row-band sharding for subsequent NeuralStyleTransfer.process calls in this process.
    group: object with rank, world and all_reduce_sum(tensor) (default: the initialized torch.distributed group)."""
    global _GROUP
    if group is None:
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError('init_sharding() needs an initialized torch.distributed process group')
        # peer memory needs CUDA ranks of one node: any other backend (the gloo host-logic tests) uses send/recv
        peer = os.environ.get('AST_HALO', DEFAULT_HALO) == 'peer' and 'nccl' in str(dist.get_backend()).lower()
        group = PeerHaloGroup() if peer else TorchDistGroup()
    _GROUP = group


def disable_sharding() -> None:
    global _GROUP
    _GROUP = None


class ShardedLevel:
    """Replaces LossBuilder.build for one level when sharding is on (same return contract)."""

    def __init__(self, group, neural_net, content_idx: int, style_idx: List[int], target_content: torch.Tensor,
                 target_grams: List[torch.Tensor], weights, height: int, width: int, precision: int):
        from . import ops
        self.ops = ops
        self.group = group
        self.net = neural_net
        self.cidx, self.sidx = content_idx, list(style_idx)
        self.weights = tuple(float(w) for w in weights)
        self.H, self.W = height, width
        self.plan = BandPlan(height, group.rank, group.world)
        self.precision = precision
        self.target_grams = target_grams
        cs = LAYER_STRIDE[content_idx]
        ga, gb = self.plan.global_feat_rows(cs)
        self.target_content_band = target_content[:, ga:gb, :].contiguous()
        self.content_numel_global = target_content.numel()
        self.channels = [g.shape[-1] for g in target_grams]
        self.offs, self.content_slot, self.n_packed = pack_layout(self.channels)
        self.wss = ops.LevelWorkspaces()
        self.fin_ws = [None] * len(self.sidx)

    def build(self, level_img: torch.Tensor):
        x = level_img[:, :, self.plan.lo:self.plan.hi, :]
        feats = self.net(x)
        return ShardLevelLossFn.apply(self, level_img, feats[self.cidx], *[feats[k] for k in self.sidx])


class ShardLevelLossFn(torch.autograd.Function):
    """inputs: full level image (TV), band content map, band style maps.  outputs as LevelLossFn."""

    @staticmethod
    def forward(ctx, sh: ShardedLevel, level_img, content_feat, *style_feats):
        ops = sh.ops
        dev = ops._require_cuda(level_img, content_feat, *style_feats)
        n_style = len(style_feats)
        plan = sh.plan
        cw, sw, tvw = sh.weights
        level_img = level_img.contiguous()
        style_feats = [f.contiguous() for f in style_feats]
        packed = torch.zeros(sh.n_packed, dtype=torch.float32, device=dev)
        geo = []
        for k, f in enumerate(style_feats):
            ch, hb, w = f.shape[-3], f.shape[-2], f.shape[-1]
            a, b, rows = plan.feat_rows(LAYER_STRIDE[sh.sidx[k]])
            assert rows == hb and ch == sh.channels[k], (rows, hb, ch)
            hw_band, ld, off = (b - a) * w, hb * w, a * w
            raw = packed[sh.offs[k]:sh.offs[k] + ch * ch]
            ops.gram_mse_fwd(f, ch, hw_band, 1.0, None, raw, None, sh.wss.for_gram(k, ch, hw_band, dev), sh.precision,
                             ld=ld, offset=off)
            geo.append((ch, hw_band, ld, off, (sh.H // LAYER_STRIDE[sh.sidx[k]]) * (sh.W // LAYER_STRIDE[sh.sidx[k]])))
        ca, cb, _ = plan.feat_rows(LAYER_STRIDE[sh.cidx])
        x_band = content_feat[0, :, ca:cb, :].contiguous()
        ops.mse_fwd(x_band, sh.target_content_band, 1.0, packed[sh.content_slot], sh.wss.for_reduce('content', dev))
        sh.group.all_reduce_sum(packed)                      # the one exchange step of the level
        vals = torch.empty(n_style + 2, dtype=torch.float32, device=dev)
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        ds = []
        for k in range(n_style):
            ch, hw_global = geo[k][0], geo[k][4]
            d = torch.empty((ch, ch), dtype=torch.float32, device=dev)
            if sh.fin_ws[k] is None or sh.fin_ws[k].buf.device != dev:
                sh.fin_ws[k] = ops.reduce_workspace(dev)
            ops.gram_finalize(packed[sh.offs[k]:sh.offs[k] + ch * ch], ch, 1.0 / (ch * hw_global), sh.target_grams[k],
                              d, vals[k], sh.fin_ws[k])
            ds.append(d)
        torch.mul(packed[sh.content_slot], 1.0 / sh.content_numel_global, out=vals[n_style])
        sums2 = torch.empty(2, dtype=torch.float32, device=dev)
        ops.tv_fwd(level_img, sums2, vals[n_style + 1], sh.wss.for_reduce('tv', dev))
        ops._launch(dev, ('combine',), 'ast_level_combine', vals.data_ptr(), n_style, vals[n_style].data_ptr(),
                    vals[n_style + 1].data_ptr(), cw, sw, tvw, out4.data_ptr())
        ctx.save_for_backward(level_img, x_band, sums2, *style_feats, *ds)
        ctx.sh, ctx.geo, ctx.n_style = sh, geo, n_style
        ctx.content_shape, ctx.content_rows = content_feat.shape, (ca, cb)
        total, content, style, tv = out4[0], out4[1], out4[2], out4[3]
        ctx.mark_non_differentiable(content, style, tv)
        return total, content, style, tv

    @staticmethod
    def backward(ctx, g_total, g_content, g_style, g_tv):
        sh = ctx.sh
        ops = sh.ops
        saved = ctx.saved_tensors
        level_img, x_band, sums2 = saved[:3]
        n = ctx.n_style
        style_feats, ds = saved[3:3 + n], saved[3 + n:3 + 2 * n]
        cw, sw, tvw = sh.weights
        dev = level_img.device
        g = ops._gscale(g_total, dev)
        need = ctx.needs_input_grad          # (sh, level_img, content_feat, *style_feats)
        d_img = d_content = None
        if need[1] and sh.group.rank == 0:   # TV is replicated: count its gradient once
            d_img = torch.empty_like(level_img)
            ops.tv_bwd(level_img, sums2, tvw, g, d_img, False)
        if need[2]:
            ca, cb = ctx.content_rows
            d_content = torch.zeros(ctx.content_shape, dtype=torch.float32, device=dev)
            band = torch.empty_like(x_band)
            ops.mse_bwd(x_band, sh.target_content_band, cw * 2.0 / sh.content_numel_global, g, band, False)
            d_content[0, :, ca:cb, :] = band
        d_style = []
        for k in range(n):
            if not need[3 + k]:
                d_style.append(None)
                continue
            f = style_feats[k]
            ch, hw_band, ld, off, hw_global = ctx.geo[k]
            w = f.shape[-1]
            a, b = off // w, (off + hw_band) // w
            df = torch.empty_like(f)
            if a > 0:
                df[..., :a, :].zero_()
            if b < f.shape[-2]:
                df[..., b:, :].zero_()
            ops.gram_bwd(ds[k], f, ch, hw_band, (sw / n) * 4.0 / (float(ch) * ch * ch * hw_global), g, df, False,
                         sh.precision, ld=ld, offset=off)
            d_style.append(df)
        return (None, d_img, d_content, *d_style)


PLAN = None       # the PyramidBands of the most recent maybe_shard() (None: equal bands or nothing sharded)


def maybe_shard(loss_builders, optimizing_img, neural_net, content_idx, style_idx, weights) -> int:
    """Hook called by NeuralStyleTransfer.process after the per-level LossBuilders exist.  Returns the number of
    levels that were sharded (0 on the single-process path).

    When every level can take the halo-exchange path the rows of the WHOLE pyramid are dealt out by PyramidBands
    (level-aware: a rank owns rows of one or two levels, AST_BANDS=uniform restores equal bands of every level);
    otherwise each shardable level is cut into `world` equal bands as before."""
    global PLAN, _GROUP
    PLAN = None
    if _GROUP is None or _GROUP.world == 1 and os.environ.get('AST_SHARD_SINGLE', '0') != '1':
        return 0
    from . import ops, neural_style_transfer as nst
    from .sharded_path import ShardedPathLevel
    h, w = optimizing_img.shape[-2], optimizing_img.shape[-1]
    rank, world_ = _GROUP.rank, _GROUP.world
    sizes = [(h >> i, w >> i) for i in range(len(loss_builders))]
    plans = [lb.path_plan(optimizing_img) for lb in loss_builders]
    uniform = os.environ.get('AST_BANDS', 'pyramid') == 'uniform'
    whole = all(p is not None for p in plans) and all(
        lh % ALIGN == 0 and lw % ALIGN == 0 and lh >= MIN_BAND_ROWS and h % (1 << i) == 0 and w % (1 << i) == 0
        for i, (lh, lw) in enumerate(sizes))
    if whole and not (uniform and not all(BandPlan.shardable(lh, world_) for lh, _ in sizes)):
        PLAN = PyramidBands(sizes, world_, uniform=uniform)
        depth = PLAN.halo_depth()
        if hasattr(_GROUP, 'prepare'):           # peer-memory halo exchange: symmetric staging for the widest row
            err = None
            try:
                _GROUP.prepare(optimizing_img.device, depth * 4 * 64 * sizes[0][1])
            except Exception as e:               # no symmetric memory on this box
                err = e
            if not all_ranks_ok(err is None, optimizing_img.device):
                import warnings
                warnings.warn(f'peer-memory halo exchange unavailable ({err!r}); using grouped NCCL send/recv')
                _GROUP = TorchDistGroup()
        for i, lb in enumerate(loss_builders):
            r0, r1 = PLAN.band(i, rank)
            up, dn = PLAN.neighbours(i, rank)
            lb.shard = ShardedPathLevel(_GROUP, plans[i], lb.target_images[0], lb.target_images[1], content_idx,
                                        style_idx, weights, sizes[i][0], sizes[i][1], band=(r0, r1, up, dn),
                                        halo_depth=depth)
        return len(loss_builders)
    n = 0
    for i, lb in enumerate(loss_builders):
        lh, lw = sizes[i]
        ok = BandPlan.shardable(lh, world_) and lw % ALIGN == 0 and (h % (1 << i) == 0)
        plan = plans[i] if ok else None
        if plan is not None:
            lb.shard = ShardedPathLevel(_GROUP, plan, lb.target_images[0], lb.target_images[1], content_idx,
                                        style_idx, weights, lh, lw)
            n += 1
        elif ok:
            grams = [g[0].detach().contiguous() for g in lb.target_style_representation]
            lb.shard = ShardedLevel(_GROUP, neural_net, content_idx, style_idx, lb.target_content_representation,
                                    grams, weights, lh, lw, ops._prec(nst.PRECISION))
            n += 1
        else:
            lb.replicated_rank0_only = _GROUP.rank != 0
    return n


def lockstep_pyramid(loss_builders):
    """ShardedPyramid over the job's levels when every one of them is a halo-exchange level, else None."""
    from .sharded_path import ShardedPathLevel, ShardedPyramid
    shards = [getattr(lb, 'shard', None) for lb in loss_builders]
    if shards and all(isinstance(s, ShardedPathLevel) for s in shards):
        return ShardedPyramid(shards)
    return None


def sync_image_grad(optimizing_img, already_global: bool = False) -> None:
    """All-reduce(sum) of the image gradient so that every rank steps an identical optimizer.  already_global: the
    closure gathered the gradient itself (PeerGradGather): nothing to do."""
    if _GROUP is None or _GROUP.world == 1 or already_global:
        return
    if optimizing_img.grad is None:          # every rank must join the collective
        optimizing_img.grad = torch.zeros_like(optimizing_img)
    from . import ops
    with ops.timed(optimizing_img.device, ('allreduce_image_grad', optimizing_img.numel())):
        _GROUP.all_reduce_sum(optimizing_img.grad)
