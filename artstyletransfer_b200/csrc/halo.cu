// Halo-row exchange between row-band neighbours through NVLink peer memory (ast_halo_exchange).
//
// One launch replaces the grouped NCCL send/recv of one lock-step step (parallel.halo_exchange): every entry is a
// symmetric exchange with ONE neighbour — push my edge row into the neighbour's staging slot with plain 16-byte
// stores over NVLink, publish an arrival counter with a system-scope release store, spin on my own counter with
// acquire loads until the neighbour's row has landed in MY staging slot, copy it into my halo row.  Entries with
// src == NULL zero a halo row instead (gradient bands at the image border).
//
// Protocol per entry (all state is per (level, direction), so entries never interfere):
//   count      local, exchanges completed so far; seq = count + 1 identifies this exchange on BOTH sides (neighbours
//              run the same number of exchanges), so nothing per-launch has to be passed in and a CUDA-graph replay
//              advances by itself;
//   staging    two slots per entry, selected by seq & 1.  A slot is overwritten at exchange e + 2, which the writer
//              reaches only after it has seen the reader's counter for e + 1, which the reader publishes in a later
//              launch than the one that drained the slot at e (stream order): no write-after-read hazard;
//   tickets    two self-resetting counters find the last CTA of a row: the one that publishes the arrival counter
//              after every CTA's stores, and the one that advances `count` after every CTA has read it.
// The CTAs of one launch must be co-resident (they spin on the neighbour, whose progress needs our stores):
// grid = (1..32) x n_rows <= 512 CTAs of 512 threads (4 per SM: all resident).  A spin that lasts ~4 s traps.
#include "ast_common.cuh"

namespace ast {

struct HaloArgs {
  ast_halo_row rows[AST_HALO_MAX_ROWS];
};

constexpr int kHaloCtasMax = 32;      // CTAs per row: ~32 KB each (a 786 KB relu1_x row of the 2048x3072 level takes 24);
constexpr int kHaloThreads = 512;     // 16 rows x 32 CTAs x 512 threads still fit the machine at once (they spin)

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kHaloThreads) halo_exchange_kernel(const __grid_constant__ HaloArgs A) {
  const ast_halo_row& r = A.rows[blockIdx.y];
  const int64_t n16 = r.bytes >> 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint4* halo = reinterpret_cast<uint4*>(r.halo);
  if (r.src == nullptr) {
    for (int64_t i = i0; i < n16; i += stride) halo[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  __shared__ uint32_t s_seq;
  if (threadIdx.x == 0) s_seq = reinterpret_cast<volatile uint32_t*>(r.state)[0] + 1u;
  __syncthreads();
  const uint32_t seq = s_seq;
  const int64_t slot = (int64_t)(seq & 1u) * r.slot_stride;

  // ---- push my edge row into the neighbour's staging slot (loads batched so four are in flight per thread)
  const uint4* src = reinterpret_cast<const uint4*>(r.src);
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(r.dst_remote) + slot);
  int64_t i = i0;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
    dst[i] = a;
    dst[i + stride] = b;
    dst[i + 2 * stride] = c;
    dst[i + 3 * stride] = d;
  }
  for (; i < n16; i += stride) dst[i] = src[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                                   // this CTA's stores before its ticket
    const uint32_t t = atomicAdd(&r.state[1], 1u);
    if (t == gridDim.x - 1) {                                 // every CTA of the row has stored
      r.state[1] = 0u;
      __threadfence_system();
      st_release_sys(r.flag_remote, seq);
    }
    // ---- wait until the neighbour's row for this exchange has landed in my staging slot
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(r.flag_local) - seq) < 0) {
      if (clock64() - t0 > 8000000000LL) __trap();            // ~4 s: the neighbour is gone
    }
  }
  __syncthreads();
  const uint4* stg = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(r.stage) + slot);
  for (int64_t k = i0; k < n16; k += stride) halo[k] = __ldcg(stg + k);     // L2 only: the peer wrote it
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(&r.state[2], 1u);
    if (t == gridDim.x - 1) {                                 // every CTA has read `count`: advance it
      r.state[2] = 0u;
      __threadfence();
      reinterpret_cast<volatile uint32_t*>(r.state)[0] = seq;
    }
  }
}

}  // namespace ast

extern "C" int ast_halo_exchange(const ast_halo_row* rows, int n_rows, void* stream) {
  using namespace ast;
  AST_REQUIRE(rows != nullptr && n_rows > 0 && n_rows <= AST_HALO_MAX_ROWS, AST_ERR_INVALID,
              "ast_halo_exchange: n_rows must be 1..%d (got %d)", AST_HALO_MAX_ROWS, n_rows);
  HaloArgs args = {};
  for (int k = 0; k < n_rows; ++k) {
    const ast_halo_row& r = rows[k];
    AST_REQUIRE(r.halo != nullptr && r.bytes > 0 && (r.bytes & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(r.halo) & 15) == 0,
                AST_ERR_INVALID, "ast_halo_exchange: row %d: halo must be 16-byte aligned, bytes a multiple of 16", k);
    if (r.src != nullptr) {
      AST_REQUIRE(r.dst_remote && r.stage && r.flag_remote && r.flag_local && r.state && (r.slot_stride & 15) == 0 &&
                      r.slot_stride >= r.bytes && (reinterpret_cast<uintptr_t>(r.src) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(r.dst_remote) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(r.stage) & 15) == 0,
                  AST_ERR_INVALID, "ast_halo_exchange: row %d: null or misaligned exchange pointers", k);
    }
    args.rows[k] = r;
  }
  int64_t max_bytes = 0;
  for (int k = 0; k < n_rows; ++k) max_bytes = rows[k].bytes > max_bytes ? rows[k].bytes : max_bytes;
  int ctas = (int)((max_bytes + 32767) / 32768);
  ctas = ctas < 1 ? 1 : (ctas > kHaloCtasMax ? kHaloCtasMax : ctas);
  halo_exchange_kernel<<<dim3(ctas, n_rows), kHaloThreads, 0, static_cast<cudaStream_t>(stream)>>>(args);
  return check_launch("halo_exchange");
}

// ---------------------------------------------------------------------------------------------------------------------
// Image-gradient all-gather through NVLink peer memory (ast_band_gather / ast_band_announce).
//
// Row-band sharding leaves every rank with the gradient of ITS rows of every pyramid level; the rows of different
// ranks are disjoint, so the "sum over ranks" the replicated optimizer needs is a concatenation, not a reduction.
// Round 1 all-reduced the whole 75.5 MB image gradient anyway (NCCL ring, ~0.4 ms of a 5 ms step at 8 GPUs).  Here every
// rank keeps the gradient pyramid in a SYMMETRIC buffer (same layout on every rank), writes its own rows there, and one
// launch stores those rows into the same place of every peer's buffer (16-byte stores over NVLink / NVSwitch), then
// waits until every peer's rows have landed here.  Afterwards all ranks hold bit-identical gradients.
//
// Protocol (all counters are monotone sequence numbers, never reset; `state` is local device memory):
//   ready[src]   in MY symmetric flags: peer `src` may be written to for exchange e once IT has announced e, i.e. once
//                its previous gradient has been consumed (ast_band_announce, launched at the start of a closure, in stream
//                order after the optimizer update that read the gradient).  I wait for ready_local[p] >= e before I
//                store into peer p: no write-after-read hazard without double buffering, so the gradient keeps ONE
//                address (CUDA-graph replays reuse it).
//   arrive[src]  in MY symmetric flags: src's rows for exchange e have landed (release store after its last CTA's stores).
// Grid: (kGatherCtas, n_peers); the CTAs of one launch must be co-resident (they wait on peers): <= 7 x 16 CTAs.
// ---------------------------------------------------------------------------------------------------------------------
namespace ast {

constexpr int kGatherCtas = 16;

struct GatherArgs {
  ast_band_gather_desc d;
};

__global__ void __launch_bounds__(64) band_announce_kernel(const __grid_constant__ GatherArgs A) {
  const ast_band_gather_desc& d = A.d;
  __shared__ uint32_t s_seq;
  if (threadIdx.x == 0) s_seq = reinterpret_cast<volatile uint32_t*>(d.state)[4] + 1u;
  __syncthreads();
  const uint32_t seq = s_seq;
  if ((int)threadIdx.x < d.n_peers) {
    __threadfence_system();                                  // everything before this launch (the gradient's readers) first
    st_release_sys(d.ready_remote[threadIdx.x], seq);
  }
  __syncthreads();
  if (threadIdx.x == 0) reinterpret_cast<volatile uint32_t*>(d.state)[4] = seq;
}

__global__ void __launch_bounds__(kHaloThreads) band_gather_kernel(const __grid_constant__ GatherArgs A) {
  const ast_band_gather_desc& d = A.d;
  const int p = blockIdx.y;                                  // peer this CTA group pushes to
  __shared__ uint32_t s_seq;
  if (threadIdx.x == 0) {
    const uint32_t seq = reinterpret_cast<volatile uint32_t*>(d.state)[0] + 1u;
    // the peer must have consumed its previous gradient before I overwrite my rows in its buffer
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(d.ready_local[p]) - seq) < 0) {
      if (clock64() - t0 > 8000000000LL) __trap();
    }
    s_seq = seq;
  }
  __syncthreads();
  const uint32_t seq = s_seq;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int sgi = 0; sgi < d.n_segs; ++sgi) {
    const uint4* src = reinterpret_cast<const uint4*>(d.local_base + d.seg_off[sgi]);
    uint4* dst = reinterpret_cast<uint4*>(d.peer_base[p] + d.seg_off[sgi]);
    const int64_t n16 = d.seg_bytes[sgi] >> 4;
    int64_t i = i0;
    for (; i + 3 * stride < n16; i += 4 * stride) {
      const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], e = src[i + 3 * stride];
      dst[i] = a;
      dst[i + stride] = b;
      dst[i + 2 * stride] = c;
      dst[i + 3 * stride] = e;
    }
    for (; i < n16; i += stride) dst[i] = src[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    uint32_t* ticket = d.state + 8 + p;                     // one self-resetting ticket per peer group
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0u;
      __threadfence_system();
      st_release_sys(d.arrive_remote[p], seq);              // my rows are in peer p's buffer
    }
    // nobody leaves before every peer's rows are here
    for (int q = 0; q < d.n_peers; ++q) {
      const long long t0 = clock64();
      while ((int32_t)(ld_acquire_sys(d.arrive_local[q]) - seq) < 0) {
        if (clock64() - t0 > 8000000000LL) __trap();
      }
    }
    uint32_t* done = d.state + 1;
    if (atomicAdd(done, 1u) == gridDim.x * gridDim.y - 1) {
      *done = 0u;
      __threadfence();
      reinterpret_cast<volatile uint32_t*>(d.state)[0] = seq;
    }
  }
}

}  // namespace ast

static int check_gather_desc(const ast_band_gather_desc* d, const char* who) {
  using namespace ast;
  AST_REQUIRE(d != nullptr && d->state != nullptr, AST_ERR_INVALID, "%s: null descriptor / state", who);
  AST_REQUIRE(d->n_peers >= 1 && d->n_peers <= AST_GATHER_MAX_PEERS, AST_ERR_INVALID, "%s: n_peers must be 1..%d (got %d)",
              who, AST_GATHER_MAX_PEERS, d->n_peers);
  for (int p = 0; p < d->n_peers; ++p)
    AST_REQUIRE(d->ready_remote[p] && d->ready_local[p] && d->arrive_remote[p] && d->arrive_local[p], AST_ERR_INVALID,
                "%s: null flag pointer for peer %d", who, p);
  return AST_OK;
}

extern "C" int ast_band_announce(const ast_band_gather_desc* desc, void* stream) {
  using namespace ast;
  int rc = check_gather_desc(desc, "ast_band_announce");
  if (rc != AST_OK) return rc;
  GatherArgs a;
  a.d = *desc;
  band_announce_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("band_announce");
}

extern "C" int ast_band_gather(const ast_band_gather_desc* desc, void* stream) {
  using namespace ast;
  int rc = check_gather_desc(desc, "ast_band_gather");
  if (rc != AST_OK) return rc;
  AST_REQUIRE(desc->n_segs >= 0 && desc->n_segs <= AST_GATHER_MAX_SEGS, AST_ERR_INVALID,
              "ast_band_gather: n_segs must be 0..%d (got %d)", AST_GATHER_MAX_SEGS, desc->n_segs);
  AST_REQUIRE(desc->local_base != nullptr && (reinterpret_cast<uintptr_t>(desc->local_base) & 15) == 0, AST_ERR_INVALID,
              "ast_band_gather: local_base must be 16-byte aligned");
  for (int p = 0; p < desc->n_peers; ++p)
    AST_REQUIRE(desc->peer_base[p] && (reinterpret_cast<uintptr_t>(desc->peer_base[p]) & 15) == 0, AST_ERR_INVALID,
                "ast_band_gather: peer base %d null or misaligned", p);
  for (int k = 0; k < desc->n_segs; ++k)
    AST_REQUIRE(desc->seg_off[k] >= 0 && (desc->seg_off[k] & 15) == 0 && desc->seg_bytes[k] > 0 && (desc->seg_bytes[k] & 15) == 0,
                AST_ERR_INVALID, "ast_band_gather: segment %d must be 16-byte aligned (offset %lld, %lld bytes)", k,
                (long long)desc->seg_off[k], (long long)desc->seg_bytes[k]);
  GatherArgs a;
  a.d = *desc;
  band_gather_kernel<<<dim3(kGatherCtas, desc->n_peers), kHaloThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("band_gather");
}
