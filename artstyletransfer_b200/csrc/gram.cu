// Gram-matrix style loss: C-ABI entry points, the split-K finalize (+ fused MSE) kernel and the exact-fp32
// FFMA kernels (AST_PREC_FP32).  The tcgen05/TMA/TMEM kernels (AST_PREC_TF32) live in gram_tc.cu.
//
// Replaces math_utils.gram_matrix (math_utils.py:26-34) + torch.nn.MSELoss (neural_style_transfer.py:100-104)
// forward, and the two bmm per layer autograd runs backward.
//
// Forward = two launches:
//   1. split-K partial sums  P_s = F[:, ks] F[:, ks]^T  (one per CTA, in TMEM or registers) -> workspace
//   2. finalize: out = scale * sum_s P_s - A, loss = mean(out^2); fixed summation order => deterministic.
#include "gram.cuh"

namespace ast {

// ------------------------------------------------------------------------------------------------------
// exact fp32 path, forward: 64x64 output tile per CTA, 4x4 micro-tile per thread, K chunks of 32.
// Thread (tx = t%16, ty = t/16) owns rows ty*4+i and columns tx+16*j.
// ------------------------------------------------------------------------------------------------------
constexpr int SG_T = 64, SG_K = 32, SG_PITCH = SG_K + 1;

__device__ __forceinline__ void load_rows_k32(const float* __restrict__ F, int64_t HW, int row0, int64_t k0,
                                             int64_t kend, bool vec_ok, float (*S)[SG_PITCH]) {
  // 64 rows x 32 k: 512 float4 slots, 256 threads x 2
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int slot = threadIdx.x + it * 256;
    const int r = slot >> 3, kq = slot & 7;
    const int64_t k = k0 + 4 * kq;
    const float* p = F + (size_t)(row0 + r) * HW + k;
    float4 v;
    if (vec_ok && k + 3 < kend) {
      v = __ldg(reinterpret_cast<const float4*>(p));
    } else {
      v.x = (k + 0 < kend) ? __ldg(p + 0) : 0.f;
      v.y = (k + 1 < kend) ? __ldg(p + 1) : 0.f;
      v.z = (k + 2 < kend) ? __ldg(p + 2) : 0.f;
      v.w = (k + 3 < kend) ? __ldg(p + 3) : 0.f;
    }
    S[r][4 * kq + 0] = v.x;
    S[r][4 * kq + 1] = v.y;
    S[r][4 * kq + 2] = v.z;
    S[r][4 * kq + 3] = v.w;
  }
}

__global__ void __launch_bounds__(256) gram_fp32_fwd_kernel(const float* __restrict__ F, int C, int64_t HW,
                                                           int64_t ld, int64_t k_per_split, int vec_ok,
                                                           float* __restrict__ partials) {
  __shared__ float As[SG_T][SG_PITCH];
  __shared__ float Bs[SG_T][SG_PITCH];
  const int ntile = C / SG_T;
  const int ti = blockIdx.x / ntile, tj = blockIdx.x % ntile;
  const int split = blockIdx.y;
  const int64_t kbeg = (int64_t)split * k_per_split;
  const int64_t kend = min(HW, kbeg + k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t k0 = kbeg; k0 < kend; k0 += SG_K) {
    load_rows_k32(F, ld, ti * SG_T, k0, kend, vec_ok, As);
    if (ti != tj) load_rows_k32(F, ld, tj * SG_T, k0, kend, vec_ok, Bs);
    __syncthreads();
    float(*Bp)[SG_PITCH] = (ti != tj) ? Bs : As;
#pragma unroll 8
    for (int kk = 0; kk < SG_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[ty * 4 + i][kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bp[tx + 16 * j][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* P = partials + (size_t)split * C * C;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      P[(size_t)(ti * SG_T + ty * 4 + i) * C + tj * SG_T + tx + 16 * j] = acc[i][j];
}

// exact fp32 path, backward: dF[c, n] (+)= scale * sum_k D[c, k] F[k, n];  64 (c) x 64 (n) tile per CTA.
__global__ void __launch_bounds__(256) gram_fp32_bwd_kernel(const float* __restrict__ D, const float* __restrict__ F,
                                                           int C, int64_t HW, int64_t ld, float scale,
                                                           const float* __restrict__ gscale, float* __restrict__ dF,
                                                           int accumulate, int vec_ok) {
  if (gscale) scale *= __ldg(gscale);
  __shared__ float As[SG_T][SG_PITCH];          // D[c-tile rows][k chunk]
  __shared__ __align__(16) float Bs[SG_K][SG_T];  // F[k chunk][n tile]
  const int64_t n0 = (int64_t)blockIdx.x * SG_T;
  const int c0 = blockIdx.y * SG_T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < C; k0 += SG_K) {
    load_rows_k32(D, C, c0, k0, C, (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(D) & 15u) == 0), As);
    // F chunk: 32 k-rows x 64 n: 512 float4 slots
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int slot = threadIdx.x + it * 256;
      const int kk = slot >> 4, nq = slot & 15;
      const int64_t n = n0 + 4 * nq;
      const float* p = F + (size_t)(k0 + kk) * ld + n;
      float4 v;
      if (vec_ok && n + 3 < HW) {
        v = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        v.x = (n + 0 < HW) ? __ldg(p + 0) : 0.f;
        v.y = (n + 1 < HW) ? __ldg(p + 1) : 0.f;
        v.z = (n + 2 < HW) ? __ldg(p + 2) : 0.f;
        v.w = (n + 3 < HW) ? __ldg(p + 3) : 0.f;
      }
      *reinterpret_cast<float4*>(&Bs[kk][4 * nq]) = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < SG_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[ty * 4 + i][kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx + 16 * j;
      if (n < HW) {
        float* o = dF + (size_t)(c0 + ty * 4 + i) * ld + n;
        const float v = scale * acc[i][j];
        *o = accumulate ? *o + v : v;
      }
    }
}

// ------------------------------------------------------------------------------------------------------
// finalize: out = scale * sum_p partial_p - A (mirrored for off-diagonal tiles); loss = mean(out^2).
// Block = (256/L) float4 element groups x L partial lanes (L = 4, 8 or 16, more lanes for small tiles so that the
// grid still fills the machine); lanes are combined in fixed order.
// ------------------------------------------------------------------------------------------------------
struct FinalizeArgs {
  GramPlan plan;
  const float* partials;
  const float* A;      // nullable
  float* out;
  float* loss;         // nullable
  float scale;
  int symmetric_src;   // 1: partials hold only bi<=bj tiles (mirror them); 0: single full tile
  int lanes;           // partial lanes per element group (4, 8, 16)
  int round_out;       // 1: store `out` rounded to nearest TF32 (it only feeds the backward's tensor-core operand)
};

__device__ __forceinline__ uint32_t bf16x2_rn(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// store 4 consecutive results: fp32, fp32 rounded to TF32 (round_out 1) or bfloat16 (round_out 2; `out` then holds
// C x C bfloat16 — the BF16 backward's D operand)
__device__ __forceinline__ void store_d4(float* out, size_t o, const float d[4], int round_out);
__device__ __forceinline__ void store_d1(float* out, size_t o, float d, int round_out);

__device__ __forceinline__ float tf32_rn(float x) {
  // same rounding as the Gram kernels' operand converter (gram_tc.cu: half an ulp added to the magnitude)
  const uint32_t u = __float_as_uint(x);
  return ((u & 0x7f800000u) == 0x7f800000u) ? x : __uint_as_float((u + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ void store_d4(float* out, size_t o, const float d[4], int round_out) {
  if (round_out == 2) {
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out) + o) = make_uint2(bf16x2_rn(d[0], d[1]), bf16x2_rn(d[2], d[3]));
  } else if (round_out == 1) {
    *reinterpret_cast<float4*>(out + o) = make_float4(tf32_rn(d[0]), tf32_rn(d[1]), tf32_rn(d[2]), tf32_rn(d[3]));
  } else {
    *reinterpret_cast<float4*>(out + o) = make_float4(d[0], d[1], d[2], d[3]);
  }
}
__device__ __forceinline__ void store_d1(float* out, size_t o, float d, int round_out) {
  if (round_out == 2) reinterpret_cast<uint16_t*>(out)[o] = (uint16_t)(bf16x2_rn(d, 0.f) & 0xffffu);
  else out[o] = round_out == 1 ? tf32_rn(d) : d;
}

__global__ void __launch_bounds__(256) gram_finalize_kernel(const __grid_constant__ FinalizeArgs a, ReduceWs* ws) {
  __shared__ float4 lanes[256];
  __shared__ double red[32];
  const GramPlan& pl = a.plan;
  const int TR = pl.TR, C = pl.C;
  const int L = a.lanes, groups = 256 / L;
  const int elems_per_block = groups * 4;
  const int blocks_per_tile = (TR * TR) / elems_per_block;
  const int t = blockIdx.x / blocks_per_tile;
  const int eb = blockIdx.x - t * blocks_per_tile;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  const int e0 = eb * elems_per_block + g * 4;  // first of 4 consecutive elements inside the tile
  const float* base = a.partials + (size_t)pl.part_off[t] * TR * TR + e0;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = lane; p < pl.part_cnt[t]; p += L) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)p * TR * TR));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  lanes[lane * groups + g] = s;
  __syncthreads();
  double sq = 0.0;
  if (lane == 0) {
    float4 tot = lanes[g];
    for (int l = 1; l < L; ++l) {
      const float4 v = lanes[l * groups + g];
      tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
    }
    const int r = e0 / TR, c = e0 - r * TR;
    const int gi = pl.tile_bi[t] * TR + r, gj = pl.tile_bj[t] * TR + c;
    float v[4] = {tot.x * a.scale, tot.y * a.scale, tot.z * a.scale, tot.w * a.scale};
    float d[4];
    const size_t o = (size_t)gi * C + gj;
    if (a.A) {
      const float4 av = __ldg(reinterpret_cast<const float4*>(a.A + o));
      d[0] = v[0] - av.x; d[1] = v[1] - av.y; d[2] = v[2] - av.z; d[3] = v[3] - av.w;
    } else {
      d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; d[3] = v[3];
    }
    store_d4(a.out, o, d, a.round_out);
    sq = (double)d[0] * d[0] + (double)d[1] * d[1] + (double)d[2] * d[2] + (double)d[3] * d[3];
    if (a.symmetric_src && pl.tile_bi[t] != pl.tile_bj[t]) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const size_t om = (size_t)(gj + k) * C + gi;
        const float dm = a.A ? v[k] - __ldg(a.A + om) : v[k];
        store_d1(a.out, om, dm, a.round_out);
        sq += (double)dm * dm;
      }
    }
  }
  if (a.loss) {
    double bs = block_sum(sq, red);
    double total;
    if (grid_reduce_last(ws, &bs, 1, &total, red)) *a.loss = (float)(total / ((double)C * (double)C));
  }
}

// ------------------------------------------------------------------------------------------------------
// batched finalize of RAW Grams (row-band sharding: the all-reduced packed buffer of every pyramid level holds five of
// them): out_i = scale_i * G_i - A_i, loss_i = mean(out_i^2), all items in ONE launch instead of one launch each
// (20 per closure at L = 3; ~6 us apiece on the critical path between the all-reduce and the backward).
// One float4 per thread, 1024 elements per block; the loss of an item is reduced through that item's own ReduceWs
// (fp64 block sums, ticketed last block, fixed order -> bit-identical on every rank and every run).
// ------------------------------------------------------------------------------------------------------
struct FinalizeBatchArgs {
  ast_finalize_item items[AST_FINALIZE_MAX_ITEMS];
  int block_off[AST_FINALIZE_MAX_ITEMS + 1];
  int n;
};

__global__ void __launch_bounds__(256) gram_finalize_batch_kernel(const __grid_constant__ FinalizeBatchArgs a,
                                                                 ReduceWs* ws_all) {
  __shared__ double red[32];
  __shared__ bool is_last;
  int it = 0;
  while (it + 1 < a.n && (int)blockIdx.x >= a.block_off[it + 1]) ++it;
  const ast_finalize_item& q = a.items[it];
  const int b = blockIdx.x - a.block_off[it], nb = a.block_off[it + 1] - a.block_off[it];
  const int64_t n4 = (int64_t)q.C * q.C / 4;
  const int64_t i = (int64_t)b * 256 + threadIdx.x;
  double sq = 0.0;
  if (i < n4) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(q.G_raw) + i);
    float d[4] = {g.x * q.scale, g.y * q.scale, g.z * q.scale, g.w * q.scale};
    if (q.A) {
      const float4 av = __ldg(reinterpret_cast<const float4*>(q.A) + i);
      d[0] -= av.x; d[1] -= av.y; d[2] -= av.z; d[3] -= av.w;
    }
    store_d4(q.out, (size_t)i * 4, d, q.round_out);
    sq = (double)d[0] * d[0] + (double)d[1] * d[1] + (double)d[2] * d[2] + (double)d[3] * d[3];
  }
  if (!q.loss) return;
  ReduceWs* ws = ws_all + it;
  const double bs = block_sum(sq, red);
  if (threadIdx.x == 0) {
    ws->partials[b] = bs;
    __threadfence();
    is_last = atomicAdd(&ws->ticket, 1u) == (unsigned)nb - 1u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc = 0.0;
  for (int k = threadIdx.x; k < nb; k += blockDim.x) acc += ((volatile double*)ws->partials)[k];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    *q.loss = (float)(acc / ((double)q.C * (double)q.C));
    ws->ticket = 0u;
  }
}

static int launch_finalize(const GramPlan& plan, const float* partials, int symmetric_src, float scale,
                           const float* A, float* out, float* loss, void* ws, cudaStream_t stream,
                           int round_out = 0) {
  FinalizeArgs fa;
  fa.plan = plan;
  fa.partials = partials;
  fa.A = A;
  fa.out = out;
  fa.loss = loss;
  fa.scale = scale;
  fa.symmetric_src = symmetric_src;
  fa.round_out = round_out;
  const int elems = plan.n_tiles * plan.TR * plan.TR;
  fa.lanes = (elems <= 4096) ? 16 : (elems <= 16384 ? 8 : 4);
  if (plan.total_parts < 2 * fa.lanes) fa.lanes = 4;
  const int blocks = elems / ((256 / fa.lanes) * 4);
  if (blocks > kReduceMaxBlocks) {
    set_error("gram finalize: %d blocks exceed the reduce workspace", blocks);
    return AST_ERR_UNSUPPORTED;
  }
  gram_finalize_kernel<<<blocks, 256, 0, stream>>>(fa, (ReduceWs*)ws);
  return check_launch("gram_finalize");
}

static int fp32_splits(int C, int64_t HW) {
  const int tiles = (C / SG_T) * (C / SG_T);
  int64_t s = (2 * 148 + tiles - 1) / tiles;          // ~2 CTAs per SM
  const int64_t max_by_k = (HW + 4 * SG_K - 1) / (4 * SG_K);  // at least 4 K-chunks per split
  if (s > max_by_k) s = max_by_k;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return (int)s;
}

static int cached_num_sms() {
  // SM count of the current device; pure function of the device, cached per thread.
  static thread_local int dev_cached = -1, sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != dev_cached) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
    dev_cached = dev;
  }
  return sms;
}

}  // namespace ast

using namespace ast;

static inline bool is16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" size_t ast_gram_workspace_bytes(int C, int64_t HW) {
  if (C <= 0 || HW <= 0) return 0;
  size_t parts_fp32 = (C % SG_T == 0) ? (size_t)fp32_splits(C, HW) * C * C : 0;
  size_t parts_tc = 0;
  if (C == 64 || C == 128 || C == 256 || C == 512) {   // either layout (the NHWC path has no HW % 4 limit)
    GramPlan plan;
    gram_tc_plan(C, HW, 148, &plan);   // workspace sized for a full B200; fewer SMs never need more
    parts_tc = (size_t)plan.total_parts * plan.TR * plan.TR;
  }
  const size_t parts = parts_fp32 > parts_tc ? parts_fp32 : parts_tc;
  return kGramWsHeaderBytes + parts * sizeof(float);
}

extern "C" int ast_gram_tf32_supported(const float* F, int C, int64_t HW, int64_t ld) {
  return (gram_tc_supported(C, HW, F) && ld >= HW && (ld % 4 == 0)) ? 1 : 0;
}

extern "C" int ast_gram_mse_fwd(const float* F, int C, int64_t HW, int64_t ld, float scale, const float* A,
                                float* out, float* loss, void* ws, size_t ws_bytes, int precision, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AST_REQUIRE(F && out && ws, AST_ERR_INVALID, "ast_gram_mse_fwd: null pointer");
  AST_REQUIRE(C > 0 && HW > 0, AST_ERR_INVALID, "ast_gram_mse_fwd: bad shape C=%d HW=%lld", C, (long long)HW);
  AST_REQUIRE(ld >= HW, AST_ERR_INVALID, "ast_gram_mse_fwd: row pitch %lld < HW %lld", (long long)ld, (long long)HW);
  AST_REQUIRE(C % SG_T == 0, AST_ERR_UNSUPPORTED, "ast_gram_mse_fwd: C must be a multiple of 64 (got %d)", C);
  AST_REQUIRE(is16(out) && (!A || is16(A)) && is16(ws), AST_ERR_INVALID, "ast_gram_mse_fwd: out/A/ws must be 16-byte aligned");
  AST_REQUIRE(precision == AST_PREC_TF32 || precision == AST_PREC_FP32, AST_ERR_INVALID, "ast_gram_mse_fwd: bad precision %d", precision);
  AST_REQUIRE(ws_bytes >= ast_gram_workspace_bytes(C, HW), AST_ERR_WORKSPACE, "ast_gram_mse_fwd: workspace %zu < %zu",
              ws_bytes, ast_gram_workspace_bytes(C, HW));
  float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kGramWsHeaderBytes);
  GramPlan plan;
  if (precision == AST_PREC_TF32) {
    AST_REQUIRE(gram_tc_supported(C, HW, F) && (ld % 4 == 0), AST_ERR_UNSUPPORTED,
                "ast_gram_mse_fwd: TF32 path needs C in {64,128,256,512}, HW %% 4 == 0, ld %% 4 == 0 and 16-byte "
                "aligned F (C=%d HW=%lld ld=%lld); use AST_PREC_FP32", C, (long long)HW, (long long)ld);
    const int sms = cached_num_sms();
    gram_tc_plan(C, HW, sms < 148 ? sms : 148, &plan);
    int rc = gram_tc_fwd(F, C, HW, ld, 0, partials, plan, sms, stream);
    if (rc != AST_OK) return rc;
    return launch_finalize(plan, partials, 1, scale, A, out, loss, ws, stream);
  }
  const int splits = fp32_splits(C, HW);
  int64_t kps = (HW + splits - 1) / splits;
  kps = (kps + SG_K - 1) / SG_K * SG_K;
  const int vec_ok = is16(F) && (ld % 4 == 0);
  dim3 grid((C / SG_T) * (C / SG_T), splits);
  gram_fp32_fwd_kernel<<<grid, 256, 0, stream>>>(F, C, HW, ld, kps, vec_ok, partials);
  int rc = check_launch("gram_fp32_fwd");
  if (rc != AST_OK) return rc;
  plan.C = C; plan.TR = C; plan.n_tiles = 1;
  plan.tile_bi[0] = plan.tile_bj[0] = 0;
  plan.part_off[0] = 0; plan.part_cnt[0] = splits; plan.total_parts = splits;
  return launch_finalize(plan, partials, 0, scale, A, out, loss, ws, stream);
}

extern "C" int ast_gram_mse_fwd_nhwc(const float* F, int C, int64_t HW, float scale, const float* A, float* out,
                                     float* loss, void* ws, size_t ws_bytes, int round_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AST_REQUIRE(F && out && ws, AST_ERR_INVALID, "ast_gram_mse_fwd_nhwc: null pointer");
  AST_REQUIRE(HW > 0 && HW <= (int64_t)0x7fffff00, AST_ERR_INVALID, "ast_gram_mse_fwd_nhwc: bad HW=%lld", (long long)HW);
  AST_REQUIRE(C == 64 || C == 128 || C == 256 || C == 512, AST_ERR_UNSUPPORTED,
              "ast_gram_mse_fwd_nhwc: C must be 64, 128, 256 or 512 (got %d)", C);
  AST_REQUIRE(is16(F) && is16(out) && (!A || is16(A)) && is16(ws), AST_ERR_INVALID,
              "ast_gram_mse_fwd_nhwc: F/out/A/ws must be 16-byte aligned");
  AST_REQUIRE(ws_bytes >= ast_gram_workspace_bytes(C, HW), AST_ERR_WORKSPACE, "ast_gram_mse_fwd_nhwc: workspace %zu < %zu",
              ws_bytes, ast_gram_workspace_bytes(C, HW));
  float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kGramWsHeaderBytes);
  GramPlan plan;
  const int sms = cached_num_sms();
  gram_tc_plan(C, HW, sms < 148 ? sms : 148, &plan);
  int rc = gram_tc_fwd(F, C, HW, C, 1, partials, plan, sms, stream);
  if (rc != AST_OK) return rc;
  return launch_finalize(plan, partials, 1, scale, A, out, loss, ws, stream, round_out);
}

extern "C" int ast_gram_bwd_nhwc(const float* D, const float* F, int C, int64_t HW, float scale, const float* gscale,
                                 float* dF, int accumulate, int d_prerounded, int relu_mask, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AST_REQUIRE(D && F && dF, AST_ERR_INVALID, "ast_gram_bwd_nhwc: null pointer");
  AST_REQUIRE(HW > 0 && HW <= (int64_t)0x7fffff00, AST_ERR_INVALID, "ast_gram_bwd_nhwc: bad HW=%lld", (long long)HW);
  AST_REQUIRE(C == 64 || C == 128 || C == 256 || C == 512, AST_ERR_UNSUPPORTED,
              "ast_gram_bwd_nhwc: C must be 64, 128, 256 or 512 (got %d)", C);
  AST_REQUIRE(is16(F) && is16(D) && is16(dF), AST_ERR_INVALID, "ast_gram_bwd_nhwc: D/F/dF must be 16-byte aligned");
  return gram_tc_bwd_nhwc(D, F, C, HW, scale, gscale, dF, accumulate, d_prerounded ? 1 : 0, relu_mask ? 1 : 0,
                          cached_num_sms(), stream);
}

extern "C" int ast_gram_bwd_nhwc_bf16(const void* D_bf16, const float* F, int C, int64_t HW, float scale,
                                      const float* gscale, float* dF, int accumulate, int relu_mask, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AST_REQUIRE(D_bf16 && F && dF, AST_ERR_INVALID, "ast_gram_bwd_nhwc_bf16: null pointer");
  AST_REQUIRE(HW > 0 && HW <= (int64_t)0x7fffff00, AST_ERR_INVALID, "ast_gram_bwd_nhwc_bf16: bad HW=%lld", (long long)HW);
  AST_REQUIRE(C == 512, AST_ERR_UNSUPPORTED,
              "ast_gram_bwd_nhwc_bf16: BF16 operands are implemented for C = 512, the tensor-bound width (got %d); "
              "narrower layers are HBM-bound and use ast_gram_bwd_nhwc", C);
  AST_REQUIRE(is16(F) && is16(D_bf16) && is16(dF), AST_ERR_INVALID, "ast_gram_bwd_nhwc_bf16: D/F/dF must be 16-byte aligned");
  return gram_tc_bwd_nhwc_bf16(D_bf16, F, C, HW, scale, gscale, dF, accumulate, relu_mask ? 1 : 0, cached_num_sms(), stream);
}

extern "C" int ast_gram_finalize(const float* G_raw, int C, float scale, const float* A, float* out, float* loss,
                                 void* ws, size_t ws_bytes, int round_out, void* stream) {
  AST_REQUIRE(G_raw && out && ws, AST_ERR_INVALID, "ast_gram_finalize: null pointer");
  AST_REQUIRE(C > 0 && C % 16 == 0, AST_ERR_UNSUPPORTED, "ast_gram_finalize: C must be a multiple of 16 (got %d)", C);
  AST_REQUIRE(is16(G_raw) && is16(out) && (!A || is16(A)), AST_ERR_INVALID, "ast_gram_finalize: pointers must be 16-byte aligned");
  AST_REQUIRE(ws_bytes >= sizeof(ReduceWs), AST_ERR_WORKSPACE, "ast_gram_finalize: workspace %zu < %zu", ws_bytes, sizeof(ReduceWs));
  GramPlan plan;
  plan.C = C; plan.TR = C; plan.n_tiles = 1;
  plan.tile_bi[0] = plan.tile_bj[0] = 0;
  plan.part_off[0] = 0; plan.part_cnt[0] = 1; plan.total_parts = 1;
  return launch_finalize(plan, G_raw, 0, scale, A, out, loss, ws, (cudaStream_t)stream, round_out);
}

extern "C" int ast_gram_bwd(const float* D, const float* F, int C, int64_t HW, int64_t ld, float scale,
                            const float* gscale, float* dF, int accumulate, int precision, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AST_REQUIRE(D && F && dF, AST_ERR_INVALID, "ast_gram_bwd: null pointer");
  AST_REQUIRE(C > 0 && HW > 0, AST_ERR_INVALID, "ast_gram_bwd: bad shape C=%d HW=%lld", C, (long long)HW);
  AST_REQUIRE(ld >= HW, AST_ERR_INVALID, "ast_gram_bwd: row pitch %lld < HW %lld", (long long)ld, (long long)HW);
  AST_REQUIRE(C % SG_T == 0, AST_ERR_UNSUPPORTED, "ast_gram_bwd: C must be a multiple of 64 (got %d)", C);
  AST_REQUIRE(precision == AST_PREC_TF32 || precision == AST_PREC_FP32, AST_ERR_INVALID, "ast_gram_bwd: bad precision %d", precision);
  if (precision == AST_PREC_TF32) {
    AST_REQUIRE(gram_tc_supported(C, HW, F) && (ld % 4 == 0) && is16(D), AST_ERR_UNSUPPORTED,
                "ast_gram_bwd: TF32 path needs C in {64,128,256,512}, HW %% 4 == 0, ld %% 4 == 0 and 16-byte aligned "
                "F/D (C=%d HW=%lld ld=%lld); use AST_PREC_FP32", C, (long long)HW, (long long)ld);
    return gram_tc_bwd(D, F, C, HW, ld, scale, gscale, dF, accumulate, cached_num_sms(), stream);
  }
  const int vec_ok = is16(F) && (ld % 4 == 0);
  dim3 grid((unsigned)((HW + SG_T - 1) / SG_T), C / SG_T);
  gram_fp32_bwd_kernel<<<grid, 256, 0, stream>>>(D, F, C, HW, ld, scale, gscale, dF, accumulate, vec_ok);
  return check_launch("gram_fp32_bwd");
}

extern "C" size_t ast_finalize_batch_workspace_bytes(int n_items) {
  return n_items > 0 ? (size_t)n_items * sizeof(ReduceWs) : 0;
}

extern "C" int ast_gram_finalize_batch(const ast_finalize_item* items, int n_items, void* ws, size_t ws_bytes,
                                       void* stream) {
  AST_REQUIRE(items && ws && n_items > 0 && n_items <= AST_FINALIZE_MAX_ITEMS, AST_ERR_INVALID,
              "ast_gram_finalize_batch: n_items must be 1..%d (got %d)", AST_FINALIZE_MAX_ITEMS, n_items);
  AST_REQUIRE(ws_bytes >= ast_finalize_batch_workspace_bytes(n_items) && is16(ws), AST_ERR_WORKSPACE,
              "ast_gram_finalize_batch: workspace %zu < %zu", ws_bytes, ast_finalize_batch_workspace_bytes(n_items));
  FinalizeBatchArgs a = {};
  a.n = n_items;
  int off = 0;
  for (int k = 0; k < n_items; ++k) {
    const ast_finalize_item& q = items[k];
    AST_REQUIRE(q.G_raw && q.out && q.C > 0 && q.C % 16 == 0 && q.C <= 1024, AST_ERR_INVALID,
                "ast_gram_finalize_batch: item %d: null pointer or bad C=%d", k, q.C);
    AST_REQUIRE(is16(q.G_raw) && is16(q.out) && (!q.A || is16(q.A)), AST_ERR_INVALID,
                "ast_gram_finalize_batch: item %d: pointers must be 16-byte aligned", k);
    a.items[k] = q;
    a.block_off[k] = off;
    off += (q.C * q.C / 4 + 255) / 256;            // <= 1024 blocks per item = kReduceMaxBlocks
  }
  a.block_off[n_items] = off;
  gram_finalize_batch_kernel<<<off, 256, 0, (cudaStream_t)stream>>>(a, (ReduceWs*)ws);
  return check_launch("gram_finalize_batch");
}
