// tcgen05 / TMEM / TMA Gram kernels for sm_100a (AST_PREC_TF32).
//
// Forward  (K1): G_partial = F[:, ks] F[:, ks]^T.  F is (C, HW) row-major, i.e. K-major for BOTH operands, so a
//   CTA TMA-loads ONE slab of rows x 32 fp32 (128-byte swizzle atom) per K chunk and points both UMMA
//   descriptors at it.  Split-K over the CTAs (one CTA per SM); accumulators live in TMEM; each CTA writes its
//   partial tile to the workspace and gram_finalize_kernel (gram.cu) reduces them in a fixed order.
// Backward (K2): dF = s * D F, computed transposed, dF^T[n, c] = sum_k F[k, n] D[c, k]:  A = F tile as an
//   MN-major operand (M = 128 spatial positions, contiguous), B = D (K-major), accumulator lane = spatial
//   position so the epilogue's global loads/stores are 128-byte coalesced.  tcgen05 accepts a swizzled
//   MN-major 32-bit operand only in the SWIZZLE_128B_BASE32B layout (32-byte swizzle atoms), which TMA
//   produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (the plain 128B swizzle silently yields zeros).  D stays resident in shared memory
//   for C <= 128 and is streamed from L2 per K chunk for C >= 256.
//
// Operands are fp32 in HBM.  tcgen05 kind::tf32 reads fp32 bit patterns and TRUNCATES the low 13 mantissa bits,
// which biases a Gram of non-negative (post-ReLU) features by ~1e-3 relative.  A converter warpgroup therefore
// rounds each staged tile to nearest TF32 in place (cvt.rna.tf32.f32) between the TMA landing and the MMA.
//
// Pipeline per stage:  TMA (warp 0) --full--> converters (4 warps) --conv--> MMA (warp 1) --empty--> TMA
// and                  MMA --acc_full--> epilogue warps (tcgen05.ld) [--acc_empty--> MMA in the backward].
#include <cstdio>
#include <cstdlib>

#include "gram.cuh"
#include "sm100_ptx.cuh"

namespace ast {

using namespace ptx;

constexpr int BK = 32;                 // fp32 elements per 128-byte swizzle row
constexpr int ROW_BYTES = BK * 4;      // 128

// ------------------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  // cuTensorMapEncodeTiled is a DRIVER call: it needs a context current on the calling thread.  A thread whose first CUDA
  // work is one of our launches (torch's autograd engine thread on device 0, when no allocation preceded it) has none
  // yet — CUDA_ERROR_INVALID_CONTEXT (201), seen on a B200 in the first backward of a sharded job.  cudaFree(0) binds
  // the primary context; once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  static EncodeTiledFn fn = nullptr;   // resolved once; pure function of the driver
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 row-major (rows, cols) tensor, box = (box_rows x 32 cols), 128-byte swizzle, zero OOB fill.
static int make_tmap(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t pitch,
                     uint32_t box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return AST_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%llu cols=%llu box_rows=%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows);
    return AST_ERR_CUDA;
  }
  return AST_OK;
}

// (HW, C) row-major fp32 operand viewed as (32 channels, HW positions, C/32 strips); box = 32 x 32 x box_strips
// lands in shared memory as box_strips consecutive 4 KB strips of [32 positions][32 channels] in the
// SWIZZLE_128B_BASE32B pattern tcgen05 needs for an MN-major 32-bit operand.  Positions >= HW read as zero.
static int make_tmap_nhwc_strips(CUtensorMap* m, const float* base, int C, uint64_t HW, uint32_t box_strips) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return AST_ERR_CUDA;
  }
  cuuint64_t gdim[3] = {32, HW, (cuuint64_t)(C / 32)};
  cuuint64_t gstride[2] = {(cuuint64_t)C * sizeof(float), 32 * sizeof(float)};
  cuuint32_t box[3] = {32, 32, box_strips};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (nhwc strips) failed (CUresult %d) C=%d HW=%llu", (int)r, C,
              (unsigned long long)HW);
    return AST_ERR_CUDA;
  }
  return AST_OK;
}

// ------------------------------------------------------------------------------------------------------
// shared device pieces
// ------------------------------------------------------------------------------------------------------
// Round `bytes` of staged fp32 to nearest TF32 in place; 128 converter threads, 16 B per access, explicit
// shared-space LDS.128 / STS.128 (a generic pointer would compile to LD.E / ST.E and double the wavefronts).
// tcgen05 drops the low 13 mantissa bits itself, so adding half a TF32 ulp (0x1000) to the bit pattern is the
// whole rounding (round-half-away on the magnitude); Inf/NaN patterns are left untouched.
__device__ __forceinline__ uint32_t tf32_half_ulp(uint32_t u) {
  return ((u & 0x7f800000u) == 0x7f800000u) ? u : u + 0x1000u;
}
template <int NT = 128>
__device__ __forceinline__ void convert_tf32_inplace(uint8_t* base, int bytes, int ctid) {
  const uint32_t s0 = smem_u32(base);
  const int n = bytes >> 4;
  // batches of 8 independent 16-byte accesses per thread (16 KB per batch for 128 threads): the eight
  // LDS.128 are in flight together, so a batch costs one shared-memory round trip instead of eight
  int i = ctid;
  for (; i + 7 * NT < n; i += 8 * NT) {
    uint32_t v[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t addr = s0 + 16u * (uint32_t)(i + j * NT);
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(v[j][0]), "=r"(v[j][1]), "=r"(v[j][2]), "=r"(v[j][3])
                   : "r"(addr));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t addr = s0 + 16u * (uint32_t)(i + j * NT);
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(tf32_half_ulp(v[j][0])),
                   "r"(tf32_half_ulp(v[j][1])), "r"(tf32_half_ulp(v[j][2])), "r"(tf32_half_ulp(v[j][3]))
                   : "memory");
    }
  }
  for (; i < n; i += NT) {
    uint32_t a, b, c, d;
    const uint32_t addr = s0 + 16u * (uint32_t)i;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
    a = tf32_half_ulp(a); b = tf32_half_ulp(b); c = tf32_half_ulp(c); d = tf32_half_ulp(d);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
  }
}

struct FwdParams {
  int64_t HW;
  int n_tiles;
  int cta_off[kGramMaxTiles + 1];   // CTAs of tile t: [cta_off[t], cta_off[t+1])
  int tile_bi[kGramMaxTiles], tile_bj[kGramMaxTiles];
  int part_off[kGramMaxTiles];
  int units_total;                  // K units (stage fills) covering HW
  int skip_rounding;                // tuning probe only (AST_GRAM_FWD_NOROUND=1): leave the operands truncated
  float* partials;
};

// NU = K units per pipeline stage (one mbarrier round trip per stage), G = converter groups of 128 threads.  The
// groups take ALTERNATE ring slots (group g rounds slots g, g + G, ...): rounding a stage is a latency chain — wait for the
// TMA, one shared-memory round trip, membar, proxy fence, arrive; ncu (profiles/r02_ncu_gram_stalls.md) shows one group
// finishing a 16 KB stage every ~700 cycles while HBM delivers one every ~600 — and two groups working on different
// stages overlap their chains.  (Round 1 split every stage over all 128 * G threads instead, which shortens the
// copy but not the chain: < 1 % gain.)
template <int C, int NU, int G>
struct FwdCfg {
  static constexpr int kBoxRows = (C == 64) ? 64 : (C == 128 ? 128 : 256);
  // shared memory per K unit: C=64 two stacked 64x32 chunks, C=128 one 128x32 chunk, C=256 one 256x32 chunk,
  // C=512 room for two 256x32 chunks (the diagonal tiles fill only the first)
  static constexpr int kUnitBytes = (C == 512) ? 65536 : (C == 256 ? 32768 : 16384);
  static constexpr int kStageBytes = NU * kUnitBytes;
  static constexpr int kStagesFit = (208 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesFit > 12 ? 12 : kStagesFit;
  static constexpr int kThreads = 64 + 128 * G;
  static constexpr int kTmemCols = (C <= 128) ? 128 : 512;
  static constexpr int kTR = (C <= 256) ? C : 256;                     // partial tile edge
  static constexpr int kUmmaN = (C <= 128) ? 128 : 256;
  // C = 512: the two diagonal tiles fill only half of a 64 KB stage, so their CTAs cut the same ring into twice as
  // many 32 KB stages (6 instead of 3: the MMAs of a 32 KB stage last ~640 cycles, three stages in flight did not
  // cover the TMA latency).  kMaxStages bounds the barrier arrays.
  static constexpr int kMaxStages = (C == 512) ? 2 * kStages : kStages;
  static constexpr int kBarBytes = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + kBarBytes;
  static_assert(kStages >= 2, "pipeline needs two stages");
  static_assert((3 * kMaxStages + 1) * 8 + 8 <= kBarBytes, "barrier area too small");
};

// warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..2+4G: converters, then epilogue.
template <int C, int NU, int G, bool NHWC>
__global__ void __launch_bounds__(64 + 128 * G, 1) gram_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                    const __grid_constant__ FwdParams P) {
  using Cfg = FwdCfg<C, NU, G>;
  constexpr int SM = Cfg::kMaxStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  // bars[0..SM) full, [SM..2SM) conv, [2SM..3SM) empty, [3SM] acc_full ; then tmem slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * SM + 1);
  const uint32_t bar_full = smem_u32(bars), bar_conv = smem_u32(bars + SM), bar_empty = smem_u32(bars + 2 * SM),
                 bar_acc = smem_u32(bars + 3 * SM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // which tile / split am I?
  int t = 0;
  while (t + 1 < P.n_tiles && (int)blockIdx.x >= P.cta_off[t + 1]) ++t;
  const int split = blockIdx.x - P.cta_off[t];
  const int nsplit = P.cta_off[t + 1] - P.cta_off[t];
  const int base_u = P.units_total / nsplit, rem_u = P.units_total % nsplit;
  const int u_begin = split * base_u + min(split, rem_u);
  const int n_units = base_u + (split < rem_u ? 1 : 0);
  const int n_stages = (n_units + NU - 1) / NU;
  const bool offdiag = (C == 512) && (P.tile_bi[t] != P.tile_bj[t]);
  const int unit_tx_bytes = (C == 512) ? (offdiag ? 65536 : 32768) : Cfg::kUnitBytes;
  // ring geometry of this CTA (see kMaxStages): stage stride and count
  const int unit_stride = (C == 512 && !offdiag) ? 32768 : Cfg::kUnitBytes;
  const int stage_stride = NU * unit_stride;
  const int S = (C == 512 && !offdiag) ? SM : Cfg::kStages;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap);
    for (int s = 0; s < SM; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, 128);       // ONE converter group per stage
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int i = 0; i < n_stages; ++i) {
        const int s = i % S;
        const uint32_t ph = (uint32_t)(i / S) & 1u;
        const int nu = min(NU, n_units - i * NU);
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)(nu * unit_tx_bytes));
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          if (j < nu) {
            const uint32_t dst = smem_u32(smem + s * stage_stride + j * unit_stride);
            const int u = u_begin + i * NU + j;
            if (NHWC) {
              // (HW, C) operand: 4 KB strips of [32 positions][32 channels]; coordinates (channel in strip,
              // position, strip)
              if (C == 64) {
                tma_load_3d(dst, &tmap, bar_full + 8 * s, 0, (2 * u) * BK, 0);
                tma_load_3d(dst + 2 * 4096, &tmap, bar_full + 8 * s, 0, (2 * u + 1) * BK, 0);
              } else if (C <= 256) {
                tma_load_3d(dst, &tmap, bar_full + 8 * s, 0, u * BK, 0);
              } else {
                tma_load_3d(dst, &tmap, bar_full + 8 * s, 0, u * BK, P.tile_bi[t] * 8);
                if (offdiag) tma_load_3d(dst + 8 * 4096, &tmap, bar_full + 8 * s, 0, u * BK, P.tile_bj[t] * 8);
              }
            } else if (C == 64) {
              tma_load_2d(dst, &tmap, bar_full + 8 * s, (2 * u) * BK, 0);
              tma_load_2d(dst + 64 * ROW_BYTES, &tmap, bar_full + 8 * s, (2 * u + 1) * BK, 0);
            } else if (C <= 256) {
              tma_load_2d(dst, &tmap, bar_full + 8 * s, u * BK, 0);
            } else {
              tma_load_2d(dst, &tmap, bar_full + 8 * s, u * BK, P.tile_bi[t] * 256);
              if (offdiag) tma_load_2d(dst + 256 * ROW_BYTES, &tmap, bar_full + 8 * s, u * BK, P.tile_bj[t] * 256);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_tf32(128, Cfg::kUmmaN, NHWC ? 1 : 0, NHWC ? 1 : 0);
      for (int i = 0; i < n_stages; ++i) {
        const int s = i % S;
        const uint32_t ph = (uint32_t)(i / S) & 1u;
        const int nu = min(NU, n_units - i * NU);
        mbar_wait(bar_conv + 8 * s, ph);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          if (j < nu) {
            const uint32_t a_base = smem_u32(smem + s * stage_stride + j * unit_stride);
            const uint32_t b_base = offdiag ? a_base + 256 * ROW_BYTES : a_base;
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint32_t acc = (i > 0 || j > 0 || k > 0) ? 1u : 0u;
              if (NHWC) {
                // MN-major operands (SWIZZLE_128B_BASE32B): 8 positions (1024 B) per K step, channel strips 4096 B
                // apart (LBO), 4-row swizzle atoms 512 B apart (SBO)
                const uint64_t bd = umma_desc(b_base + k * 1024, 4096, 512, 1);
                umma_tf32(tmem_base, umma_desc(a_base + k * 1024, 4096, 512, 1), bd, idesc, acc);
                if (C > 128) umma_tf32(tmem_base + 256, umma_desc(a_base + 4 * 4096 + k * 1024, 4096, 512, 1), bd, idesc, acc);
              } else {
                const uint64_t bd = umma_desc_sw128(b_base + k * 32, 16, 1024);
                umma_tf32(tmem_base, umma_desc_sw128(a_base + k * 32, 16, 1024), bd, idesc, acc);
                if (C > 128)
                  umma_tf32(tmem_base + 256, umma_desc_sw128(a_base + 128 * ROW_BYTES + k * 32, 16, 1024), bd, idesc,
                            acc);
              }
            }
          }
        }
        umma_commit(bar_empty + 8 * s);
      }
      umma_commit(bar_acc);
    }
  } else {
    // ===== converters (G groups of 128 threads taking alternate stages), then epilogue =====
    const int grp = (warp - 2) >> 2;
    const int ctid = (threadIdx.x - 64) & 127;       // thread within its group
    for (int i = 0; i < n_stages; ++i) {
      const int s = i % S;
      // a ring slot always belongs to the same group: a group that skipped a phase of full[s] could not tell the
      // phase it waits for from the one before it by parity alone (S = 3 with G = 2 raced exactly like that)
      if (G > 1 && (s % G) != grp) continue;
      const uint32_t ph = (uint32_t)(i / S) & 1u;
      const int nu = min(NU, n_units - i * NU);
      mbar_wait(bar_full + 8 * s, ph);
      if (P.skip_rounding) {
      } else if (unit_tx_bytes == unit_stride) {
        convert_tf32_inplace<128>(smem + s * stage_stride, nu * unit_stride, ctid);
      } else {
        for (int j = 0; j < nu; ++j)
          convert_tf32_inplace<128>(smem + s * stage_stride + j * unit_stride, unit_tx_bytes, ctid);
      }
      fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(bar_conv + 8 * s);
    }
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int sub = warp & 3;            // TMEM sub-partition this warp may read
    const int row = sub * 32 + lane;     // accumulator lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(sub * 32) << 16);
    constexpr int TR = Cfg::kTR;
    constexpr int kRowBlocks = (C <= 128) ? 1 : 2;
    constexpr int kColGroups = (C == 64) ? 2 : (C == 128 ? 4 : 8);   // 32-column groups per accumulator row
    uint32_t v[32];
    // C == 64, stacked units: rows 0..63 x cols 0..63 and rows 64..127 x cols 64..127 are two partial Grams
    const int half = (C == 64) ? (row >> 6) : 0;
    float* tile = (C == 64)
                      ? P.partials + ((size_t)(P.part_off[t] + 2 * split + half) * TR + (row & 63)) * TR
                      : P.partials + (size_t)(P.part_off[t] + split) * TR * TR + (size_t)row * TR;
#pragma unroll 1
    for (int q = grp; q < kRowBlocks * kColGroups; q += G) {
      const int rb = q / kColGroups, cg = q % kColGroups;
      tmem_ld_x32(lane_addr + rb * 256 + half * 64 + cg * 32, v);
      tmem_ld_wait();
      float* dst = tile + (size_t)(rb * 128) * TR + cg * 32;
#pragma unroll
      for (int e = 0; e < 8; ++e)
        reinterpret_cast<float4*>(dst)[e] =
            make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                        __uint_as_float(v[4 * e + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------
struct BwdParams {
  int64_t HW;
  int64_t ld;
  int n_tiles;       // ceil(HW / 128)
  float scale;
  const float* gscale;   // nullable device scalar multiplied into scale
  int accumulate;
  float* dF;
};

template <int C>
struct BwdCfg {
  static constexpr bool kResidentD = (C <= 128);
  static constexpr int kFBytes = 128 * BK * 4;                          // 16 KB: 32 channels x 128 positions
  static constexpr int kDChunkBytes = C * ROW_BYTES;                    // C rows x 32 k
  static constexpr int kStageBytes = kFBytes + (kResidentD ? 0 : kDChunkBytes);
  static constexpr int kStages = (C == 512) ? 2 : (C == 256 ? 4 : 6);
  static constexpr int kDResBytes = kResidentD ? C * C * 4 : 0;
  static constexpr int kAccBufs = (C == 512) ? 1 : 2;
  static constexpr int kTmemCols = (C == 64) ? 128 : (C == 128 ? 256 : 512);
  static constexpr int kUmmaN = (C <= 256) ? C : 256;
  static constexpr int kDBoxRows = (C <= 256) ? C : 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kDResBytes + 1024 + 256;
};

// warp 0: TMA, warp 1: MMA + TMEM, warps 2..5: converters, warps 6..9: epilogue.
template <int C>
__global__ void __launch_bounds__(320, 1) gram_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmapF,
                                                            const __grid_constant__ CUtensorMap tmapD,
                                                            const __grid_constant__ BwdParams P) {
  using Cfg = BwdCfg<C>;
  constexpr int S = Cfg::kStages;
  constexpr int KC = C / BK;            // K chunks per tile
  constexpr int NB = Cfg::kAccBufs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* dres = smem + S * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dres + Cfg::kDResBytes);
  // full[S] conv[S] empty[S] acc_full[2] acc_empty[2] d_full d_conv
  const uint32_t bar_full = smem_u32(bars), bar_conv = smem_u32(bars + S), bar_empty = smem_u32(bars + 2 * S),
                 bar_accf = smem_u32(bars + 3 * S), bar_acce = smem_u32(bars + 3 * S + 2),
                 bar_dfull = smem_u32(bars + 3 * S + 4), bar_dconv = smem_u32(bars + 3 * S + 5);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 6);
  // q[epilogue warp][kOutBufs]: "the running-gradient block has landed in this output buffer" (fused ReLU backward)
  const uint32_t bar_qall = smem_u32(bars + 3 * S + 7);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = ((int)blockIdx.x < P.n_tiles) ? (P.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmapF);
    prefetch_tmap(&tmapD);
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, 128);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accf + 8 * b, 1);
      mbar_init(bar_acce + 8 * b, 128);
    }
    mbar_init(bar_dfull, 1);
    mbar_init(bar_dconv, 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      if (Cfg::kResidentD) {
        mbar_arrive_expect_tx(bar_dfull, (uint32_t)Cfg::kDResBytes);
        for (int kc = 0; kc < KC; ++kc)
          tma_load_2d(smem_u32(dres + kc * Cfg::kDChunkBytes), &tmapD, bar_dfull, kc * BK, 0);
      }
      int it = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int64_t n0 = ((int64_t)blockIdx.x + (int64_t)ti * gridDim.x) * 128;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)Cfg::kStageBytes);
          const uint32_t dst = smem_u32(smem + s * Cfg::kStageBytes);
#pragma unroll
          for (int j = 0; j < 4; ++j)   // four 32-position strips, 32 channel rows each (4 KB)
            tma_load_2d(dst + j * 4096, &tmapF, bar_full + 8 * s, (int)(n0 + 32 * j), kc * BK);
          if (!Cfg::kResidentD) {
            tma_load_2d(dst + Cfg::kFBytes, &tmapD, bar_full + 8 * s, kc * BK, 0);
            if (C == 512) tma_load_2d(dst + Cfg::kFBytes + 256 * ROW_BYTES, &tmapD, bar_full + 8 * s, kc * BK, 256);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_tf32(128, Cfg::kUmmaN, /*A MN-major*/ 1, /*B K-major*/ 0);
      if (Cfg::kResidentD) {
        mbar_wait(bar_dconv, 0);
        tc_fence_after();
      }
      int it = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int b = ti % NB;
        const uint32_t bph = (uint32_t)(ti / NB) & 1u;
        mbar_wait(bar_acce + 8 * b, bph ^ 1u);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * C;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(bar_conv + 8 * s, ph);
          tc_fence_after();
          const uint32_t f_base = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t d_base = Cfg::kResidentD ? smem_u32(dres + kc * Cfg::kDChunkBytes) : f_base + Cfg::kFBytes;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint32_t acc = (kc > 0 || k > 0) ? 1u : 0u;
            // A: 8 channel rows (1024 B) per K step; 32-position strips 4096 B apart (LBO)
            // SWIZZLE_128B_BASE32B: 4-row (512 B) swizzle atoms, two per K step (SBO); strips 4096 B apart (LBO)
            const uint64_t ad = umma_desc(f_base + k * 1024, 4096, 512, 1);
            umma_tf32(d_tmem, ad, umma_desc_sw128(d_base + k * 32, 16, 1024), idesc, acc);
            if (C == 512)
              umma_tf32(d_tmem + 256, ad, umma_desc_sw128(d_base + 256 * ROW_BYTES + k * 32, 16, 1024), idesc, acc);
          }
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_accf + 8 * b);
      }
    }
  } else if (warp < 6) {
    // ===== converters =====
    const int ctid = threadIdx.x - 64;
    if (Cfg::kResidentD) {
      mbar_wait(bar_dfull, 0);
      convert_tf32_inplace(dres, Cfg::kDResBytes, ctid);
      fence_proxy_async_smem();
      mbar_arrive(bar_dconv);
    }
    const int total = my_tiles * KC;
    for (int it = 0; it < total; ++it) {
      const int s = it % S;
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      mbar_wait(bar_full + 8 * s, ph);
      convert_tf32_inplace(smem + s * Cfg::kStageBytes, Cfg::kStageBytes, ctid);
      fence_proxy_async_smem();
      mbar_arrive(bar_conv + 8 * s);
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (+=) global, 128-byte coalesced per channel =====
    const int sub = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(sub * 32) << 16);
    const float scale = P.gscale ? P.scale * __ldg(P.gscale) : P.scale;
    uint32_t v[32];
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int b = ti % NB;
      const uint32_t bph = (uint32_t)(ti / NB) & 1u;
      const int64_t n = ((int64_t)blockIdx.x + (int64_t)ti * gridDim.x) * 128 + sub * 32 + lane;
      mbar_wait(bar_accf + 8 * b, bph);
      tc_fence_after();
      float* out = P.dF + n;
#pragma unroll 1
      for (int g = 0; g < C / 32; ++g) {
        tmem_ld_x32(lane_addr + b * C + g * 32, v);
        tmem_ld_wait();
        if (n < P.HW) {
          float* o = out + (size_t)(g * 32) * P.ld;
          if (P.accumulate) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[(size_t)j * P.ld] += scale * __uint_as_float(v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[(size_t)j * P.ld] = scale * __uint_as_float(v[j]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acce + 8 * b);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------------
// backward, (HW, C) layout (torch channels_last):  dF[p, c] (+)= s * sum_k F[p, k] D[c, k]
// ------------------------------------------------------------------------------------------------------
// A = F tile [128 positions x 32 channels] K-major (channels are contiguous), B = D chunk [C x 32] K-major, the
// accumulator lane is the position and the column the output channel.  The epilogue stages each 32 x 32 block in
// shared memory (128B swizzle, conflict-free) and hands it to TMA: a tiled store, or a tiled reduce-add when the
// result accumulates into an existing gradient (the add happens in L2; the SM never reads dF).
struct BwdNhwcParams {
  int64_t HW;
  int n_tiles;       // ceil(HW / 128)
  float scale;
  const float* gscale;
  int accumulate;
  int d_prerounded;  // D is already TF32-representable: the converters leave it alone
  int relu_mask;     // fuse the backward of the ReLU that produced F: out = (accumulate ? dF + v : v) * (F > 0)
  const float* F;    // (HW, C), read again by the epilogue when relu_mask is set (L2-hot: the tile was just staged)
  float* dF;
  int C;
  int relay_release; // CTA-pair kernel: relay with release.cluster arrives (A/B switch, default relaxed)
};

// W ("wide epilogue", C = 256 only): 3 stages + 2 epilogue groups instead of 4 + 1 — the pipeline shape for the
// fused ReLU backward, whose epilogue keeps one TMA load of the running gradient in flight per warp.
template <int C, bool W = false>
struct BwdNhwcCfg {
  static_assert(!W || C == 256, "the wide-epilogue variant exists for C = 256 only");
  static constexpr bool kResidentD = (C <= 128);
  static constexpr int kFBytes = 128 * ROW_BYTES;                       // 16 KB: 128 positions x 32 channels
  static constexpr int kDChunkBytes = C * ROW_BYTES;                    // C rows x 32 k
  static constexpr int kStageBytes = kFBytes + (kResidentD ? 0 : kDChunkBytes);
  static constexpr int kStages = (C == 512) ? 2 : (C == 256 ? (W ? 3 : 4) : (C == 128 ? 4 : 6));
  static constexpr int kDResBytes = kResidentD ? C * C * 4 : 0;
  static constexpr int kAccBufs = (C == 512) ? 1 : 2;
  static constexpr int kTmemCols = (C == 64) ? 128 : (C == 128 ? 256 : 512);
  static constexpr int kUmmaN = (C <= 256) ? C : 256;
  static constexpr int kDBoxRows = (C <= 256) ? C : 256;
  // epilogue groups of 4 warps (one warp per TMEM sub-partition); group e takes the 32-column groups g with
  // g % kEpiGroups == e.  Two groups where the epilogue also reads global memory (fused ReLU backward): the loads
  // in flight per SM, not the sectors per request, bound that path.
  static constexpr int kEpiGroups = (C == 256 && !W) ? 1 : 2;   // measured: plain C = 256 is faster with 4 stages + 1 group
  static constexpr int kThreads = 192 + 128 * kEpiGroups;
  // per epilogue warp: kOutBufs 32 x 32 fp32 blocks.  Two are enough for store-only epilogues; the fused ReLU
  // backward also LOADS the running gradient block through TMA into the buffer it will be stored from, and a third
  // buffer keeps two loads in flight per warp (C = 64 / 128, where the shared-memory budget allows it).
  static constexpr int kOutBufs = (C <= 128) ? 3 : 2;
  static constexpr int kOutBytes = 4 * kEpiGroups * kOutBufs * 4096;
  static constexpr int kSmemBytes = kStages * kStageBytes + kDResBytes + kOutBytes + 1024 + 512;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// warp 0: TMA loads, warp 1: MMA + TMEM, warps 2..5: converters, warps 6..: epilogue groups + TMA stores.
template <int C, bool W = false>
__global__ void __launch_bounds__(BwdNhwcCfg<C, W>::kThreads, 1) gram_bwd_nhwc_tc_kernel(const __grid_constant__ CUtensorMap tmapF,
                                                                 const __grid_constant__ CUtensorMap tmapD,
                                                                 const __grid_constant__ CUtensorMap tmapO,
                                                                 const __grid_constant__ BwdNhwcParams P) {
  using Cfg = BwdNhwcCfg<C, W>;
  constexpr int S = Cfg::kStages;
  constexpr int KC = C / BK;            // K chunks per tile
  constexpr int NB = Cfg::kAccBufs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* dres = smem + S * Cfg::kStageBytes;
  uint8_t* ostage = dres + Cfg::kDResBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ostage + Cfg::kOutBytes);
  // full[S] conv[S] empty[S] acc_full[2] acc_empty[2] d_full d_conv
  const uint32_t bar_full = smem_u32(bars), bar_conv = smem_u32(bars + S), bar_empty = smem_u32(bars + 2 * S),
                 bar_accf = smem_u32(bars + 3 * S), bar_acce = smem_u32(bars + 3 * S + 2),
                 bar_dfull = smem_u32(bars + 3 * S + 4), bar_dconv = smem_u32(bars + 3 * S + 5);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 6);
  // q[epilogue warp][kOutBufs]: "the running-gradient block has landed in this output buffer" (fused ReLU backward)
  const uint32_t bar_qall = smem_u32(bars + 3 * S + 7);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = ((int)blockIdx.x < P.n_tiles) ? (P.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmapF);
    prefetch_tmap(&tmapD);
    prefetch_tmap(&tmapO);
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, 128);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accf + 8 * b, 1);
      mbar_init(bar_acce + 8 * b, 128 * Cfg::kEpiGroups);
    }
    mbar_init(bar_dfull, 1);
    mbar_init(bar_dconv, 128);
    for (int q = 0; q < 4 * Cfg::kEpiGroups * Cfg::kOutBufs; ++q) mbar_init(bar_qall + 8 * q, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      if (Cfg::kResidentD) {
        mbar_arrive_expect_tx(bar_dfull, (uint32_t)Cfg::kDResBytes);
        for (int kc = 0; kc < KC; ++kc)
          tma_load_2d(smem_u32(dres + kc * Cfg::kDChunkBytes), &tmapD, bar_dfull, kc * BK, 0);
      }
      int it = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int64_t n0 = ((int64_t)blockIdx.x + (int64_t)ti * gridDim.x) * 128;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)Cfg::kStageBytes);
          const uint32_t dst = smem_u32(smem + s * Cfg::kStageBytes);
          tma_load_2d(dst, &tmapF, bar_full + 8 * s, kc * BK, (int)n0);
          if (!Cfg::kResidentD) {
            tma_load_2d(dst + Cfg::kFBytes, &tmapD, bar_full + 8 * s, kc * BK, 0);
            if (C == 512) tma_load_2d(dst + Cfg::kFBytes + 256 * ROW_BYTES, &tmapD, bar_full + 8 * s, kc * BK, 256);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_tf32(128, Cfg::kUmmaN, 0, 0);
      if (Cfg::kResidentD) {
        mbar_wait(bar_dconv, 0);
        tc_fence_after();
      }
      int it = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int b = ti % NB;
        const uint32_t bph = (uint32_t)(ti / NB) & 1u;
        mbar_wait(bar_acce + 8 * b, bph ^ 1u);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * C;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(bar_conv + 8 * s, ph);
          tc_fence_after();
          const uint32_t f_base = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t d_base = Cfg::kResidentD ? smem_u32(dres + kc * Cfg::kDChunkBytes) : f_base + Cfg::kFBytes;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint32_t acc = (kc > 0 || k > 0) ? 1u : 0u;
            const uint64_t ad = umma_desc_sw128(f_base + k * 32, 16, 1024);
            umma_tf32(d_tmem, ad, umma_desc_sw128(d_base + k * 32, 16, 1024), idesc, acc);
            if (C == 512)
              umma_tf32(d_tmem + 256, ad, umma_desc_sw128(d_base + 256 * ROW_BYTES + k * 32, 16, 1024), idesc, acc);
          }
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_accf + 8 * b);
      }
    }
  } else if (warp < 6) {
    // ===== converters =====
    const int ctid = threadIdx.x - 64;
    if (Cfg::kResidentD) {
      mbar_wait(bar_dfull, 0);
      if (!P.d_prerounded) convert_tf32_inplace(dres, Cfg::kDResBytes, ctid);
      fence_proxy_async_smem();
      mbar_arrive(bar_dconv);
    }
    const int total = my_tiles * KC;
    const int conv_bytes = P.d_prerounded ? Cfg::kFBytes : Cfg::kStageBytes;   // the F tile leads every stage
    for (int it = 0; it < total; ++it) {
      const int s = it % S;
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      mbar_wait(bar_full + 8 * s, ph);
      convert_tf32_inplace(smem + s * Cfg::kStageBytes, conv_bytes, ctid);
      fence_proxy_async_smem();
      mbar_arrive(bar_conv + 8 * s);
    }
  } else if (P.relu_mask && P.accumulate) {
    // ===== epilogue with the fused ReLU backward on top of a running gradient =====
    //   out = F > 0 ? s * acc + dF : 0.   The running-gradient block (32 positions x 32 channels, the HBM stream of
    //   this path) arrives through TMA in the very buffer the result is stored from: lane 0 keeps kOutBufs - 1 block
    //   loads in flight — independent of the accumulator, so they run ahead across tile boundaries — every lane
    //   reads its 128-byte row with conflict-free LDS.128, writes the result back in place, and the buffer goes out
    //   as a plain TMA store.  Only the mask still comes through per-thread loads (F is L2-hot: just staged).
    constexpr int NBUF = Cfg::kOutBufs;
    constexpr int GP = (C / 32) / Cfg::kEpiGroups;          // 32-channel blocks per tile for this epilogue group
    const int ew = warp - 6;
    const int sub = warp & 3;
    const int egrp = ew >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(sub * 32) << 16);
    const float scale = P.gscale ? P.scale * __ldg(P.gscale) : P.scale;
    const uint32_t stg = smem_u32(ostage + ew * (NBUF * 4096));
    const uint32_t bar_q = bar_qall + 8 * (ew * NBUF);
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    // tiles whose 32-position sub-block of this warp starts inside the tensor (a prefix of my_tiles)
    int vt = 0;
    while (vt < my_tiles && ((int64_t)blockIdx.x + (int64_t)vt * gridDim.x) * 128 + sub * 32 < P.HW) ++vt;
    const int nblk = vt * GP;
    auto issue = [&](int m) {                                // lane 0: running-gradient block m -> buffer m % NBUF
      const int ti = m / GP, g = egrp + (m % GP) * Cfg::kEpiGroups;
      const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)ti * gridDim.x) * 128 + sub * 32;
      const uint32_t bq = bar_q + 8 * (m % NBUF);
      mbar_arrive_expect_tx(bq, 4096u);
      tma_load_2d(stg + (uint32_t)(m % NBUF) * 4096u, &tmapO, bq, g * 32, (int)p0);
    };
    if (lane == 0)
      for (int m = 0; m < NBUF - 1 && m < nblk; ++m) issue(m);
    uint32_t v[32];
    int m = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int b = ti % NB;
      const uint32_t bph = (uint32_t)(ti / NB) & 1u;
      const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)ti * gridDim.x) * 128 + sub * 32;
      mbar_wait(bar_accf + 8 * b, bph);
      tc_fence_after();
      if (ti < vt) {
#pragma unroll 1
        for (int gi = 0; gi < GP; ++gi, ++m) {
          const int g = egrp + gi * Cfg::kEpiGroups;
          const int64_t p = p0 + lane;
          float4 f[8];
          if (p < P.HW) {
            const float4* fr = reinterpret_cast<const float4*>(P.F + p * C + g * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __ldg(fr + j);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (lane == 0) {
            tma_store_wait_read<0>();                        // every earlier store has left its buffer
            if (m + NBUF - 1 < nblk) issue(m + NBUF - 1);    // -> the buffer block m - 1 was stored from
          }
          tmem_ld_x32(lane_addr + b * C + g * 32, v);
          tmem_ld_wait();
          mbar_wait(bar_q + 8 * (m % NBUF), (uint32_t)(m / NBUF) & 1u);
          const uint32_t buf = stg + (uint32_t)(m % NBUF) * 4096u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t addr = buf + row_off + ((((uint32_t)j) ^ sw) << 4);
            float4 q;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                         : "r"(addr)
                         : "memory");
            q.x = f[j].x > 0.f ? __fadd_rn(__fmul_rn(scale, __uint_as_float(v[4 * j])), q.x) : 0.f;
            q.y = f[j].y > 0.f ? __fadd_rn(__fmul_rn(scale, __uint_as_float(v[4 * j + 1])), q.y) : 0.f;
            q.z = f[j].z > 0.f ? __fadd_rn(__fmul_rn(scale, __uint_as_float(v[4 * j + 2])), q.z) : 0.f;
            q.w = f[j].w > 0.f ? __fadd_rn(__fmul_rn(scale, __uint_as_float(v[4 * j + 3])), q.w) : 0.f;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(q.x), "f"(q.y), "f"(q.z), "f"(q.w)
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmapO, buf, g * 32, (int)p0);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acce + 8 * b);
    }
    if (lane == 0) tma_store_wait<0>();
  } else {
    // ===== epilogue: TMEM -> registers -> swizzled shared block -> TMA store / reduce-add =====
    const int sub = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(sub * 32) << 16);
    const float scale = P.gscale ? P.scale * __ldg(P.gscale) : P.scale;
    const uint32_t stg = smem_u32(ostage + (warp - 6) * (Cfg::kOutBufs * 4096));
    const int egrp = (warp - 6) >> 2;
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    uint32_t v[32];
    uint32_t nbuf = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int b = ti % NB;
      const uint32_t bph = (uint32_t)(ti / NB) & 1u;
      const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)ti * gridDim.x) * 128 + sub * 32;
      mbar_wait(bar_accf + 8 * b, bph);
      tc_fence_after();
#pragma unroll 1
      for (int g = egrp; g < C / 32; g += Cfg::kEpiGroups) {
        tmem_ld_x32(lane_addr + b * C + g * 32, v);
        tmem_ld_wait();
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = scale * __uint_as_float(v[j]);
        if (P.relu_mask) {
          // Fused ReLU backward (+ running gradient).  This thread's position row, 32 channels = one 128-byte line
          // of F (L2-hot: the tile was just staged) and of dF; all 16 loads are in flight together.  (A variant
          // that re-walks the staged block with 8 lanes per row for coalesced loads was 2x SLOWER on B200: with
          // 4 epilogue warps per SM the loads in flight, not the sectors per request, bound this path.  The
          // planned version takes the mask from the converter warps and dF through a TMA load ring.)
          const int64_t p = p0 + lane;
          if (p < P.HW) {
            const float4* fr = reinterpret_cast<const float4*>(P.F + p * C + g * 32);
            const float4* gr = reinterpret_cast<const float4*>(P.dF + p * C + g * 32);
            float4 f[8], q[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[j] = __ldg(fr + j);
              q[j] = P.accumulate ? gr[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              o[4 * j] = f[j].x > 0.f ? o[4 * j] + q[j].x : 0.f;
              o[4 * j + 1] = f[j].y > 0.f ? o[4 * j + 1] + q[j].y : 0.f;
              o[4 * j + 2] = f[j].z > 0.f ? o[4 * j + 2] + q[j].z : 0.f;
              o[4 * j + 3] = f[j].w > 0.f ? o[4 * j + 3] + q[j].w : 0.f;
            }
          }
        }
        if (lane == 0) tma_store_wait_read<1>();    // the block stored two rounds ago has left shared memory
        __syncwarp();
        const uint32_t buf = stg + (nbuf & 1u) * 4096u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t addr = buf + row_off + ((((uint32_t)j) ^ sw) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o[4 * j]), "f"(o[4 * j + 1]),
                       "f"(o[4 * j + 2]), "f"(o[4 * j + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && p0 < P.HW) {
          if (P.accumulate && !P.relu_mask) tma_reduce_add_2d(&tmapO, buf, g * 32, (int)p0);   // add happens in L2
          else tma_store_2d(&tmapO, buf, g * 32, (int)p0);
          tma_store_commit();
        }
        ++nbuf;
      }
      tc_fence_before();
      mbar_arrive(bar_acce + 8 * b);
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------------
// backward, (HW, C) layout, C = 512: CTA pair (cta_group::2)
// ------------------------------------------------------------------------------------------------------
// The one-CTA kernel above is shared-memory-bandwidth bound at C = 512: per 32-channel K chunk a CTA writes 80 KB
// (F 16 + D 64), its tensor core reads 96 KB (A 4 + B 8 KB per MMA) and the epilogue moves 32 KB, for 1024 cycles of
// MMA.  A pair of CTAs (two SMs of a TPC) computes M = 256 positions per MMA: each CTA keeps its own 128 positions
// of F and only HALF of the B operand (128 of the 256 output channels of each N block), so per CTA the D traffic
// halves (TMA 48 KB, tensor-core reads 64 KB per chunk) and four stages fit where two did.
// Protocol (s = stage): each CTA's TMA fills its full[s]; its converters round the F part and arrive on their OWN
// CTA's conv[s] (count 128, cta scope: a MEMBAR.ALL.CTA).  Round 1 had all 256 converter threads arrive on the LEADER's
// barrier with a cluster-scope release, which ptxas turns into MEMBAR.ALL.GPU + ERRBAR per thread and stage: ncu's
// source page (profiles/r02_ncu_gram_stalls.md) shows the converter warps stalled there (~1 400 of 8 700 samples, the
// largest single stall of the kernel), i.e. one converter warp-group finished a stage only every ~2 500 cycles for
// 1 300 cycles of MMA.  Now ONE otherwise idle thread of the peer CTA (lane 0 of warp 1) relays: it waits on the peer's
// local conv[s] and does the single cluster-scope arrive on the leader's peer[s]; the leader's MMA thread waits for
// conv[s] (its own converters) and peer[s], issues tcgen05.mma.cta_group::2 and commits with a multicast arrive on
// empty[s] of BOTH CTAs; the accumulator-full commit is multicast too; the epilogue warps arrive on their own CTA's
// acc_empty and the same relay thread forwards the peer's to the leader.
// BF16 = true (AST_PREC_BF16): operands in bfloat16, fp32 accumulation.  A K chunk is 64 channels: the F tile lands as
// two fp32 boxes of [128 positions][32 channels]; converter thread p rounds ITS position row of both boxes to 64 bf16
// (cvt.rn) and writes them over the first box's row p — a 128-byte bf16 row occupies exactly the 128-byte fp32 row it
// came from, with the same 128B-swizzle key (p & 7), so the conversion is in place without any cross-thread hazard.
// D arrives ALREADY in bf16 (ast_gram_finalize*, round_out = 2) through a bf16 tensor map: half the L2 traffic and no
// conversion.  Per stage: 8 MMAs of K = 16 (twice the flops of the TF32 stage for the same tensor-core time).
template <bool BF16>
struct Bwd2CtaCfg {
  static constexpr int C = 512;
  static constexpr int kChunk = BF16 ? 64 : 32;                   // channels of K per stage
  static constexpr int kFBytes = 128 * kChunk * 4;                // fp32 landing: this CTA's 128 positions
  static constexpr int kDHalfBytes = 128 * ROW_BYTES;             // 16 KB: 128 output channels x one 128-byte K row
  static constexpr int kStageBytes = kFBytes + 2 * kDHalfBytes;   // 48 KB (TF32) / 64 KB (BF16)
  static constexpr int kStages = BF16 ? 3 : 4;
  static constexpr int kOutBytes = 4 * 2 * 4096;
  static constexpr int kThreads = 320;
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 + 256;
  static_assert((4 * kStages + 4) * 8 <= 256, "barrier area too small");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// Round one position row (64 channels) of a staged BF16 chunk in place: fp32 rows of box 0 / box 1 -> one bf16 row.
__device__ __forceinline__ void convert_bf16_row_inplace(uint8_t* stage, int p) {
  const uint32_t row0 = smem_u32(stage) + (uint32_t)p * 128u;     // box 0, row p (and the bf16 row)
  const uint32_t row1 = row0 + 128u * 128u;                       // box 1, row p (16 KB further)
  const uint32_t sw = (uint32_t)(p & 7);
  uint32_t v[2][8][4];
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t addr = (b ? row1 : row0) + ((((uint32_t)j) ^ sw) << 4);
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(v[b][j][0]), "=r"(v[b][j][1]), "=r"(v[b][j][2]), "=r"(v[b][j][3])
                   : "r"(addr));
    }
  // bf16 16-byte chunk c (channels 8c .. 8c+7) = fp32 chunks 2c, 2c+1 of box c / 4
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int b = c >> 2, j0 = (2 * c) & 7;
    uint32_t o[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[2 * h]) : "f"(__uint_as_float(v[b][j0 + h][1])), "f"(__uint_as_float(v[b][j0 + h][0])));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[2 * h + 1]) : "f"(__uint_as_float(v[b][j0 + h][3])), "f"(__uint_as_float(v[b][j0 + h][2])));
    }
    const uint32_t addr = row0 + ((((uint32_t)c) ^ sw) << 4);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
  }
}

template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
gram_bwd_nhwc_2cta_kernel(const __grid_constant__ CUtensorMap tmapF, const __grid_constant__ CUtensorMap tmapD,
                          const __grid_constant__ CUtensorMap tmapO, const __grid_constant__ BwdNhwcParams P) {
  using Cfg = Bwd2CtaCfg<BF16>;
  constexpr int C = 512, S = Cfg::kStages, KC = C / Cfg::kChunk;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ostage = smem + S * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ostage + Cfg::kOutBytes);
  // full[S] conv[S] empty[S] peer[S] acc_full acc_empty acc_empty_peer
  const uint32_t bar_full = smem_u32(bars), bar_conv = smem_u32(bars + S), bar_empty = smem_u32(bars + 2 * S),
                 bar_peer = smem_u32(bars + 3 * S), bar_accf = smem_u32(bars + 4 * S),
                 bar_acce = smem_u32(bars + 4 * S + 1), bar_accp = smem_u32(bars + 4 * S + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * S + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int my_tiles = (pair < P.n_tiles) ? (P.n_tiles - 1 - pair) / n_pairs + 1 : 0;   // P.n_tiles: 256-position tiles

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmapF);
    prefetch_tmap(&tmapD);
    prefetch_tmap(&tmapO);
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, 128);     // this CTA's converters
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_peer + 8 * s, 1);       // the peer's relay thread (only the leader's copy is used)
    }
    mbar_init(bar_accf, 1);
    mbar_init(bar_acce, 128);               // this CTA's epilogue warps
    mbar_init(bar_accp, 1);                 // the peer's relay thread (leader's copy)
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(smem_u32(tmem_slot), 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();                       // barrier inits + TMEM allocation visible in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (each CTA: its 128 positions of F, its half of both N blocks of D) =====
      int it = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int64_t n0 = ((int64_t)pair + (int64_t)ti * n_pairs) * 256 + rank * 128;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)Cfg::kStageBytes);
          const uint32_t dst = smem_u32(smem + s * Cfg::kStageBytes);
          tma_load_2d(dst, &tmapF, bar_full + 8 * s, kc * Cfg::kChunk, (int)n0);
          if (BF16) tma_load_2d(dst + 128 * ROW_BYTES, &tmapF, bar_full + 8 * s, kc * Cfg::kChunk + BK, (int)n0);
          // D: fp32 rows of 32 k (TF32) or bf16 rows of 64 k (BF16) — 128 bytes either way
          tma_load_2d(dst + Cfg::kFBytes, &tmapD, bar_full + 8 * s, kc * Cfg::kChunk, (int)(rank * 128));
          tma_load_2d(dst + Cfg::kFBytes + Cfg::kDHalfBytes, &tmapD, bar_full + 8 * s, kc * Cfg::kChunk, (int)(256 + rank * 128));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (leader CTA only) =====
      constexpr uint32_t idesc = BF16 ? umma_idesc_bf16(256, 256, 0, 0) : umma_idesc_tf32(256, 256, 0, 0);
      int it = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        mbar_wait(bar_acce, ((uint32_t)ti & 1u) ^ 1u);    // both epilogues have drained the accumulator
        mbar_wait(bar_accp, ((uint32_t)ti & 1u) ^ 1u);
        tc_fence_after();
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(bar_conv + 8 * s, ph);                // my converters ...
          mbar_wait(bar_peer + 8 * s, ph);                // ... and the peer's (relayed)
          tc_fence_after();
          const uint32_t f_base = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t d_base = f_base + Cfg::kFBytes;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint32_t acc = (kc > 0 || k > 0) ? 1u : 0u;
            // one K step = 32 bytes of every 128-byte row: 8 tf32 or 16 bf16
            const uint64_t ad = umma_desc_sw128(f_base + k * 32, 16, 1024);
            const uint64_t bd0 = umma_desc_sw128(d_base + k * 32, 16, 1024);
            const uint64_t bd1 = umma_desc_sw128(d_base + Cfg::kDHalfBytes + k * 32, 16, 1024);
            if (BF16) {
              umma_bf16_2cta(tmem_base, ad, bd0, idesc, acc);
              umma_bf16_2cta(tmem_base + 256, ad, bd1, idesc, acc);
            } else {
              umma_tf32_2cta(tmem_base, ad, bd0, idesc, acc);
              umma_tf32_2cta(tmem_base + 256, ad, bd1, idesc, acc);
            }
          }
          umma_commit_2cta(bar_empty + 8 * s, 3);
        }
        umma_commit_2cta(bar_accf, 3);
      }
    } else if (lane == 0) {
      // ===== relay (peer CTA only): forward "my converters are done with stage s" / "my epilogue has drained the
      // accumulator" to the leader with ONE cluster-scope arrive each =====
      // AST_2CTA_RELAY_RELEASE=1 (P.relay_release): cluster-scope release arrives instead of relaxed ones (A/B switch)
      const uint32_t peer0 = mapa_shared(bar_peer, 0), accp0 = mapa_shared(bar_accp, 0);
      int it = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        if (ti > 0) {
          mbar_wait(bar_acce, ((uint32_t)ti & 1u) ^ 1u);  // my epilogue warps have drained tile ti - 1
          if (P.relay_release) mbar_arrive_cluster(accp0); else mbar_arrive_cluster_relaxed(accp0);
        }
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % S;
          mbar_wait(bar_conv + 8 * s, (uint32_t)(it / S) & 1u);
          if (P.relay_release) mbar_arrive_cluster(peer0 + 8 * s); else mbar_arrive_cluster_relaxed(peer0 + 8 * s);
        }
      }
    }
  } else if (warp < 6) {
    // ===== converters: round this CTA's F tile, arrive on this CTA's conv[s] (cta scope) =====
    const int ctid = threadIdx.x - 64;
    const int total = my_tiles * KC;
    for (int it = 0; it < total; ++it) {
      const int s = it % S;
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      mbar_wait(bar_full + 8 * s, ph);
      if (BF16) convert_bf16_row_inplace(smem + s * Cfg::kStageBytes, ctid);
      else convert_tf32_inplace(smem + s * Cfg::kStageBytes, P.d_prerounded ? Cfg::kFBytes : Cfg::kStageBytes, ctid);
      fence_proxy_async_smem();
      mbar_arrive(bar_conv + 8 * s);
    }
  } else {
    // ===== epilogue (each CTA: its own 128 positions = its own TMEM lanes) =====
    const int sub = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(sub * 32) << 16);
    const float scale = P.gscale ? P.scale * __ldg(P.gscale) : P.scale;
    const uint32_t stg = smem_u32(ostage + (warp - 6) * 8192);
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    uint32_t v[32];
    uint32_t nbuf = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int64_t p0 = ((int64_t)pair + (int64_t)ti * n_pairs) * 256 + rank * 128 + sub * 32;
      mbar_wait(bar_accf, (uint32_t)ti & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < C / 32; ++g) {
        tmem_ld_x32(lane_addr + g * 32, v);
        tmem_ld_wait();
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = scale * __uint_as_float(v[j]);
        if (P.relu_mask) {
          const int64_t p = p0 + lane;
          if (p < P.HW) {
            const float4* fr = reinterpret_cast<const float4*>(P.F + p * C + g * 32);
            const float4* gr = reinterpret_cast<const float4*>(P.dF + p * C + g * 32);
            float4 f[8], q[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[j] = __ldg(fr + j);
              q[j] = P.accumulate ? gr[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              o[4 * j] = f[j].x > 0.f ? o[4 * j] + q[j].x : 0.f;
              o[4 * j + 1] = f[j].y > 0.f ? o[4 * j + 1] + q[j].y : 0.f;
              o[4 * j + 2] = f[j].z > 0.f ? o[4 * j + 2] + q[j].z : 0.f;
              o[4 * j + 3] = f[j].w > 0.f ? o[4 * j + 3] + q[j].w : 0.f;
            }
          }
        }
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        const uint32_t buf = stg + (nbuf & 1u) * 4096u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t addr = buf + row_off + ((((uint32_t)j) ^ sw) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o[4 * j]), "f"(o[4 * j + 1]),
                       "f"(o[4 * j + 2]), "f"(o[4 * j + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && p0 < P.HW) {
          if (P.accumulate && !P.relu_mask) tma_reduce_add_2d(&tmapO, buf, g * 32, (int)p0);
          else tma_store_2d(&tmapO, buf, g * 32, (int)p0);
          tma_store_commit();
        }
        ++nbuf;
      }
      tc_fence_before();
      mbar_arrive(bar_acce);
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  cluster_sync_all();                       // nobody may still be using the pair's TMEM / shared memory
  if (warp == 1) tmem_dealloc_2cta(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
bool gram_tc_supported(int C, int64_t HW, const void* F) {
  if (!(C == 64 || C == 128 || C == 256 || C == 512)) return false;
  if (HW <= 0 || (HW % 4) != 0 || HW > (int64_t)0x7fffff00) return false;
  if (F && (reinterpret_cast<uintptr_t>(F) & 15u)) return false;
  return true;
}

void gram_tc_plan(int C, int64_t HW, int num_sms, GramPlan* plan) {
  const int64_t chunks = (HW + BK - 1) / BK;
  const int64_t units = (C == 64) ? (chunks + 1) / 2 : chunks;
  const int min_units = 4;   // do not split finer than 4 stage fills per CTA
  plan->C = C;
  plan->TR = (C <= 256) ? C : 256;
  if (C <= 256) {
    int64_t ns = units / min_units;
    if (ns > num_sms) ns = num_sms;
    if (ns < 1) ns = 1;
    plan->n_tiles = 1;
    plan->tile_bi[0] = plan->tile_bj[0] = 0;
    plan->part_off[0] = 0;
    plan->part_cnt[0] = (int)ns * (C == 64 ? 2 : 1);
    plan->total_parts = plan->part_cnt[0];
  } else {
    // three 256x256 tiles (0,0) (0,1) (1,1); the off-diagonal one loads twice the bytes per unit -> 2x the CTAs
    int64_t sd = num_sms / 4;
    if (sd > units / min_units) sd = units / min_units;
    if (sd < 1) sd = 1;
    int64_t so = 2 * sd;
    if (so > units) so = units;
    plan->n_tiles = 3;
    const int bi[3] = {0, 0, 1}, bj[3] = {0, 1, 1};
    const int cnt[3] = {(int)sd, (int)so, (int)sd};
    int off = 0;
    for (int t = 0; t < 3; ++t) {
      plan->tile_bi[t] = bi[t];
      plan->tile_bj[t] = bj[t];
      plan->part_off[t] = off;
      plan->part_cnt[t] = cnt[t];
      off += cnt[t];
    }
    plan->total_parts = off;
  }
}

template <int C, int NU, int G, bool NHWC>
static int launch_fwd(const float* F, int64_t HW, int64_t ld, float* partials, const GramPlan& plan,
                      cudaStream_t stream) {
  using Cfg = FwdCfg<C, NU, G>;
  CUtensorMap tmap;
  int rc = NHWC ? make_tmap_nhwc_strips(&tmap, F, C, HW, (C == 64) ? 2 : (C == 128 ? 4 : 8))
                : make_tmap(&tmap, F, C, HW, ld, Cfg::kBoxRows);
  if (rc != AST_OK) return rc;
  FwdParams P;
  P.HW = HW;
  P.n_tiles = plan.n_tiles;
  const int64_t chunks = (HW + BK - 1) / BK;
  P.units_total = (int)((C == 64) ? (chunks + 1) / 2 : chunks);
  int off = 0;
  for (int t = 0; t < plan.n_tiles; ++t) {
    P.cta_off[t] = off;
    P.tile_bi[t] = plan.tile_bi[t];
    P.tile_bj[t] = plan.tile_bj[t];
    P.part_off[t] = plan.part_off[t];
    off += (C == 64) ? plan.part_cnt[t] / 2 : plan.part_cnt[t];
  }
  P.cta_off[plan.n_tiles] = off;
  P.partials = partials;
  static const int noround = (getenv("AST_GRAM_FWD_NOROUND") && atoi(getenv("AST_GRAM_FWD_NOROUND")) == 1) ? 1 : 0;
  P.skip_rounding = noround;
  cudaError_t e = cudaFuncSetAttribute(gram_fwd_tc_kernel<C, NU, G, NHWC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (e != cudaSuccess) {
    set_error("gram_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return AST_ERR_CUDA;
  }
  gram_fwd_tc_kernel<C, NU, G, NHWC><<<off, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(tmap, P);
  return check_launch("gram_fwd_tc");
}

// Pipeline shape of the forward kernel: (units per stage, converter groups).  AST_GRAM_FWD_CFG="NU,G" overrides
// the per-C default for tuning sweeps (tests/tools/gram_sweep.py); every shape gives bit-identical results
// because the accumulation order inside a CTA does not depend on it.
static void fwd_cfg(int C, int* nu, int* g) {
  static int env_nu = -1, env_g = -1;
  if (env_nu < 0) {
    int a = 0, b = 0;
    const char* e = getenv("AST_GRAM_FWD_CFG");
    if (e && sscanf(e, "%d,%d", &a, &b) == 2 && (a == 1 || a == 2) && (b == 1 || b == 2)) {
      env_g = b;
      env_nu = a;
    } else {
      env_g = 0;
      env_nu = 0;
    }
  }
  // Round 1 measured NU and G as neutral (profiles/r01_fwd_cfg_sweep.log) — with G groups splitting EVERY stage.
  // Since round 2 the groups take alternate stages, which is what hides the converter's latency chain; G = 2 is the
  // default, G = 1 the comparison.  (The one launch failure in the round-1 sweep log at NCHW C = 512 G = 2 did not
  // reproduce in two reruns of the same sweep on a B200, profiles/r02_fwd_cfg12_rerun.log; the shape has a test.)
  *nu = env_nu ? env_nu : 1;
  *g = env_g ? env_g : 2;      // two converter groups on alternate stages (round 2)
  if (C >= 256) *nu = 1;   // a stage already holds 32/64 KB
}

int gram_tc_fwd(const float* F, int C, int64_t HW, int64_t ld, int nhwc, float* partials, const GramPlan& plan,
                int num_sms, cudaStream_t stream) {
  (void)num_sms;
  int nu, g;
  fwd_cfg(C, &nu, &g);
#define AST_FWD_CASE(CC, NN, GG)                                                                   \
  if (C == CC && nu == NN && g == GG)                                                              \
    return nhwc ? launch_fwd<CC, NN, GG, true>(F, HW, ld, partials, plan, stream)                  \
                : launch_fwd<CC, NN, GG, false>(F, HW, ld, partials, plan, stream);
  AST_FWD_CASE(64, 1, 1) AST_FWD_CASE(64, 1, 2) AST_FWD_CASE(64, 2, 1) AST_FWD_CASE(64, 2, 2)
  AST_FWD_CASE(128, 1, 1) AST_FWD_CASE(128, 1, 2) AST_FWD_CASE(128, 2, 1) AST_FWD_CASE(128, 2, 2)
  AST_FWD_CASE(256, 1, 1) AST_FWD_CASE(256, 1, 2)
  AST_FWD_CASE(512, 1, 1) AST_FWD_CASE(512, 1, 2)
#undef AST_FWD_CASE
  set_error("gram_tc_fwd: unsupported C=%d", C);
  return AST_ERR_UNSUPPORTED;
}

template <int C>
static int launch_bwd(const float* D, const float* F, int64_t HW, int64_t ld, float scale, const float* gscale,
                      float* dF, int accumulate, int num_sms, cudaStream_t stream) {
  using Cfg = BwdCfg<C>;
  CUtensorMap tmF, tmD;
  // MN-major fp32 operand: tcgen05 only accepts the 128B swizzle with 32-byte atoms (SWIZZLE_128B_BASE32B)
  int rc = make_tmap(&tmF, F, C, HW, ld, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != AST_OK) return rc;
  rc = make_tmap(&tmD, D, C, C, C, Cfg::kDBoxRows);
  if (rc != AST_OK) return rc;
  BwdParams P;
  P.HW = HW;
  P.ld = ld;
  P.n_tiles = (int)((HW + 127) / 128);
  P.scale = scale;
  P.gscale = gscale;
  P.accumulate = accumulate;
  P.dF = dF;
  cudaError_t e = cudaFuncSetAttribute(gram_bwd_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (e != cudaSuccess) {
    set_error("gram_tc_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return AST_ERR_CUDA;
  }
  const int grid = P.n_tiles < num_sms ? P.n_tiles : num_sms;
  gram_bwd_tc_kernel<C><<<grid, 320, Cfg::kSmemBytes, stream>>>(tmF, tmD, P);
  return check_launch("gram_bwd_tc");
}

int gram_tc_bwd(const float* D, const float* F, int C, int64_t HW, int64_t ld, float scale, const float* gscale,
                float* dF, int accumulate, int num_sms, cudaStream_t stream) {
  switch (C) {
    case 64: return launch_bwd<64>(D, F, HW, ld, scale, gscale, dF, accumulate, num_sms, stream);
    case 128: return launch_bwd<128>(D, F, HW, ld, scale, gscale, dF, accumulate, num_sms, stream);
    case 256: return launch_bwd<256>(D, F, HW, ld, scale, gscale, dF, accumulate, num_sms, stream);
    case 512: return launch_bwd<512>(D, F, HW, ld, scale, gscale, dF, accumulate, num_sms, stream);
  }
  set_error("gram_tc_bwd: unsupported C=%d", C);
  return AST_ERR_UNSUPPORTED;
}

template <int C, bool W = false>
static int launch_bwd_nhwc(const float* D, const float* F, int64_t HW, float scale, const float* gscale, float* dF,
                           int accumulate, int d_prerounded, int relu_mask, int num_sms, cudaStream_t stream) {
  using Cfg = BwdNhwcCfg<C, W>;
  CUtensorMap tmF, tmD, tmO;
  int rc = make_tmap(&tmF, F, (uint64_t)HW, C, C, 128);
  if (rc != AST_OK) return rc;
  rc = make_tmap(&tmD, D, C, C, C, Cfg::kDBoxRows);
  if (rc != AST_OK) return rc;
  rc = make_tmap(&tmO, dF, (uint64_t)HW, C, C, 32);
  if (rc != AST_OK) return rc;
  BwdNhwcParams P;
  P.HW = HW;
  P.n_tiles = (int)((HW + 127) / 128);
  P.scale = scale;
  P.gscale = gscale;
  P.accumulate = accumulate;
  P.d_prerounded = d_prerounded;
  P.relu_mask = relu_mask;
  P.F = F;
  P.dF = dF;
  P.C = C;
  P.relay_release = 0;
  cudaError_t e = cudaFuncSetAttribute(gram_bwd_nhwc_tc_kernel<C, W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes);
  if (e != cudaSuccess) {
    set_error("gram_tc_bwd_nhwc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return AST_ERR_CUDA;
  }
  const int grid = P.n_tiles < num_sms ? P.n_tiles : num_sms;
  gram_bwd_nhwc_tc_kernel<C, W><<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(tmF, tmD, tmO, P);
  return check_launch("gram_bwd_nhwc_tc");
}

// D: fp32 (C, C) for TF32, bfloat16 (C, C) for BF16 (as written by ast_gram_finalize* with round_out = 2).
template <bool BF16>
static int launch_bwd_nhwc_2cta(const void* D, const float* F, int64_t HW, float scale, const float* gscale, float* dF,
                                int accumulate, int d_prerounded, int relu_mask, int num_sms, cudaStream_t stream) {
  using Cfg = Bwd2CtaCfg<BF16>;
  constexpr int C = 512;
  CUtensorMap tmF, tmD, tmO;
  int rc = make_tmap(&tmF, F, (uint64_t)HW, C, C, 128);
  if (rc != AST_OK) return rc;
  if (BF16) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
      set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
      return AST_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)C};
    cuuint64_t gstride[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(D), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (bf16 D) failed (CUresult %d)", (int)r);
      return AST_ERR_CUDA;
    }
  } else {
    rc = make_tmap(&tmD, static_cast<const float*>(D), C, C, C, 128);
    if (rc != AST_OK) return rc;
  }
  rc = make_tmap(&tmO, dF, (uint64_t)HW, C, C, 32);
  if (rc != AST_OK) return rc;
  BwdNhwcParams P;
  P.HW = HW;
  P.n_tiles = (int)((HW + 255) / 256);      // 256-position tiles, one per CTA pair
  P.scale = scale;
  P.gscale = gscale;
  P.accumulate = accumulate;
  P.d_prerounded = d_prerounded;
  P.relu_mask = relu_mask;
  P.F = F;
  P.dF = dF;
  P.C = C;
  static const int relay_release = (getenv("AST_2CTA_RELAY_RELEASE") && atoi(getenv("AST_2CTA_RELAY_RELEASE")) == 1) ? 1 : 0;
  P.relay_release = relay_release;
  cudaError_t e = cudaFuncSetAttribute(gram_bwd_nhwc_2cta_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes);
  if (e != cudaSuccess) {
    set_error("gram_tc_bwd_nhwc(2cta): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return AST_ERR_CUDA;
  }
  int pairs = num_sms / 2;
  if (pairs > P.n_tiles) pairs = P.n_tiles;
  if (pairs < 1) pairs = 1;
  gram_bwd_nhwc_2cta_kernel<BF16><<<2 * pairs, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(tmF, tmD, tmO, P);
  return check_launch("gram_bwd_nhwc_2cta");
}

int gram_tc_bwd_nhwc_bf16(const void* D_bf16, const float* F, int C, int64_t HW, float scale, const float* gscale,
                          float* dF, int accumulate, int relu_mask, int num_sms, cudaStream_t stream) {
  if (C != 512) {
    set_error("gram_tc_bwd_nhwc_bf16: BF16 operands are implemented for C = 512 (the tensor-bound width); got C=%d", C);
    return AST_ERR_UNSUPPORTED;
  }
  return launch_bwd_nhwc_2cta<true>(D_bf16, F, HW, scale, gscale, dF, accumulate, 1, relu_mask, num_sms, stream);
}

int gram_tc_bwd_nhwc(const float* D, const float* F, int C, int64_t HW, float scale, const float* gscale, float* dF,
                     int accumulate, int d_prerounded, int relu_mask, int num_sms, cudaStream_t stream) {
  switch (C) {
    case 64: return launch_bwd_nhwc<64>(D, F, HW, scale, gscale, dF, accumulate, d_prerounded, relu_mask, num_sms, stream);
    case 128: return launch_bwd_nhwc<128>(D, F, HW, scale, gscale, dF, accumulate, d_prerounded, relu_mask, num_sms, stream);
    case 256:
      if (relu_mask && accumulate)
        return launch_bwd_nhwc<256, true>(D, F, HW, scale, gscale, dF, accumulate, d_prerounded, relu_mask, num_sms, stream);
      return launch_bwd_nhwc<256>(D, F, HW, scale, gscale, dF, accumulate, d_prerounded, relu_mask, num_sms, stream);
    case 512: {
      // CTA-pair kernel by default; AST_GRAM_BWD_2CTA=0 selects the one-CTA kernel (comparison / fallback)
      static const int use_pair = (getenv("AST_GRAM_BWD_2CTA") && atoi(getenv("AST_GRAM_BWD_2CTA")) == 0) ? 0 : 1;
      if (use_pair)
        return launch_bwd_nhwc_2cta<false>(D, F, HW, scale, gscale, dF, accumulate, d_prerounded, relu_mask, num_sms, stream);
      return launch_bwd_nhwc<512>(D, F, HW, scale, gscale, dF, accumulate, d_prerounded, relu_mask, num_sms, stream);
    }
  }
  set_error("gram_tc_bwd_nhwc: unsupported C=%d", C);
  return AST_ERR_UNSUPPORTED;
}

}  // namespace ast
