// Internal interface between the Gram entry points (gram.cu), the exact-fp32 FFMA kernels (gram.cu) and the
// tcgen05 kernels (gram_tc.cu).
#pragma once

#include "ast_common.cuh"

namespace ast {

constexpr int kGramMaxTiles = 3;

// How the split-K partial sums of F F^T are laid out in the workspace, for the finalize kernel.
// The C x C output is cut into (C/TR)^2 square tiles of TR x TR; only tiles with bi <= bj are stored
// (the Gram is symmetric); tile t has part_cnt[t] partial sums, partial p at
//   partials + (part_off[t] + p) * TR * TR   (row-major TR x TR fp32).
struct GramPlan {
  int C;
  int TR;
  int n_tiles;
  int tile_bi[kGramMaxTiles], tile_bj[kGramMaxTiles];
  int part_off[kGramMaxTiles], part_cnt[kGramMaxTiles];
  int total_parts;
};

constexpr size_t kGramWsHeaderBytes = 32768;  // ReduceWs (ticket + per-block loss partials) lives here

static_assert(sizeof(ReduceWs) <= kGramWsHeaderBytes, "reduce header too small");

// gram_tc.cu -------------------------------------------------------------------------------------------
// True when the tcgen05 path handles this problem: C in {64,128,256,512}, HW % 4 == 0 (TMA row pitch is a
// multiple of 16 bytes) and a 16-byte aligned base pointer.
bool gram_tc_supported(int C, int64_t HW, const void* F);
// Plans the split-K decomposition for `num_sms` SMs (no device access).
void gram_tc_plan(int C, int64_t HW, int num_sms, GramPlan* plan);
// nhwc = 0: F is (C, HW) with row pitch ld (torch NCHW).  nhwc = 1: F is (HW, C) dense (torch channels_last).
int gram_tc_fwd(const float* F, int C, int64_t HW, int64_t ld, int nhwc, float* partials, const GramPlan& plan,
                int num_sms, cudaStream_t stream);
int gram_tc_bwd(const float* D, const float* F, int C, int64_t HW, int64_t ld, float scale, const float* gscale,
                float* dF, int accumulate, int num_sms, cudaStream_t stream);

// (HW, C) layout: dF[p, c] (+)= scale * sum_k F[p, k] D[c, k]
// d_prerounded: D already holds TF32-representable values (ast_gram_mse_fwd_nhwc / ast_gram_finalize with round_out),
// so the kernel's converter warps skip it (for C >= 256 D is 2/3 .. 4/5 of every staged chunk).
// relu_mask: also apply the backward of the ReLU whose output F is: dF = (accumulate ? dF + v : v) * (F > 0).
int gram_tc_bwd_nhwc(const float* D, const float* F, int C, int64_t HW, float scale, const float* gscale, float* dF,
                     int accumulate, int d_prerounded, int relu_mask, int num_sms, cudaStream_t stream);

// BF16 operands (AST_PREC_BF16), C = 512 only: D is bfloat16 (C, C) as written by the finalize kernels with round_out = 2.
int gram_tc_bwd_nhwc_bf16(const void* D_bf16, const float* F, int C, int64_t HW, float scale, const float* gscale,
                          float* dF, int accumulate, int relu_mask, int num_sms, cudaStream_t stream);

}  // namespace ast
