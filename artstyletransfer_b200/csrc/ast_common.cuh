// Shared helpers for libast_sm100.so (error reporting, reductions, launch checks).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ast_sm100.h"

namespace ast {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return AST_ERR_CUDA;
  }
  return AST_OK;
}

#define AST_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::ast::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

constexpr int kReduceMaxBlocks = 1024;  // per-block partials kept in the reduce workspace

// Layout of the small reduction workspace (ast_reduce_workspace_bytes()):
//   [0]                unsigned ticket counter (zero at entry, reset to zero by the last block)
//   [64 ...]           double partials[2 * kReduceMaxBlocks]
struct ReduceWs {
  unsigned int ticket;
  unsigned int pad[15];
  double partials[2 * kReduceMaxBlocks];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum (all threads must call; result valid in thread 0).  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v, double* smem32) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem32[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? smem32[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// Deterministic grid reduction: every block deposits `nvals` (1 or 2) doubles; the block that takes the
// last ticket sums all deposits in block order and returns true in thread 0 (totals in out[0..nvals)).
__device__ __forceinline__ bool grid_reduce_last(ReduceWs* ws, const double* vals, int nvals, double* out,
                                                 double* smem32) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    for (int k = 0; k < nvals; ++k) ws->partials[k * kReduceMaxBlocks + blockIdx.x] = vals[k];
    __threadfence();
    const unsigned t = atomicAdd(&ws->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  for (int k = 0; k < nvals; ++k) {
    double acc = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x)
      acc += ((volatile double*)ws->partials)[k * kReduceMaxBlocks + i];
    acc = block_sum(acc, smem32);
    if (threadIdx.x == 0) out[k] = acc;
  }
  if (threadIdx.x == 0) ws->ticket = 0u;  // leave the workspace reusable
  return threadIdx.x == 0;
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace ast
