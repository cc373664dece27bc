// Bandwidth-bound pieces of the Gatys loss: content MSE and total variation, forward + backward.
// Replaces torch.nn.MSELoss at neural_style_transfer.py:95 and math_utils.total_variation
// (math_utils.py:37-41) together with their autograd backward graphs.
//
// All kernels are streaming: 16-byte vector loads through the read-only path, one pass over the
// data, fp64 block/grid reductions made deterministic by a ticketed last-block sum.
#include "ast_common.cuh"

namespace ast {

constexpr int kThreads = 256;

static inline int grid_for(int64_t work_items, int per_block) {
  int64_t b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > kReduceMaxBlocks) b = kReduceMaxBlocks;
  return (int)b;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------ content MSE forward
__global__ void __launch_bounds__(kThreads) mse_fwd_kernel(const float* __restrict__ X, const float* __restrict__ T,
                                                          int64_t n, int vec_ok, float scale,
                                                          float* __restrict__ loss, ReduceWs* ws) {
  __shared__ double red[32];
  float acc = 0.f;
  double acc_d = 0.0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* X4 = reinterpret_cast<const float4*>(X);
    const float4* T4 = reinterpret_cast<const float4*>(T);
    int iter = 0;
    for (int64_t i = tid; i < n4; i += nthreads) {
      const float4 a = ldg_stream(X4 + i), b = ldg_stream(T4 + i);
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      if ((++iter & 63) == 0) { acc_d += (double)acc; acc = 0.f; }  // bound fp32 accumulation length
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nthreads) {
    const float d = X[i] - T[i];
    acc += d * d;
  }
  acc_d += (double)acc;
  double v = block_sum(acc_d, red);
  double total;
  if (grid_reduce_last(ws, &v, 1, &total, red)) *loss = (float)(total * (double)scale);
}

// ------------------------------------------------------------------ content MSE backward
__global__ void __launch_bounds__(kThreads) mse_bwd_kernel(const float* __restrict__ X, const float* __restrict__ T,
                                                          int64_t n, int vec_ok, float scale,
                                                          const float* __restrict__ gscale, float* __restrict__ dX,
                                                          int accumulate, int relu_mask) {
  // relu_mask: X is a ReLU output; its backward (X > 0) is applied to the accumulated gradient in the same pass
  if (gscale) scale *= __ldg(gscale);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* X4 = reinterpret_cast<const float4*>(X);
    const float4* T4 = reinterpret_cast<const float4*>(T);
    float4* D4 = reinterpret_cast<float4*>(dX);
    for (int64_t i = tid; i < n4; i += nthreads) {
      const float4 a = ldg_stream(X4 + i), b = ldg_stream(T4 + i);
      float4 g = make_float4(scale * (a.x - b.x), scale * (a.y - b.y), scale * (a.z - b.z), scale * (a.w - b.w));
      if (accumulate) {
        const float4 o = D4[i];
        g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w;
      }
      if (relu_mask) {
        g.x = a.x > 0.f ? g.x : 0.f; g.y = a.y > 0.f ? g.y : 0.f;
        g.z = a.z > 0.f ? g.z : 0.f; g.w = a.w > 0.f ? g.w : 0.f;
      }
      D4[i] = g;
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nthreads) {
    float g = scale * (X[i] - T[i]);
    if (accumulate) g += dX[i];
    dX[i] = (relu_mask && !(X[i] > 0.f)) ? 0.f : g;
  }
}

// ------------------------------------------------------------------ total variation forward
// One thread per 4 consecutive pixels of a row (W % 4 == 0 and 16B-aligned rows), else scalar.
__global__ void __launch_bounds__(kThreads) tv_fwd_kernel(const float* __restrict__ Y, int C, int H, int W, int vec_ok,
                                                         float* __restrict__ sums2, float* __restrict__ tv,
                                                         ReduceWs* ws) {
  __shared__ double red[32];
  float sx = 0.f, sy = 0.f;
  double sxd = 0.0, syd = 0.0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t rows = (int64_t)C * H;
  if (vec_ok) {
    // one image row per block iteration, threads stride along the row: no per-element 64-bit divisions
    const int w4 = W >> 2;
    int iter = 0;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
      const int y = (int)(row % H);
      const float* r = Y + row * W;
#pragma unroll 4
      for (int xq = threadIdx.x; xq < w4; xq += blockDim.x) {
        const float* p = r + (xq << 2);
        const float4 a = ldg_stream(reinterpret_cast<const float4*>(p));
        sx += fabsf(a.x - a.y) + fabsf(a.y - a.z) + fabsf(a.z - a.w);
        if (xq + 1 < w4) sx += fabsf(a.w - __ldg(p + 4));
        if (y + 1 < H) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p + W));
          sy += fabsf(a.x - b.x) + fabsf(a.y - b.y) + fabsf(a.z - b.z) + fabsf(a.w - b.w);
        }
        if ((++iter & 31) == 0) { sxd += sx; syd += sy; sx = sy = 0.f; }
      }
    }
  } else {
    const int64_t n = rows * W;
    for (int64_t i = tid; i < n; i += nthreads) {
      const int64_t row = i / W;
      const int x = (int)(i - row * W);
      const int y = (int)(row % H);
      const float a = Y[i];
      if (x + 1 < W) sx += fabsf(a - Y[i + 1]);
      if (y + 1 < H) sy += fabsf(a - Y[i + W]);
    }
  }
  sxd += sx; syd += sy;
  double v[2];
  v[0] = block_sum(sxd, red);
  v[1] = block_sum(syd, red);
  double tot[2];
  if (grid_reduce_last(ws, v, 2, tot, red)) {
    sums2[0] = (float)tot[0];
    sums2[1] = (float)tot[1];
    if (tv) {
      const double mx = tot[0] / ((double)C * H * (W - 1)), my = tot[1] / ((double)C * (H - 1) * W);
      *tv = (float)(mx * mx + my * my);
    }
  }
}

__device__ __forceinline__ float sgn(float v) { return (float)((v > 0.f) - (v < 0.f)); }

// ------------------------------------------------------------------ total variation backward (gather form)
__global__ void __launch_bounds__(kThreads) tv_bwd_kernel(const float* __restrict__ Y, int C, int H, int W,
                                                         const float* __restrict__ sums2, float kx, float ky,
                                                         const float* __restrict__ gscale, float* __restrict__ dY,
                                                         int accumulate, int vec_ok, int r0, int r1) {
  // rows [r0, r1) of every plane (the whole image: 0, H); neighbours outside the range are still read
  const float gs = gscale ? __ldg(gscale) : 1.f;
  const float cx = kx * sums2[0] * gs, cy = ky * sums2[1] * gs;
  const int hb = r1 - r0;
  const int64_t n = (int64_t)C * hb * W;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok) {
    // 4 consecutive pixels of a row per thread: 16-byte loads of the row, the row above and the row below
    const int w4 = W >> 2;
    const int64_t rows = (int64_t)C * hb;
    for (int64_t rr = blockIdx.x; rr < rows; rr += gridDim.x) {
     const int y = r0 + (int)(rr % hb);
     const int64_t row = (rr / hb) * H + y;
#pragma unroll 2
     for (int xq = threadIdx.x; xq < w4; xq += blockDim.x) {
      const float* p = Y + row * W + (xq << 2);
      const float4 a = __ldg(reinterpret_cast<const float4*>(p));
      const float v[6] = {xq > 0 ? __ldg(p - 1) : 0.f, a.x, a.y, a.z, a.w, xq + 1 < w4 ? __ldg(p + 4) : 0.f};
      float g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float t = 0.f;
        if (k < 3 || xq + 1 < w4) t += cx * sgn(v[k + 1] - v[k + 2]);
        if (k > 0 || xq > 0) t -= cx * sgn(v[k] - v[k + 1]);
        g[k] = t;
      }
      if (y + 1 < H) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p + W));
        g[0] += cy * sgn(a.x - b.x); g[1] += cy * sgn(a.y - b.y); g[2] += cy * sgn(a.z - b.z); g[3] += cy * sgn(a.w - b.w);
      }
      if (y > 0) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(p - W));
        g[0] -= cy * sgn(u.x - a.x); g[1] -= cy * sgn(u.y - a.y); g[2] -= cy * sgn(u.z - a.z); g[3] -= cy * sgn(u.w - a.w);
      }
      float4* o = reinterpret_cast<float4*>(dY + row * W + (xq << 2));
      float4 r = make_float4(g[0], g[1], g[2], g[3]);
      if (accumulate) {
        const float4 old = *o;
        r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
      }
      *o = r;
     }
    }
    return;
  }
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += nthreads) {
    const int64_t rr = j / W;
    const int x = (int)(j - rr * W);
    const int y = r0 + (int)(rr % hb);
    const int64_t i = ((rr / hb) * H + y) * W + x;
    const float a = __ldg(Y + i);
    float g = 0.f;
    if (x + 1 < W) g += cx * sgn(a - __ldg(Y + i + 1));
    if (x > 0) g -= cx * sgn(__ldg(Y + i - 1) - a);
    if (y + 1 < H) g += cy * sgn(a - __ldg(Y + i + W));
    if (y > 0) g -= cy * sgn(__ldg(Y + i - W) - a);
    dY[i] = accumulate ? dY[i] + g : g;
  }
}

__global__ void level_combine_kernel(const float* __restrict__ style_mse, int n_style,
                                     const float* __restrict__ content, const float* __restrict__ tv, float cw,
                                     float sw, float tvw, float* __restrict__ out4) {
  // neural_style_transfer.py:100-110: style_loss = sum(mse_k) / n; total = cw*content + sw*style + tvw*tv
  float s = 0.f;
  for (int k = 0; k < n_style; ++k) s += style_mse[k];
  s /= (float)n_style;
  const float c = *content, t = *tv;
  out4[0] = cw * c + sw * s + tvw * t;
  out4[1] = c;
  out4[2] = s;
  out4[3] = t;
}

}  // namespace ast

using namespace ast;

extern "C" int ast_level_combine(const float* style_mse, int n_style, const float* content, const float* tv,
                                 float content_weight, float style_weight, float tv_weight, float* out4,
                                 void* stream) {
  AST_REQUIRE(style_mse && content && tv && out4, AST_ERR_INVALID, "ast_level_combine: null pointer");
  AST_REQUIRE(n_style > 0, AST_ERR_INVALID, "ast_level_combine: n_style must be positive");
  level_combine_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(style_mse, n_style, content, tv, content_weight,
                                                          style_weight, tv_weight, out4);
  return check_launch("ast_level_combine");
}

extern "C" size_t ast_reduce_workspace_bytes(void) { return sizeof(ReduceWs); }

extern "C" int ast_mse_fwd(const float* X, const float* T, int64_t n, float scale, float* loss, void* ws,
                           size_t ws_bytes, void* stream) {
  AST_REQUIRE(X && T && loss && ws, AST_ERR_INVALID, "ast_mse_fwd: null pointer");
  AST_REQUIRE(n > 0, AST_ERR_INVALID, "ast_mse_fwd: n must be positive (got %lld)", (long long)n);
  AST_REQUIRE(ws_bytes >= sizeof(ReduceWs), AST_ERR_WORKSPACE, "ast_mse_fwd: workspace %zu < %zu", ws_bytes,
              sizeof(ReduceWs));
  const int vec_ok = aligned16(X) && aligned16(T);
  const int grid = grid_for(n, kThreads * 16);
  mse_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(X, T, n, vec_ok, scale, loss, (ReduceWs*)ws);
  return check_launch("ast_mse_fwd");
}

extern "C" int ast_mse_bwd(const float* X, const float* T, int64_t n, float scale, const float* gscale, float* dX,
                           int accumulate, int relu_mask, void* stream) {
  AST_REQUIRE(X && T && dX, AST_ERR_INVALID, "ast_mse_bwd: null pointer");
  AST_REQUIRE(n > 0, AST_ERR_INVALID, "ast_mse_bwd: n must be positive (got %lld)", (long long)n);
  const int vec_ok = aligned16(X) && aligned16(T) && aligned16(dX);
  int64_t blocks = (n + (int64_t)kThreads * 8 - 1) / ((int64_t)kThreads * 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  mse_bwd_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(X, T, n, vec_ok, scale, gscale, dX, accumulate,
                                                                     relu_mask ? 1 : 0);
  return check_launch("ast_mse_bwd");
}

extern "C" int ast_tv_fwd(const float* Y, int C, int H, int W, float* sums2, float* tv, void* ws, size_t ws_bytes,
                          void* stream) {
  AST_REQUIRE(Y && sums2 && ws, AST_ERR_INVALID, "ast_tv_fwd: null pointer");
  AST_REQUIRE(C > 0 && H > 1 && W > 1, AST_ERR_INVALID, "ast_tv_fwd: bad shape %dx%dx%d", C, H, W);
  AST_REQUIRE(ws_bytes >= sizeof(ReduceWs), AST_ERR_WORKSPACE, "ast_tv_fwd: workspace %zu < %zu", ws_bytes,
              sizeof(ReduceWs));
  const int vec_ok = aligned16(Y) && (W % 4 == 0);
  const int64_t n = (int64_t)C * H * W;
  const int grid = grid_for(n, kThreads * 16);
  tv_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(Y, C, H, W, vec_ok, sums2, tv, (ReduceWs*)ws);
  return check_launch("ast_tv_fwd");
}

extern "C" int ast_tv_bwd(const float* Y, int C, int H, int W, const float* sums2, float kx, float ky,
                          const float* gscale, float* dY, int accumulate, void* stream) {
  AST_REQUIRE(Y && sums2 && dY, AST_ERR_INVALID, "ast_tv_bwd: null pointer");
  AST_REQUIRE(C > 0 && H > 0 && W > 0, AST_ERR_INVALID, "ast_tv_bwd: bad shape %dx%dx%d", C, H, W);
  const int64_t n = (int64_t)C * H * W;
  int64_t blocks = (n + (int64_t)kThreads * 4 - 1) / ((int64_t)kThreads * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const int vec_ok = aligned16(Y) && aligned16(dY) && (W % 4 == 0);
  tv_bwd_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(Y, C, H, W, sums2, kx, ky, gscale, dY, accumulate,
                                                                    vec_ok, 0, H);
  return check_launch("ast_tv_bwd");
}

extern "C" int ast_tv_bwd_rows(const float* Y, int C, int H, int W, int r0, int r1, const float* sums2, float kx,
                               float ky, const float* gscale, float* dY, int accumulate, void* stream) {
  AST_REQUIRE(Y && sums2 && dY, AST_ERR_INVALID, "ast_tv_bwd_rows: null pointer");
  AST_REQUIRE(C > 0 && H > 0 && W > 0, AST_ERR_INVALID, "ast_tv_bwd_rows: bad shape %dx%dx%d", C, H, W);
  AST_REQUIRE(0 <= r0 && r0 < r1 && r1 <= H, AST_ERR_INVALID, "ast_tv_bwd_rows: rows [%d, %d) of %d", r0, r1, H);
  const int64_t n = (int64_t)C * (r1 - r0) * W;
  int64_t blocks = (n + (int64_t)kThreads * 4 - 1) / ((int64_t)kThreads * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const int vec_ok = aligned16(Y) && aligned16(dY) && (W % 4 == 0);
  tv_bwd_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(Y, C, H, W, sums2, kx, ky, gscale, dY, accumulate,
                                                                    vec_ok, r0, r1);
  return check_launch("ast_tv_bwd_rows");
}
