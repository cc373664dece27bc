// Bicubic resampling kernels (Keys cubic, A = -0.75, per-tap index clamp, no antialias).
//
//   K4  bicubic_down2x      : the in-loop pyramid step, F.interpolate(x, (H/2, W/2), 'bicubic') at
//                             neural_style_transfer.py:173-176 for even H, W.  src = 2d + 0.5 exactly, so the
//                             taps are the constants [-3/32, 19/32, 19/32, -3/32] at rows/cols 2d-1..2d+2.
//                             Input tile (+1/+2 halo) staged in shared memory from 16-byte global loads.
//   K5  bicubic_down2x_adj  : its exact transpose in gather form (replaces the atomicAdd scatter of
//                             upsample_bicubic2d_backward): each thread owns a 2x4 block of the input gradient
//                             and reads the 3x4 output-gradient neighbourhood; deterministic.
//   K6  bicubic_resize(_adj): arbitrary in/out sizes (odd pyramid levels, and cv2.resize(..., INTER_CUBIC) at
//                             neural_style_transfer.py:226, :304, :427), CHW or HWC.
//
// All are HBM-bound: algorithmic traffic 4*C*(Hin*Win + Hout*Wout) bytes.
#include "ast_common.cuh"

namespace ast {

__device__ __constant__ float kTap[4] = {-0.09375f, 0.59375f, 0.59375f, -0.09375f};

// ---------------------------------------------------------------------------------------------------------
// K4: exact 2x down.  Block = 256 threads, output tile 16 rows x 128 cols per plane.
// Shared tile: 34 input rows x 258 input cols (cols 2*e0-1 .. 2*e0+256), local col = gcol - (2*e0-1) so that
// every thread's first tap sits on an even (8-byte aligned) local column -> conflict-free LDS.64.
// ---------------------------------------------------------------------------------------------------------
constexpr int D2_TH = 16, D2_TW = 128;
constexpr int D2_SROWS = 2 * D2_TH + 2;        // 34 input rows
constexpr int D2_HALF = D2_TW + 4;             // 132 entries per parity per row
// Shared tile, split by column parity so that both the 16-byte global loads (stored as two 8-byte pairs) and
// the per-output tap reads (stride-1 across the warp) are bank-conflict free:
//   E[r][k+2] = x[row r][2*(e0+k)]      (even global columns), k = -2 .. 129
//   O[r][k+1] = x[row r][2*(e0+k) - 1]  (odd  global columns), k = -1 .. 130
// Output column e = e0+le reads O[le], E[le], O[le+1], E[le+1] = global columns 2e-1 .. 2e+2.

// TV = true (ast_bicubic_down2x_tv): the block also sums |x[r,c] - x[r,c+1]| and |x[r,c] - x[r+1,c]| over the 32 x 256
// input pixels it owns — they are all in the staged tile, neighbours included — so total_variation(x)
// (math_utils.py:37-41) costs no pass of its own over the image the pyramid step reads anyway.  Per-block fp64 partials
// go to the workspace; the last block (ticket) adds them in block order: deterministic.
struct TvTileWs {
  unsigned int ticket;
  unsigned int pad[15];
  double partials[1];       // [2][n_blocks]
};

template <bool TV>
__global__ void __launch_bounds__(256) down2x_kernel(const float* __restrict__ x, int H, int W,
                                                    float* __restrict__ y, int vec_ok, int C, float* __restrict__ sums2,
                                                    float* __restrict__ tv, TvTileWs* __restrict__ ws) {
  __shared__ __align__(16) float tileE[D2_SROWS * D2_HALF];
  __shared__ __align__(16) float tileO[D2_SROWS * D2_HALF];
  const int Ho = H >> 1, Wo = W >> 1;
  const int e0 = blockIdx.x * D2_TW, d0 = blockIdx.y * D2_TH;
  const float* xp = x + (size_t)blockIdx.z * H * W;
  float* yp = y + (size_t)blockIdx.z * Ho * Wo;
  const int gr0 = 2 * d0 - 1;  // global row of local row 0

  if (vec_ok) {
    // aligned float4 columns 2*e0-4 .. 2*e0+259 -> 66 float4 per row (W % 4 == 0, rows 16-byte aligned)
    constexpr int V = 66;
    constexpr int kIters = (D2_SROWS * V + 255) / 256;   // 9
    float4 vals[kIters];
    // all of the thread's 16-byte loads are issued before the first shared-memory store: 9 loads in flight per
    // thread instead of 1 (the load -> store dependency otherwise serialises them and HBM latency bounds the kernel)
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int idx = threadIdx.x + it * 256;
      const int r = idx / V, v = idx - r * V;
      const int gr = min(max(gr0 + min(r, D2_SROWS - 1), 0), H - 1);
      const int c4 = 2 * e0 - 4 + 4 * v;
      const float* row = xp + (size_t)gr * W;
      if (c4 < 0) {
        const float e = __ldg(row);
        vals[it] = make_float4(e, e, e, e);
      } else if (c4 < W) {
        vals[it] = ldg_stream(reinterpret_cast<const float4*>(row + c4));
      } else {
        const float e = __ldg(row + W - 1);
        vals[it] = make_float4(e, e, e, e);
      }
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int idx = threadIdx.x + it * 256;
      if (idx < D2_SROWS * V) {
        const int r = idx / V, v = idx - r * V;
        // columns c4, c4+2 are even -> E[2v-2 .. 2v-1] (+2); columns c4+1, c4+3 odd -> O[2v-1 .. 2v] (+1)
        *reinterpret_cast<float2*>(tileE + r * D2_HALF + 2 * v) = make_float2(vals[it].x, vals[it].z);
        *reinterpret_cast<float2*>(tileO + r * D2_HALF + 2 * v) = make_float2(vals[it].y, vals[it].w);
      }
    }
  } else {
    constexpr int SC = 2 * D2_TW + 2;  // local columns 0..257 <-> global 2*e0-1 .. 2*e0+256
    for (int idx = threadIdx.x; idx < D2_SROWS * SC; idx += 256) {
      const int r = idx / SC, c = idx - r * SC;
      const int gr = min(max(gr0 + r, 0), H - 1);
      const int gc = min(max(2 * e0 - 1 + c, 0), W - 1);
      const float val = __ldg(xp + (size_t)gr * W + gc);
      if ((c & 1) == 0) tileO[r * D2_HALF + (c >> 1) + 1] = val;   // unclamped column is odd
      else              tileE[r * D2_HALF + (c >> 1) + 2] = val;   // unclamped column is even
    }
  }
  __syncthreads();

  // thread -> output column e0 + (tid % 128), 8 consecutive output rows starting at d0 + 8*(tid / 128)
  const int le = threadIdx.x & (D2_TW - 1);
  const int lr0 = (threadIdx.x >> 7) * 8;
  const int e = e0 + le;
  const float w0 = kTap[0], w1 = kTap[1];
  float h[18];  // horizontally filtered local input rows 2*lr0 .. 2*lr0+17
#pragma unroll
  for (int r = 0; r < 18; ++r) {
    const float* pe = tileE + (2 * lr0 + r) * D2_HALF + le + 2;
    const float* po = tileO + (2 * lr0 + r) * D2_HALF + le + 1;
    // same association as upsample_bicubic2d's cubic_interp1d: x0*c0 + x1*c1 + x2*c2 + x3*c3
    h[r] = ((po[0] * w0 + pe[0] * w1) + po[1] * w1) + pe[1] * w0;
  }
  if (e < Wo) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int d = d0 + lr0 + k;
      if (d < Ho) yp[(size_t)d * Wo + e] = ((h[2 * k] * w0 + h[2 * k + 1] * w1) + h[2 * k + 2] * w1) + h[2 * k + 3] * w0;
    }
  }
  if (TV) {
    // this thread: input columns 2e, 2e+1 (E[le+2], O[le+2]; right neighbour E[le+3]) x input rows 2(d0+lr0) .. +15 =
    // local rows 2*lr0+1 .. 2*lr0+16 (row below: +1, staged up to local row 33).  Out-of-image neighbours were
    // staged as clamped copies (difference 0); out-of-image OWNERS (partial tiles) are masked.
    __shared__ double red[32];
    __shared__ bool is_last;
    float sx = 0.f, sy = 0.f;
    const bool c0ok = 2 * e < W, c1ok = 2 * e + 1 < W;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int l = 2 * lr0 + 1 + r;
      if (gr0 + l < H) {
        const float e0v = tileE[l * D2_HALF + le + 2], o0v = tileO[l * D2_HALF + le + 2], e1v = tileE[l * D2_HALF + le + 3];
        const float edn = tileE[(l + 1) * D2_HALF + le + 2], odn = tileO[(l + 1) * D2_HALF + le + 2];
        if (c0ok) { sx += fabsf(e0v - o0v); sy += fabsf(e0v - edn); }
        if (c1ok) { sx += fabsf(o0v - e1v); sy += fabsf(o0v - odn); }
      }
    }
    const int nb = gridDim.x * gridDim.y * gridDim.z;
    const int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const double bx = block_sum((double)sx, red), by = block_sum((double)sy, red);
    if (threadIdx.x == 0) {
      ws->partials[bid] = bx;
      ws->partials[nb + bid] = by;
      __threadfence();
      is_last = atomicAdd(&ws->ticket, 1u) == (unsigned)nb - 1u;
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      double ax = 0.0, ay = 0.0;
      for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        ax += ((volatile double*)ws->partials)[i];
        ay += ((volatile double*)ws->partials)[nb + i];
      }
      ax = block_sum(ax, red);
      ay = block_sum(ay, red);
      if (threadIdx.x == 0) {
        sums2[0] = (float)ax;
        sums2[1] = (float)ay;
        if (tv) {
          const double mx = ax / ((double)C * H * (W - 1)), my = ay / ((double)C * (H - 1) * W);
          *tv = (float)(mx * mx + my * my);
        }
        ws->ticket = 0u;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K5: transpose of K4, gather form.  gx (C, H, W) with H = 2*Ho, W = 2*Wo; gy (C, Ho, Wo).
// Per axis (m outputs, n = 2m inputs):
//   out[2p]   = w1*g[p] + w3*g[p-1]   (+ w0*g[0]   folded from the clamped tap -1 when p == 0)
//   out[2p+1] = w2*g[p] + w0*g[p+1]   (+ w3*g[m-1] folded from the clamped tap n  when p == m-1)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void adj_axis_weights(int p, int m, float& c_m1, float& c_0e, float& c_0o, float& c_p1) {
  // contributions to (even, odd) input samples of position p from g[p-1], g[p], g[p+1]
  c_m1 = (p >= 1) ? kTap[3] : 0.f;                       // even <- g[p-1]
  c_0e = kTap[1] + ((p == 0) ? kTap[0] : 0.f);            // even <- g[p]
  c_0o = kTap[2] + ((p == m - 1) ? kTap[3] : 0.f);        // odd  <- g[p]
  c_p1 = (p + 1 <= m - 1) ? kTap[0] : 0.f;                // odd  <- g[p+1]
}

// Thread = output-gradient row p, FOUR output-gradient columns q0..q0+3 -> writes 2 rows x 8 floats of gx from a
// 3 x 6 neighbourhood of gy (16-byte loads for the aligned middle four); block = 64 column quads x 4 rows.
__global__ void __launch_bounds__(256) down2x_adj_kernel(const float* __restrict__ gy, int Ho, int Wo,
                                                        float* __restrict__ gx, int accumulate, int vec_ok) {
  const int W = 2 * Wo;
  const int q0 = 4 * (blockIdx.x * 64 + (threadIdx.x & 63));
  const int p = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (q0 >= Wo || p >= Ho) return;
  const float* gp = gy + (size_t)blockIdx.z * Ho * Wo;
  float* xp = gx + (size_t)blockIdx.z * (2 * Ho) * W;
  // vertical weights for rows 2p (even) and 2p+1 (odd)
  float vy_m1, vy_0e, vy_0o, vy_p1;
  adj_axis_weights(p, Ho, vy_m1, vy_0e, vy_0o, vy_p1);
  // g rows p-1, p, p+1 at columns q0-1 .. q0+4
  float ra[6], rb[6], rd[6];
  const bool vec_in = vec_ok && (q0 + 3 < Wo);
#pragma unroll
  for (int rr = 0; rr < 3; ++rr) {
    float* dst = rr == 0 ? ra : (rr == 1 ? rb : rd);
    const int row = p - 1 + rr;
    if (row < 0 || row >= Ho) {
#pragma unroll
      for (int k = 0; k < 6; ++k) dst[k] = 0.f;
      continue;
    }
    const float* g = gp + (size_t)row * Wo;
    dst[0] = (q0 >= 1) ? __ldg(g + q0 - 1) : 0.f;
    if (vec_in) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g + q0));
      dst[1] = v.x; dst[2] = v.y; dst[3] = v.z; dst[4] = v.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[1 + k] = (q0 + k < Wo) ? __ldg(g + q0 + k) : 0.f;
    }
    dst[5] = (q0 + 4 < Wo) ? __ldg(g + q0 + 4) : 0.f;
  }
  float ge[6], go[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    ge[k] = vy_m1 * ra[k] + vy_0e * rb[k];
    go[k] = vy_0o * rb[k] + vy_p1 * rd[k];
  }
  float r0[8], r1[8];
#pragma unroll
  for (int s = 0; s < 4; ++s) {  // the four output-gradient columns q0+s
    float hx_m1, hx_0e, hx_0o, hx_p1;
    adj_axis_weights(q0 + s, Wo, hx_m1, hx_0e, hx_0o, hx_p1);
    r0[2 * s] = hx_m1 * ge[s] + hx_0e * ge[s + 1];
    r0[2 * s + 1] = hx_0o * ge[s + 1] + hx_p1 * ge[s + 2];
    r1[2 * s] = hx_m1 * go[s] + hx_0e * go[s + 1];
    r1[2 * s + 1] = hx_0o * go[s + 1] + hx_p1 * go[s + 2];
  }
  float* o0 = xp + (size_t)(2 * p) * W + 2 * q0;
  float* o1 = o0 + W;
  if (vec_in) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 a = make_float4(r0[4 * h], r0[4 * h + 1], r0[4 * h + 2], r0[4 * h + 3]);
      float4 b = make_float4(r1[4 * h], r1[4 * h + 1], r1[4 * h + 2], r1[4 * h + 3]);
      if (accumulate) {
        const float4 pa = *reinterpret_cast<const float4*>(o0 + 4 * h), pb = *reinterpret_cast<const float4*>(o1 + 4 * h);
        a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
        b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
      }
      *reinterpret_cast<float4*>(o0 + 4 * h) = a;
      *reinterpret_cast<float4*>(o1 + 4 * h) = b;
    }
  } else {
    const int ncol = 2 * min(4, Wo - q0);
    for (int k = 0; k < ncol; ++k) {
      o0[k] = accumulate ? o0[k] + r0[k] : r0[k];
      o1[k] = accumulate ? o1[k] + r1[k] : r1[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K6: general-ratio bicubic gather.
// ---------------------------------------------------------------------------------------------------------
struct AxisMap {
  float scale_f;    // torch: (float)in / out
  double scale_d;   // cv2 : 1.0 / ((double)out / in)
};

__host__ __device__ inline AxisMap make_axis(int n_in, int n_out) {
  AxisMap a;
  a.scale_f = (float)n_in / (float)n_out;
  a.scale_d = 1.0 / ((double)n_out / (double)n_in);
  return a;
}

__device__ __forceinline__ void src_index(const AxisMap& a, int d, int coord_mode, int& ix, float& t) {
  float src;
  if (coord_mode == AST_COORD_TORCH) {
    src = a.scale_f * ((float)d + 0.5f) - 0.5f;   // area_pixel_compute_source_index (cubic: no clamp at 0)
  } else {
    src = (float)(((double)d + 0.5) * a.scale_d - 0.5);  // cv2 resizeGeneric: fx = (float)((dx+0.5)*scale - 0.5)
  }
  const float fl = floorf(src);
  ix = (int)fl;
  t = src - fl;
}

__device__ __forceinline__ void cubic_coeffs(float t, int coord_mode, float w[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f;
  w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  w[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
  const float u = 1.0f - t;
  w[2] = ((A + 2.0f) * u - (A + 3.0f)) * u * u + 1.0f;
  if (coord_mode == AST_COORD_TORCH) {
    const float x3 = u + 1.0f;
    w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;   // get_cubic_upsample_coefficients
  } else {
    w[3] = 1.0f - w[0] - w[1] - w[2];                               // cv2 interpolateCubic
  }
}

// One thread per output pixel (x fastest).  CHW: plane = blockIdx.z.  HWC: thread loops the C channels.
__global__ void __launch_bounds__(256) resize_kernel(const float* __restrict__ x, int C, int Hin, int Win,
                                                    float* __restrict__ y, int Hout, int Wout, int layout,
                                                    int coord_mode, AxisMap ay, AxisMap ax) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  if (ox >= Wout) return;
  int iy, ixx;
  float ty, tx;
  src_index(ay, oy, coord_mode, iy, ty);
  src_index(ax, ox, coord_mode, ixx, tx);
  float wy[4], wx[4];
  cubic_coeffs(ty, coord_mode, wy);
  cubic_coeffs(tx, coord_mode, wx);
  int ry[4], rx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ry[k] = min(max(iy - 1 + k, 0), Hin - 1);
    rx[k] = min(max(ixx - 1 + k, 0), Win - 1);
  }
  if (layout == AST_LAYOUT_CHW) {
    const float* xp = x + (size_t)blockIdx.z * Hin * Win;
    float acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* row = xp + (size_t)ry[i] * Win;
      acc[i] = ((__ldg(row + rx[0]) * wx[0] + __ldg(row + rx[1]) * wx[1]) + __ldg(row + rx[2]) * wx[2]) +
               __ldg(row + rx[3]) * wx[3];
    }
    y[((size_t)blockIdx.z * Hout + oy) * Wout + ox] = ((acc[0] * wy[0] + acc[1] * wy[1]) + acc[2] * wy[2]) + acc[3] * wy[3];
  } else {
    for (int c = 0; c < C; ++c) {
      float acc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* row = x + (size_t)ry[i] * Win * C + c;
        acc[i] = ((__ldg(row + (size_t)rx[0] * C) * wx[0] + __ldg(row + (size_t)rx[1] * C) * wx[1]) +
                  __ldg(row + (size_t)rx[2] * C) * wx[2]) + __ldg(row + (size_t)rx[3] * C) * wx[3];
      }
      y[((size_t)oy * Wout + ox) * C + c] = ((acc[0] * wy[0] + acc[1] * wy[1]) + acc[2] * wy[2]) + acc[3] * wy[3];
    }
  }
}

// Weight with which output sample d touches input sample i along one axis (sum over clamped taps).
__device__ __forceinline__ float axis_weight(const AxisMap& a, int d, int i, int n_in, int coord_mode) {
  int ix;
  float t;
  src_index(a, d, coord_mode, ix, t);
  float w[4];
  cubic_coeffs(t, coord_mode, w);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (min(max(ix - 1 + k, 0), n_in - 1) == i) s += w[k];
  return s;
}

__device__ __forceinline__ void adj_range(int i, int n_in, int n_out, int& lo, int& hi) {
  // conservative range of output samples whose (clamped) 4 taps can reach input sample i
  const double inv = (double)n_out / (double)n_in;
  lo = (int)floor(((double)i - 2.0 + 0.5) * inv - 0.5) - 1;
  hi = (int)ceil(((double)i + 2.0 + 0.5) * inv - 0.5) + 1;
  if (i == 0) lo = 0;
  if (i == n_in - 1) hi = n_out - 1;
  lo = max(lo, 0);
  hi = min(hi, n_out - 1);
}

// Transpose of resize_kernel (CHW) in gather form: one thread per INPUT pixel.
__global__ void __launch_bounds__(256) resize_adj_kernel(const float* __restrict__ gy, int Hin, int Win, int Hout,
                                                        int Wout, float* __restrict__ gx, int accumulate,
                                                        int coord_mode, AxisMap ay, AxisMap ax) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= Win) return;
  int dlo, dhi, elo, ehi;
  adj_range(i, Hin, Hout, dlo, dhi);
  adj_range(j, Win, Wout, elo, ehi);
  const float* gp = gy + (size_t)blockIdx.z * Hout * Wout;
  float acc = 0.f;
  for (int d = dlo; d <= dhi; ++d) {
    const float wy = axis_weight(ay, d, i, Hin, coord_mode);
    if (wy == 0.f) continue;
    float rowacc = 0.f;
    for (int e = elo; e <= ehi; ++e) {
      const float wx = axis_weight(ax, e, j, Win, coord_mode);
      if (wx != 0.f) rowacc += wx * __ldg(gp + (size_t)d * Wout + e);
    }
    acc += wy * rowacc;
  }
  float* o = gx + ((size_t)blockIdx.z * Hin + i) * Win + j;
  *o = accumulate ? *o + acc : acc;
}

}  // namespace ast

using namespace ast;

static inline bool aligned16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int ast_bicubic_down2x(const float* x, int C, int H, int W, float* y, void* stream) {
  AST_REQUIRE(x && y, AST_ERR_INVALID, "ast_bicubic_down2x: null pointer");
  AST_REQUIRE(C > 0 && H >= 2 && W >= 2, AST_ERR_INVALID, "ast_bicubic_down2x: bad shape %dx%dx%d", C, H, W);
  AST_REQUIRE((H % 2 == 0) && (W % 2 == 0), AST_ERR_UNSUPPORTED,
              "ast_bicubic_down2x: H and W must be even (got %dx%d); use ast_bicubic_resize", H, W);
  AST_REQUIRE(C <= 65535, AST_ERR_INVALID, "ast_bicubic_down2x: too many planes");
  const int Ho = H / 2, Wo = W / 2;
  const int vec_ok = aligned16p(x) && (W % 4 == 0);
  dim3 grid((Wo + D2_TW - 1) / D2_TW, (Ho + D2_TH - 1) / D2_TH, C);
  down2x_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, H, W, y, vec_ok, C, nullptr, nullptr, nullptr);
  return check_launch("ast_bicubic_down2x");
}

extern "C" size_t ast_bicubic_down2x_tv_workspace_bytes(int C, int H, int W) {
  if (C <= 0 || H < 2 || W < 2) return 0;
  const size_t nb = (size_t)((W / 2 + D2_TW - 1) / D2_TW) * ((H / 2 + D2_TH - 1) / D2_TH) * C;
  return 64 + 2 * nb * sizeof(double);
}

extern "C" int ast_bicubic_down2x_tv(const float* x, int C, int H, int W, float* y, float* sums2, float* tv, void* ws,
                                     size_t ws_bytes, void* stream) {
  AST_REQUIRE(x && y && sums2 && ws, AST_ERR_INVALID, "ast_bicubic_down2x_tv: null pointer");
  AST_REQUIRE(C > 0 && H >= 2 && W >= 2, AST_ERR_INVALID, "ast_bicubic_down2x_tv: bad shape %dx%dx%d", C, H, W);
  AST_REQUIRE((H % 2 == 0) && (W % 2 == 0), AST_ERR_UNSUPPORTED,
              "ast_bicubic_down2x_tv: H and W must be even (got %dx%d)", H, W);
  AST_REQUIRE(C <= 65535, AST_ERR_INVALID, "ast_bicubic_down2x_tv: too many planes");
  AST_REQUIRE(ws_bytes >= ast_bicubic_down2x_tv_workspace_bytes(C, H, W) && aligned16p(ws), AST_ERR_WORKSPACE,
              "ast_bicubic_down2x_tv: workspace %zu < %zu", ws_bytes, ast_bicubic_down2x_tv_workspace_bytes(C, H, W));
  const int Ho = H / 2, Wo = W / 2;
  const int vec_ok = aligned16p(x) && (W % 4 == 0);
  dim3 grid((Wo + D2_TW - 1) / D2_TW, (Ho + D2_TH - 1) / D2_TH, C);
  down2x_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, H, W, y, vec_ok, C, sums2, tv, (TvTileWs*)ws);
  return check_launch("ast_bicubic_down2x_tv");
}

extern "C" int ast_bicubic_down2x_adj(const float* gy, int C, int H, int W, float* gx, int accumulate,
                                      void* stream) {
  AST_REQUIRE(gy && gx, AST_ERR_INVALID, "ast_bicubic_down2x_adj: null pointer");
  AST_REQUIRE(C > 0 && H >= 2 && W >= 2, AST_ERR_INVALID, "ast_bicubic_down2x_adj: bad shape %dx%dx%d", C, H, W);
  AST_REQUIRE((H % 2 == 0) && (W % 2 == 0), AST_ERR_UNSUPPORTED,
              "ast_bicubic_down2x_adj: H and W must be even (got %dx%d); use ast_bicubic_resize_adj", H, W);
  AST_REQUIRE(C <= 65535, AST_ERR_INVALID, "ast_bicubic_down2x_adj: too many planes");
  const int Ho = H / 2, Wo = W / 2;
  // 16-byte accesses need gy rows (Wo % 4 == 0) and gx rows (W % 8 == 0 follows) aligned
  const int vec_ok = aligned16p(gx) && aligned16p(gy) && (Wo % 4 == 0);
  dim3 grid((unsigned)(((Wo + 3) / 4 + 63) / 64), (unsigned)((Ho + 3) / 4), C);
  AST_REQUIRE(grid.y <= 65535, AST_ERR_UNSUPPORTED, "ast_bicubic_down2x_adj: image too tall (%d rows)", H);
  down2x_adj_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gy, Ho, Wo, gx, accumulate, vec_ok);
  return check_launch("ast_bicubic_down2x_adj");
}

extern "C" int ast_bicubic_resize(const float* x, int C, int Hin, int Win, float* y, int Hout, int Wout, int layout,
                                  int coord_mode, void* stream) {
  AST_REQUIRE(x && y, AST_ERR_INVALID, "ast_bicubic_resize: null pointer");
  AST_REQUIRE(C > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, AST_ERR_INVALID, "ast_bicubic_resize: bad shape");
  AST_REQUIRE(layout == AST_LAYOUT_CHW || layout == AST_LAYOUT_HWC, AST_ERR_INVALID, "ast_bicubic_resize: bad layout %d", layout);
  AST_REQUIRE(coord_mode == AST_COORD_TORCH || coord_mode == AST_COORD_CV2, AST_ERR_INVALID, "ast_bicubic_resize: bad coord_mode %d", coord_mode);
  AST_REQUIRE(Hout <= 65535 && C <= 65535, AST_ERR_UNSUPPORTED, "ast_bicubic_resize: Hout/C exceed grid limits");
  dim3 grid((Wout + 255) / 256, Hout, layout == AST_LAYOUT_CHW ? C : 1);
  resize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, C, Hin, Win, y, Hout, Wout, layout, coord_mode,
                                                        make_axis(Hin, Hout), make_axis(Win, Wout));
  return check_launch("ast_bicubic_resize");
}

extern "C" int ast_bicubic_resize_adj(const float* gy, int C, int Hin, int Win, int Hout, int Wout, float* gx,
                                      int accumulate, int coord_mode, void* stream) {
  AST_REQUIRE(gy && gx, AST_ERR_INVALID, "ast_bicubic_resize_adj: null pointer");
  AST_REQUIRE(C > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, AST_ERR_INVALID, "ast_bicubic_resize_adj: bad shape");
  AST_REQUIRE(coord_mode == AST_COORD_TORCH || coord_mode == AST_COORD_CV2, AST_ERR_INVALID, "ast_bicubic_resize_adj: bad coord_mode %d", coord_mode);
  AST_REQUIRE(Hin <= 65535 && C <= 65535, AST_ERR_UNSUPPORTED, "ast_bicubic_resize_adj: Hin/C exceed grid limits");
  dim3 grid((Win + 255) / 256, Hin, C);
  resize_adj_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gy, Hin, Win, Hout, Wout, gx, accumulate, coord_mode,
                                                            make_axis(Hin, Hout), make_axis(Win, Wout));
  return check_launch("ast_bicubic_resize_adj");
}
