// Bandwidth-bound glue between the cuDNN convolutions of the VGG19 feature path (neural_nets.py:53-68 of the
// reference; the convolutions themselves stay on cuDNN), for activations held as (H, W, C) row-major — torch
// channels_last, the layout cuDNN's TF32 kernels run in without NCHW<->NHWC transposes.
//
// Replaces, per closure and level, what torch launches around the convolutions:
//   bias add + ReLU(inplace)            (2 passes)        -> bias_relu_kernel              (1 pass, in place)
//   max_pool2d_with_indices (+int64 indices)              -> maxpool2x2_kernel             (no indices)
//   max_pool2d backward + threshold_backward (2 kernels)  -> maxpool2x2_relu_bwd_kernel    (1 pass)
//   threshold_backward                                    -> relu_bwd_kernel               (in place)
//   NCHW<->NHWC copies of the 3-channel image / gradient  -> chw_to_hwc_kernel / hwc_to_chw_kernel
//
// All kernels: 16-byte vector accesses along C (C % 4 == 0), streaming loads, grid sized to a multiple of the
// SM count, grid-stride loops.  Max-pool ties go to the first element in (row, column) scan order, which is what
// torch's max_pool2d does (strict > against the running maximum).
#include "ast_common.cuh"

namespace ast {

constexpr int kGlueThreads = 256;

static inline int glue_grid(int64_t items) {
  // 148 SMs x 8 resident 256-thread CTAs
  int64_t b = (items + kGlueThreads - 1) / kGlueThreads;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// y[p, c] = max(y[p, c] + bias[c], 0), in place.  n4 = n_pos * C / 4, c4 = C / 4.
__global__ void __launch_bounds__(kGlueThreads) bias_relu_kernel(float4* __restrict__ y,
                                                                const float4* __restrict__ bias, int64_t n4,
                                                                int c4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = y[i];
    const float4 b = __ldg(bias + (int)(i % c4));
    v.x = fmaxf(v.x + b.x, 0.f);
    v.y = fmaxf(v.y + b.y, 0.f);
    v.z = fmaxf(v.z + b.z, 0.f);
    v.w = fmaxf(v.w + b.w, 0.f);
    y[i] = v;
  }
}

// g[i] = r[i] > 0 ? g[i] : 0, in place (r = the ReLU's output).
__global__ void __launch_bounds__(kGlueThreads) relu_bwd_kernel(float4* __restrict__ g, const float4* __restrict__ r,
                                                               int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = g[i];
    const float4 a = ldg_stream(r + i);
    v.x = a.x > 0.f ? v.x : 0.f;
    v.y = a.y > 0.f ? v.y : 0.f;
    v.z = a.z > 0.f ? v.z : 0.f;
    v.w = a.w > 0.f ? v.w : 0.f;
    g[i] = v;
  }
}

__device__ __forceinline__ float4 max4(const float4& a, const float4& b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// y[ho, wo, c] = max over the 2x2 window of x (floor mode: an odd last row / column is dropped).
__global__ void __launch_bounds__(kGlueThreads) maxpool2x2_kernel(const float4* __restrict__ x, int c4, int H, int W,
                                                                 float4* __restrict__ y) {
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t total = (int64_t)Ho * Wo * c4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % c4);
    const int64_t q = i / c4;
    const int wo = (int)(q % Wo), ho = (int)(q / Wo);
    const float4* p = x + ((int64_t)(2 * ho) * W + 2 * wo) * c4 + c;
    const float4 a = ldg_stream(p), b = ldg_stream(p + c4);
    const float4 d = ldg_stream(p + (int64_t)W * c4), e = ldg_stream(p + (int64_t)W * c4 + c4);
    stg_stream(y + i, max4(max4(a, b), max4(d, e)));
  }
}

// Route gy through the window's arg-max (first in scan order) and through the ReLU that produced x:
//   gx[h, w, c] = (x is the window's first maximum and x > 0) ? gy[h/2, w/2, c] : 0
// One thread per window and 4 channels; a dropped odd row / column gets zeros.
__device__ __forceinline__ void route1(float a, float b, float d, float e, float g, bool relu_mask, float& oa,
                                       float& ob, float& od, float& oe) {
  // torch keeps the first element that is strictly greater than the running maximum
  int k = 0;
  float m = a;
  if (b > m) { m = b; k = 1; }
  if (d > m) { m = d; k = 2; }
  if (e > m) { m = e; k = 3; }
  const float v = (!relu_mask || m > 0.f) ? g : 0.f;
  oa = k == 0 ? v : 0.f;
  ob = k == 1 ? v : 0.f;
  od = k == 2 ? v : 0.f;
  oe = k == 3 ? v : 0.f;
}

__global__ void __launch_bounds__(kGlueThreads) maxpool2x2_relu_bwd_kernel(const float4* __restrict__ gy,
                                                                          const float4* __restrict__ x, int c4,
                                                                          int H, int W, int relu_mask,
                                                                          float4* __restrict__ gx) {
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t total = (int64_t)Ho * Wo * c4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool mask = relu_mask != 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % c4);
    const int64_t q = i / c4;
    const int wo = (int)(q % Wo), ho = (int)(q / Wo);
    const int64_t o = ((int64_t)(2 * ho) * W + 2 * wo) * c4 + c;
    const int64_t row = (int64_t)W * c4;
    const float4 a = ldg_stream(x + o), b = ldg_stream(x + o + c4);
    const float4 d = ldg_stream(x + o + row), e = ldg_stream(x + o + row + c4);
    const float4 g = ldg_stream(gy + i);
    float4 oa, ob, od, oe;
    route1(a.x, b.x, d.x, e.x, g.x, mask, oa.x, ob.x, od.x, oe.x);
    route1(a.y, b.y, d.y, e.y, g.y, mask, oa.y, ob.y, od.y, oe.y);
    route1(a.z, b.z, d.z, e.z, g.z, mask, oa.z, ob.z, od.z, oe.z);
    route1(a.w, b.w, d.w, e.w, g.w, mask, oa.w, ob.w, od.w, oe.w);
    stg_stream(gx + o, oa);
    stg_stream(gx + o + c4, ob);
    stg_stream(gx + o + row, od);
    stg_stream(gx + o + row + c4, oe);
  }
  // odd tails (floor-mode pooling never read them): zero gradient
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (W & 1) {
    const int64_t n = (int64_t)H * c4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const int c = (int)(i % c4);
      const int64_t h = i / c4;
      gx[(h * W + (W - 1)) * c4 + c] = z;
    }
  }
  if (H & 1) {
    const int64_t n = (int64_t)W * c4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      gx[(int64_t)(H - 1) * W * c4 + i] = z;
  }
}

// (C, HW) planar <-> (HW, C) interleaved for small C (the 3-channel image and its gradient).  `plane` is the
// element stride between channel planes of the planar side (>= HW: a row band of a taller image).
__global__ void __launch_bounds__(kGlueThreads) chw_to_hwc_kernel(const float* __restrict__ x, int C, int64_t HW,
                                                                 int64_t plane, float* __restrict__ y) {
  const int64_t n = HW * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = (int)(i % C);
    const int64_t p = i / C;
    y[i] = __ldg(x + (int64_t)c * plane + p);
  }
}

__global__ void __launch_bounds__(kGlueThreads) hwc_to_chw_kernel(const float* __restrict__ x, int C, int64_t HW,
                                                                 int64_t plane, float* __restrict__ y,
                                                                 int accumulate) {
  const int64_t n = HW * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t p = i % HW;
    const int c = (int)(i / HW);
    const float v = __ldg(x + p * C + c);
    float* o = y + (int64_t)c * plane + p;
    *o = accumulate ? *o + v : v;
  }
}

// unprepare_img (neural_style_transfer.py:388-393 of the reference) on the device: planar (3, H*W) "x*255 - mean"
// image -> interleaved (H*W, 3) in [0,1].  The reference adds a float64 mean to the float32 array in place (numpy
// computes the sum in double and rounds to float), then divides the float32 array by 255: y = fl32(fl32(x + mean) / 255).
// Four pixels per thread: three coalesced 16-byte plane loads, three 16-byte stores of the 12 interleaved floats.
__global__ void __launch_bounds__(kGlueThreads) unprepare_hwc_kernel(const float* __restrict__ x, int64_t HW, double m0,
                                                                    double m1, double m2, float* __restrict__ y) {
  const int64_t n4 = HW >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = ldg_stream(reinterpret_cast<const float4*>(x) + i);
    const float4 b = ldg_stream(reinterpret_cast<const float4*>(x + HW) + i);
    const float4 c = ldg_stream(reinterpret_cast<const float4*>(x + 2 * HW) + i);
#define AST_UNPREP(v, m) ((float)((double)(v) + (m)) / 255.0f)
    float4* o = reinterpret_cast<float4*>(y) + 3 * i;
    stg_stream(o, make_float4(AST_UNPREP(a.x, m0), AST_UNPREP(b.x, m1), AST_UNPREP(c.x, m2), AST_UNPREP(a.y, m0)));
    stg_stream(o + 1, make_float4(AST_UNPREP(b.y, m1), AST_UNPREP(c.y, m2), AST_UNPREP(a.z, m0), AST_UNPREP(b.z, m1)));
    stg_stream(o + 2, make_float4(AST_UNPREP(c.z, m2), AST_UNPREP(a.w, m0), AST_UNPREP(b.w, m1), AST_UNPREP(c.w, m2)));
  }
  // tail pixels (HW % 4) by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (HW & 3)) {
    const int64_t p = (n4 << 2) + threadIdx.x;
    y[3 * p] = AST_UNPREP(x[p], m0);
    y[3 * p + 1] = AST_UNPREP(x[HW + p], m1);
    y[3 * p + 2] = AST_UNPREP(x[2 * HW + p], m2);
  }
#undef AST_UNPREP
}

}  // namespace ast

using namespace ast;

extern "C" int ast_bias_relu_nhwc(float* y, const float* bias, int C, int64_t n_pos, void* stream) {
  AST_REQUIRE(y && bias, AST_ERR_INVALID, "ast_bias_relu_nhwc: null pointer");
  AST_REQUIRE(C > 0 && C % 4 == 0 && n_pos > 0, AST_ERR_INVALID, "ast_bias_relu_nhwc: bad shape C=%d n_pos=%lld", C,
              (long long)n_pos);
  AST_REQUIRE(al16(y) && al16(bias), AST_ERR_INVALID, "ast_bias_relu_nhwc: pointers must be 16-byte aligned");
  const int64_t n4 = n_pos * (C / 4);
  bias_relu_kernel<<<glue_grid(n4), kGlueThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float4*>(y), reinterpret_cast<const float4*>(bias), n4, C / 4);
  return check_launch("bias_relu");
}

extern "C" int ast_relu_bwd(float* g, const float* r, int64_t n, void* stream) {
  AST_REQUIRE(g && r, AST_ERR_INVALID, "ast_relu_bwd: null pointer");
  AST_REQUIRE(n > 0 && n % 4 == 0, AST_ERR_INVALID, "ast_relu_bwd: n must be a positive multiple of 4 (got %lld)",
              (long long)n);
  AST_REQUIRE(al16(g) && al16(r), AST_ERR_INVALID, "ast_relu_bwd: pointers must be 16-byte aligned");
  relu_bwd_kernel<<<glue_grid(n / 4), kGlueThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float4*>(g), reinterpret_cast<const float4*>(r), n / 4);
  return check_launch("relu_bwd");
}

extern "C" int ast_maxpool2x2_nhwc(const float* x, int C, int H, int W, float* y, void* stream) {
  AST_REQUIRE(x && y, AST_ERR_INVALID, "ast_maxpool2x2_nhwc: null pointer");
  AST_REQUIRE(C > 0 && C % 4 == 0 && H >= 2 && W >= 2, AST_ERR_INVALID, "ast_maxpool2x2_nhwc: bad shape C=%d H=%d W=%d",
              C, H, W);
  AST_REQUIRE(al16(x) && al16(y), AST_ERR_INVALID, "ast_maxpool2x2_nhwc: pointers must be 16-byte aligned");
  const int64_t total = (int64_t)(H / 2) * (W / 2) * (C / 4);
  maxpool2x2_kernel<<<glue_grid(total), kGlueThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), C / 4, H, W, reinterpret_cast<float4*>(y));
  return check_launch("maxpool2x2");
}

extern "C" int ast_maxpool2x2_bwd_nhwc(const float* gy, const float* x, int C, int H, int W, int relu_mask, float* gx,
                                       void* stream) {
  AST_REQUIRE(gy && x && gx, AST_ERR_INVALID, "ast_maxpool2x2_bwd_nhwc: null pointer");
  AST_REQUIRE(C > 0 && C % 4 == 0 && H >= 2 && W >= 2, AST_ERR_INVALID,
              "ast_maxpool2x2_bwd_nhwc: bad shape C=%d H=%d W=%d", C, H, W);
  AST_REQUIRE(al16(gy) && al16(x) && al16(gx), AST_ERR_INVALID, "ast_maxpool2x2_bwd_nhwc: pointers must be 16-byte aligned");
  const int64_t total = (int64_t)(H / 2) * (W / 2) * (C / 4);
  maxpool2x2_relu_bwd_kernel<<<glue_grid(total), kGlueThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(gy), reinterpret_cast<const float4*>(x), C / 4, H, W, relu_mask ? 1 : 0,
      reinterpret_cast<float4*>(gx));
  return check_launch("maxpool2x2_relu_bwd");
}

extern "C" int ast_chw_to_hwc(const float* x, int C, int64_t HW, int64_t plane_stride, float* y, void* stream) {
  AST_REQUIRE(x && y, AST_ERR_INVALID, "ast_chw_to_hwc: null pointer");
  AST_REQUIRE(C > 0 && C <= 16 && HW > 0 && plane_stride >= HW, AST_ERR_INVALID,
              "ast_chw_to_hwc: bad shape C=%d HW=%lld plane=%lld", C, (long long)HW, (long long)plane_stride);
  chw_to_hwc_kernel<<<glue_grid(HW * C), kGlueThreads, 0, (cudaStream_t)stream>>>(x, C, HW, plane_stride, y);
  return check_launch("chw_to_hwc");
}

extern "C" int ast_hwc_to_chw(const float* x, int C, int64_t HW, float* y, int64_t plane_stride, int accumulate,
                              void* stream) {
  AST_REQUIRE(x && y, AST_ERR_INVALID, "ast_hwc_to_chw: null pointer");
  AST_REQUIRE(C > 0 && C <= 16 && HW > 0 && plane_stride >= HW, AST_ERR_INVALID,
              "ast_hwc_to_chw: bad shape C=%d HW=%lld plane=%lld", C, (long long)HW, (long long)plane_stride);
  hwc_to_chw_kernel<<<glue_grid(HW * C), kGlueThreads, 0, (cudaStream_t)stream>>>(x, C, HW, plane_stride, y,
                                                                                 accumulate ? 1 : 0);
  return check_launch("hwc_to_chw");
}

extern "C" int ast_unprepare_hwc(const float* x_chw, int64_t HW, double mean0, double mean1, double mean2, float* y_hwc,
                                 void* stream) {
  AST_REQUIRE(x_chw && y_hwc, AST_ERR_INVALID, "ast_unprepare_hwc: null pointer");
  AST_REQUIRE(HW > 0 && HW % 4 == 0, AST_ERR_INVALID, "ast_unprepare_hwc: H*W must be a positive multiple of 4 (got %lld)",
              (long long)HW);
  AST_REQUIRE(al16(x_chw) && al16(y_hwc), AST_ERR_INVALID, "ast_unprepare_hwc: pointers must be 16-byte aligned");
  unprepare_hwc_kernel<<<glue_grid(HW / 4), kGlueThreads, 0, (cudaStream_t)stream>>>(x_chw, HW, mean0, mean1, mean2, y_hwc);
  return check_launch("unprepare_hwc");
}
