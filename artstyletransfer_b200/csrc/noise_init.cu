// K7: fused structured-noise init (neural_style_transfer.py:265-362, gaussian_mask :396-418).
//
// One pass over the top-level (H, W, 3) HWC image.  Per 16x32-pixel tile a CTA
//   1. stages the content tile with a 3-pixel REFLECT_101 halo in shared memory,
//   2. forms the clipped 5x5 Sobel magnitude (cv2.Sobel ksize=5, CV_64F; :331-337) with a 1-pixel halo,
//   3. applies the (numerically 3-tap) GaussianBlur((101,101), 0.2) of :340 and r = 5*nf/(5+s) (:342-343),
//   4. accumulates the noise levels — bicubic-upsampled low-res grid (cv2 INTER_CUBIC, :304) times the separable
//      Gaussian envelope (:404-413), float32 accumulation per level exactly as the reference's in-place += —
//   5. blends (1-r)*content + r*noise in fp64 and writes float32 (:353-357), or noise*0.5 for 'random' (:350-352).
// HBM traffic ~ 24*H*W bytes (read content once, write init once); the low-res grids and 1-D Gaussian vectors
// are L1/L2 resident, except a granularity -1 grid (full resolution), which adds 12*H*W.
#include "ast_common.cuh"

namespace ast {

constexpr int NI_TH = 16, NI_TW = 32;
constexpr int NI_CH = 3;
constexpr int NI_CROWS = NI_TH + 6, NI_CCOLS = NI_TW + 6;   // content tile with halo 3
constexpr int NI_SROWS = NI_TH + 2, NI_SCOLS = NI_TW + 2;   // sobel tile with halo 1

struct NoiseParams {
  ast_noise_level lv[AST_NOISE_MAX_LEVELS];
  int n_levels;
  int H, W;
  int mode;
  int use_gradient_map;
  double noise_factor;
  double blur_w0, blur_w1;
};

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    else p = 2 * (n - 1) - p;
  }
  return p;
}

__device__ __forceinline__ void cv_coord(int d, double scale, int& ix, float w[4]) {
  float fx = (float)(((double)d + 0.5) * scale - 0.5);
  const float fl = floorf(fx);
  ix = (int)fl;
  const float t = fx - fl;
  const float A = -0.75f;
  const float x0 = t + 1.0f;
  w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  w[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
  const float u = 1.0f - t;
  w[2] = ((A + 2.0f) * u - (A + 3.0f)) * u * u + 1.0f;
  w[3] = 1.0f - w[0] - w[1] - w[2];
}

__global__ void __launch_bounds__(256) noise_init_kernel(const float* __restrict__ content,
                                                        float* __restrict__ out,
                                                        const __grid_constant__ NoiseParams P) {
  __shared__ float cs[NI_CROWS][NI_CCOLS * NI_CH];
  __shared__ double sm[NI_SROWS][NI_SCOLS * NI_CH];
  const int H = P.H, W = P.W;
  const int x0 = blockIdx.x * NI_TW, y0 = blockIdx.y * NI_TH;
  const bool need_map = (P.mode == AST_INIT_CONTENT_NOISE) && P.use_gradient_map;

  if (P.mode == AST_INIT_CONTENT_NOISE) {
    for (int idx = threadIdx.x; idx < NI_CROWS * NI_CCOLS; idx += blockDim.x) {
      const int r = idx / NI_CCOLS, c = idx - r * NI_CCOLS;
      const int gy = reflect101(y0 - 3 + r, H), gx = reflect101(x0 - 3 + c, W);
      const float* src = content + ((size_t)gy * W + gx) * NI_CH;
      cs[r][c * NI_CH + 0] = __ldg(src + 0);
      cs[r][c * NI_CH + 1] = __ldg(src + 1);
      cs[r][c * NI_CH + 2] = __ldg(src + 2);
    }
    __syncthreads();
  }
  if (need_map) {
    const double kd[5] = {-1.0, -2.0, 0.0, 2.0, 1.0};
    const double ks[5] = {1.0, 4.0, 6.0, 4.0, 1.0};
    for (int idx = threadIdx.x; idx < NI_SROWS * NI_SCOLS * NI_CH; idx += blockDim.x) {
      const int ch = idx % NI_CH;
      const int pc = (idx / NI_CH) % NI_SCOLS;
      const int pr = idx / (NI_CH * NI_SCOLS);
      // sobel position (y0-1+pr, x0-1+pc) -> content tile rows pr..pr+4, cols pc..pc+4
      double sx = 0.0, sy = 0.0;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        double rowd = 0.0, rows = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const double v = (double)cs[pr + i][(pc + j) * NI_CH + ch];
          rowd += kd[j] * v;
          rows += ks[j] * v;
        }
        sx += ks[i] * rowd;   // dx=1: derivative along x, smoothing along y
        sy += kd[i] * rows;   // dy=1
      }
      double m = sqrt(sx * sx + sy * sy);
      m = fmin(fmax(m, 0.0), 100.0);
      sm[pr][pc * NI_CH + ch] = m;
    }
    __syncthreads();
  }

  for (int pix = threadIdx.x; pix < NI_TH * NI_TW; pix += blockDim.x) {
    const int ly = pix / NI_TW, lx = pix - ly * NI_TW;
    const int y = y0 + ly, x = x0 + lx;
    if (y >= H || x >= W) continue;
    float acc[NI_CH] = {0.f, 0.f, 0.f};
    for (int l = 0; l < P.n_levels; ++l) {
      const ast_noise_level& L = P.lv[l];
      double env = 0.0;
      if (L.kind != 2) {
        const double gn = (__ldg(L.gy + y) * __ldg(L.gx + x)) / L.center;
        env = L.peripheral + gn * (L.central - L.peripheral);
      }
      if (L.kind == 0) {
#pragma unroll
        for (int c = 0; c < NI_CH; ++c) acc[c] = (float)((double)acc[c] + env);
        continue;
      }
      float up[NI_CH];
      if (L.lh == H && L.lw == W) {
        // scale 1: source index = d, t = 0 -> weights (0,1,0,0): identity resample
        const float* src = L.lowres + ((size_t)y * W + x) * NI_CH;
        up[0] = __ldg(src); up[1] = __ldg(src + 1); up[2] = __ldg(src + 2);
      } else {
        int iy, ix;
        float wy[4], wx[4];
        cv_coord(y, 1.0 / ((double)H / (double)L.lh), iy, wy);
        cv_coord(x, 1.0 / ((double)W / (double)L.lw), ix, wx);
        int ry[4], rx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ry[k] = min(max(iy - 1 + k, 0), L.lh - 1);
          rx[k] = min(max(ix - 1 + k, 0), L.lw - 1);
        }
#pragma unroll
        for (int c = 0; c < NI_CH; ++c) {
          float rowv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float* row = L.lowres + (size_t)ry[i] * L.lw * NI_CH + c;
            rowv[i] = ((__ldg(row + rx[0] * NI_CH) * wx[0] + __ldg(row + rx[1] * NI_CH) * wx[1]) +
                       __ldg(row + rx[2] * NI_CH) * wx[2]) + __ldg(row + rx[3] * NI_CH) * wx[3];
          }
          up[c] = ((rowv[0] * wy[0] + rowv[1] * wy[1]) + rowv[2] * wy[2]) + rowv[3] * wy[3];
        }
      }
      if (L.kind == 1) {
#pragma unroll
        for (int c = 0; c < NI_CH; ++c) acc[c] = (float)((double)acc[c] + (double)up[c] * env);
      } else {
#pragma unroll
        for (int c = 0; c < NI_CH; ++c) acc[c] = acc[c] + up[c];
      }
    }
    float* dst = out + ((size_t)y * W + x) * NI_CH;
    if (P.mode == AST_INIT_RANDOM) {
#pragma unroll
      for (int c = 0; c < NI_CH; ++c) dst[c] = acc[c] * 0.5f;
    } else if (!P.use_gradient_map) {
      const float r = (float)P.noise_factor;
#pragma unroll
      for (int c = 0; c < NI_CH; ++c) dst[c] = (1.0f - r) * cs[ly + 3][(lx + 3) * NI_CH + c] + r * acc[c];
    } else {
#pragma unroll
      for (int c = 0; c < NI_CH; ++c) {
        // GaussianBlur: row pass then column pass, taps (w1, w0, w1); further taps < 2e-22
        double b = 0.0;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const double* row = &sm[ly + dy][0];
          const double h = P.blur_w1 * row[(lx + 0) * NI_CH + c] + P.blur_w0 * row[(lx + 1) * NI_CH + c] +
                           P.blur_w1 * row[(lx + 2) * NI_CH + c];
          b += (dy == 1 ? P.blur_w0 : P.blur_w1) * h;
        }
        const double r = 5.0 * P.noise_factor / (5.0 + b);
        const double v = (1.0 - r) * (double)cs[ly + 3][(lx + 3) * NI_CH + c] + r * (double)acc[c];
        dst[c] = (float)v;
      }
    }
  }
}

}  // namespace ast

using namespace ast;

extern "C" int ast_noise_init(const float* content_hwc, int H, int W, const ast_noise_level* levels, int n_levels,
                              double noise_factor, int mode, int use_gradient_map, double blur_w0, double blur_w1,
                              float* out_hwc, void* stream) {
  AST_REQUIRE(out_hwc, AST_ERR_INVALID, "ast_noise_init: null output");
  AST_REQUIRE(H > 0 && W > 0, AST_ERR_INVALID, "ast_noise_init: bad shape %dx%d", H, W);
  AST_REQUIRE(mode == AST_INIT_RANDOM || mode == AST_INIT_CONTENT_NOISE, AST_ERR_INVALID, "ast_noise_init: bad mode %d", mode);
  AST_REQUIRE(mode == AST_INIT_RANDOM || content_hwc, AST_ERR_INVALID, "ast_noise_init: content image required");
  AST_REQUIRE(n_levels >= 0 && n_levels <= AST_NOISE_MAX_LEVELS, AST_ERR_UNSUPPORTED,
              "ast_noise_init: %d noise levels (max %d)", n_levels, AST_NOISE_MAX_LEVELS);
  AST_REQUIRE(n_levels == 0 || levels, AST_ERR_INVALID, "ast_noise_init: null levels");
  AST_REQUIRE(H >= 4 && W >= 4, AST_ERR_UNSUPPORTED, "ast_noise_init: image smaller than the Sobel halo");
  NoiseParams P;
  for (int l = 0; l < n_levels; ++l) {
    P.lv[l] = levels[l];
    AST_REQUIRE(levels[l].kind >= 0 && levels[l].kind <= 2, AST_ERR_INVALID, "ast_noise_init: level %d bad kind", l);
    AST_REQUIRE(levels[l].kind == 2 || (levels[l].gy && levels[l].gx && levels[l].center != 0.0), AST_ERR_INVALID,
                "ast_noise_init: level %d missing envelope", l);
    AST_REQUIRE(levels[l].kind == 0 || (levels[l].lowres && levels[l].lh > 0 && levels[l].lw > 0), AST_ERR_INVALID,
                "ast_noise_init: level %d missing low-res grid", l);
  }
  P.n_levels = n_levels;
  P.H = H;
  P.W = W;
  P.mode = mode;
  P.use_gradient_map = use_gradient_map;
  P.noise_factor = noise_factor;
  P.blur_w0 = blur_w0;
  P.blur_w1 = blur_w1;
  dim3 grid((W + NI_TW - 1) / NI_TW, (H + NI_TH - 1) / NI_TH);
  noise_init_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(content_hwc, out_hwc, P);
  return check_launch("ast_noise_init");
}
