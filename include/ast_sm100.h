/*
 * ast_sm100.h — C ABI of libast_sm100.so: the pyramid Gatys-loss hot path of
 * irenemizus/ArtStyleTransfer as hand-written sm_100a (B200) CUDA.
 *
 * The reference has no FFI of its own (it is pure Python on torch / OpenCV); the boundary it
 * offers is its Python call surface.  Each entry point below replaces the library call the
 * reference makes at the cited site (file:line into the reference repository).  The Python
 * binding a maintainer would add is the ctypes stub in INTEGRATION.md (shipped as
 * artstyletransfer_b200/_lib.py).
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, sizes, a cudaStream_t passed as void*; no torch types;
 *   - every call only enqueues work on `stream`; it never synchronises and never allocates;
 *   - return 0 on success, a negative AST_ERR_* otherwise; ast_last_error() (thread-local)
 *     holds the message;
 *   - no hidden mutable state: scratch memory is caller-provided (`ws`).  A workspace must be
 *     zero-filled once when allocated; kernels leave it reusable.  Two host threads may call
 *     concurrently as long as they pass different workspaces (the reference runs closures of
 *     up to 2 jobs from different threads — task_executor.py:9, neural_style_transfer.py:206);
 *   - all tensors are contiguous fp32.  Feature maps are (C, HW) row-major, i.e. torch NCHW
 *     with batch 1; images are (C, H, W) unless a layout argument says otherwise.
 */
#ifndef AST_SM100_H_
#define AST_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AST_ABI_VERSION 9

#define AST_OK               0
#define AST_ERR_INVALID     -1   /* bad argument (null pointer, non-positive size, misalignment) */
#define AST_ERR_CUDA        -2   /* a CUDA runtime / driver call failed                          */
#define AST_ERR_UNSUPPORTED -3   /* shape or mode not implemented by this build                  */
#define AST_ERR_WORKSPACE   -4   /* workspace too small                                          */

/* Operand precision of the Gram contractions (accumulation is always fp32). */
#define AST_PREC_TF32 0   /* tcgen05 kind::tf32, operands rounded to nearest TF32 in shared memory */
#define AST_PREC_FP32 1   /* exact fp32 FFMA path (what torch.bmm does with allow_tf32 = False)    */
#define AST_PREC_BF16 2   /* TF32 everywhere except the C = 512 backward: tcgen05 kind::f16, bfloat16 operands */

#define AST_LAYOUT_CHW 0
#define AST_LAYOUT_HWC 1

/* Source-coordinate arithmetic of the bicubic resampler (both: Keys A=-0.75, per-tap clamp). */
#define AST_COORD_TORCH 0 /* F.interpolate(mode='bicubic'): float scale, float source index  */
#define AST_COORD_CV2   1 /* cv2.resize(INTER_CUBIC): double scale, index cast to float      */

int         ast_version(void);
const char* ast_last_error(void);
/* 0 when the current CUDA device is compute capability 10.x (B200), AST_ERR_UNSUPPORTED otherwise. */
int         ast_device_check(void);

/* ---- Gram matrix + fused MSE --------------------------------------------------------------
 * Replaces math_utils.gram_matrix (math_utils.py:26-34: view, transpose, bmm, /= ch*h*w) and
 * the per-layer torch.nn.MSELoss against the target Gram (neural_style_transfer.py:100-104).
 *
 *   out[C,C] = scale * F F^T - (A ? A : 0);    *loss = mean(out^2) over C*C   (if loss != NULL)
 *
 * Single GPU: scale = 1/(C*HW), A = target Gram -> out is D = G - A, loss the layer's MSE.
 * Plain Gram: A = NULL, loss = NULL.  Row-band sharding: scale = 1, A = NULL gives the raw
 * partial sum to all-reduce, then ast_gram_finalize applies scale / A / MSE on every rank.
 * ws must hold ast_gram_workspace_bytes(C, HW) bytes.
 * ld = row pitch of F in elements (ld >= HW; pass HW for a dense (C, HW) map).  A row band
 * [a, b) of an NCHW map (C, h, w) is F + a*w with HW = (b-a)*w and ld = h*w: no copy needed.
 */
size_t ast_gram_workspace_bytes(int C, int64_t HW);
/* 1 when the tcgen05/TMA path (AST_PREC_TF32) can address this operand: C in {64,128,256,512}, HW and ld
 * multiples of 4 elements (TMA strides are multiples of 16 bytes), F 16-byte aligned; else 0 (use AST_PREC_FP32). */
int ast_gram_tf32_supported(const float* F, int C, int64_t HW, int64_t ld);
int ast_gram_mse_fwd(const float* F, int C, int64_t HW, int64_t ld, float scale, const float* A,
                     float* out, float* loss, void* ws, size_t ws_bytes, int precision,
                     void* stream);
/* round_out = 1: `out` is stored rounded to nearest TF32 (the loss is still computed from the exact values).  Use it
 * when `out` = D only feeds ast_gram_bwd_nhwc(..., d_prerounded = 1): the backward's converter warps then skip D.
 * round_out = 2: `out` is stored as C*C bfloat16 (round to nearest even) for ast_gram_bwd_nhwc_bf16. */
int ast_gram_finalize(const float* G_raw, int C, float scale, const float* A, float* out,
                      float* loss, void* ws, size_t ws_bytes, int round_out, void* stream);

/* The same finalize for up to AST_FINALIZE_MAX_ITEMS raw Grams in ONE launch (row-band sharding: the all-reduced
 * packed buffer of every pyramid level holds five of them; 20 launches per L=3 closure become one).  Every item is
 * out = scale * G_raw - (A ? A : 0), *loss = mean(out^2) (loss may be NULL), out rounded to TF32 when round_out.
 * ws: ast_finalize_batch_workspace_bytes(n_items) bytes, zero-filled once, reusable. */
#define AST_FINALIZE_MAX_ITEMS 32
typedef struct ast_finalize_item {
  const float* G_raw;
  const float* A;        /* nullable */
  float*       out;
  float*       loss;     /* nullable */
  float        scale;
  int32_t      C;
  int32_t      round_out;
  int32_t      pad_;
} ast_finalize_item;
size_t ast_finalize_batch_workspace_bytes(int n_items);
int ast_gram_finalize_batch(const ast_finalize_item* items, int n_items, void* ws, size_t ws_bytes,
                            void* stream);

/* Backward of the style term (autograd of bmm + MSELoss in the reference, 2 bmm per layer):
 *   dF[C,HW] (+)= scale * D D-symmetric [C,C] * F[C,HW]
 * (F and dF share the row pitch ld) with scale = 4 / (C^2 * C*HW) on the host and the upstream gradient read on the device:
 * every *_bwd entry point multiplies its host scale by *gscale when gscale != NULL (a device
 * float, e.g. autograd's grad_output), so no host synchronisation is needed. */
int ast_gram_bwd(const float* D, const float* F, int C, int64_t HW, int64_t ld, float scale,
                 const float* gscale, float* dF, int accumulate, int precision, void* stream);

/* Same two operations for a feature map held as (HW, C) row-major — torch channels_last, the layout in which
 * cuDNN's TF32 convolutions run without NCHW<->NHWC transposes (profiles/r01_vgg_layout.md).  C in
 * {64,128,256,512}, any HW, dense rows (a row band of an NHWC map is a contiguous range of positions: offset the
 * pointer).  TF32 operands only.  out/loss/ws as for ast_gram_mse_fwd; dF may alias nothing else.
 * accumulate != 0 adds into dF with TMA reduce-add (the SM never reads dF). */
int ast_gram_mse_fwd_nhwc(const float* F, int C, int64_t HW, float scale, const float* A, float* out,
                          float* loss, void* ws, size_t ws_bytes, int round_out, void* stream);
/* relu_mask != 0 (both here and in ast_mse_bwd): F / X is the output of a ReLU and the gradient written is the one
 * w.r.t. that ReLU's INPUT: dF = (accumulate ? dF + v : v) * (F > 0) — the tap's loss gradient, the running
 * activation gradient and torch's threshold_backward in one pass over the tensor. */
int ast_gram_bwd_nhwc(const float* D, const float* F, int C, int64_t HW, float scale,
                      const float* gscale, float* dF, int accumulate, int d_prerounded,
                      int relu_mask, void* stream);

/* BF16 operands (AST_PREC_BF16 — north_star: "TF32 or BF16 operands with FP32 accumulation") for the backward at the
 * one tensor-bound width, C = 512: D_bf16 is (C, C) bfloat16 as written by ast_gram_mse_fwd_nhwc / ast_gram_finalize /
 * ast_gram_finalize_batch with round_out = 2 (`out` then holds C*C bfloat16 instead of C*C float); the kernel rounds the
 * staged fp32 feature tile to bfloat16 (cvt.rn) in shared memory.  Narrower layers are HBM-bound: use
 * ast_gram_bwd_nhwc.  The forward contraction (the loss and D itself) always takes TF32 operands. */
int ast_gram_bwd_nhwc_bf16(const void* D_bf16, const float* F, int C, int64_t HW, float scale,
                           const float* gscale, float* dF, int accumulate, int relu_mask, void* stream);

/* ---- Content MSE (neural_style_transfer.py:95) ---------------------------------------------
 *   *loss = scale * sum((X - T)^2)      (scale = 1/n for MSELoss(reduction='mean'))
 *   dX (+)= scale * (X - T)             (scale = 2 * content_weight / n * upstream)       */
size_t ast_reduce_workspace_bytes(void);
int ast_mse_fwd(const float* X, const float* T, int64_t n, float scale, float* loss, void* ws,
                size_t ws_bytes, void* stream);
int ast_mse_bwd(const float* X, const float* T, int64_t n, float scale, const float* gscale,
                float* dX, int accumulate, int relu_mask, void* stream);

/* ---- Glue between the cuDNN convolutions of the VGG19 feature path (neural_nets.py:53-68) --------------
 * Activations are (H, W, C) row-major (torch channels_last), C % 4 == 0, 16-byte aligned.  The convolutions
 * themselves stay on cuDNN; these replace the bias add, ReLU(inplace), max_pool2d (+ its int64 indices) and
 * their autograd backward kernels, which are 40 % of the reference's closure on a B200.
 *   ast_bias_relu_nhwc      : y[p,c] = max(y[p,c] + bias[c], 0) in place
 *   ast_relu_bwd            : g[i] = r[i] > 0 ? g[i] : 0 in place (r = ReLU output)
 *   ast_maxpool2x2_nhwc     : 2x2 / stride 2 max, floor mode (y is (H/2, W/2, C))
 *   ast_maxpool2x2_bwd_nhwc : gx = gy routed to the window's first maximum (torch's tie rule); relu_mask != 0 also
 *                             applies the backward of the ReLU that produced x (x > 0), fusing two torch kernels
 *   ast_chw_to_hwc / ast_hwc_to_chw : (C, HW) planar <-> (HW, C) interleaved for C <= 16 (the image, its gradient);
 *                             plane_stride = elements between channel planes of the planar side (HW when dense,
 *                             H*W of the whole image when converting a band of rows) */
int ast_bias_relu_nhwc(float* y, const float* bias, int C, int64_t n_pos, void* stream);
int ast_relu_bwd(float* g, const float* r, int64_t n, void* stream);
int ast_maxpool2x2_nhwc(const float* x, int C, int H, int W, float* y, void* stream);
int ast_maxpool2x2_bwd_nhwc(const float* gy, const float* x, int C, int H, int W, int relu_mask,
                            float* gx, void* stream);
int ast_chw_to_hwc(const float* x, int C, int64_t HW, int64_t plane_stride, float* y, void* stream);
int ast_hwc_to_chw(const float* x, int C, int64_t HW, float* y, int64_t plane_stride, int accumulate,
                   void* stream);
/* unprepare_img (neural_style_transfer.py:388-393), the device half of the per-step image yield (:207-208):
 * planar (3, HW) image in "x*255 - mean" units -> interleaved (HW, 3) in [0,1], y = fl32(fl32(x + mean_c) / 255)
 * with the mean added in double exactly as numpy's in-place float32 += float64 does.  HW % 4 == 0.  The result is
 * a snapshot: the optimizer may overwrite x as soon as this kernel has run, while the snapshot travels to the host. */
int ast_unprepare_hwc(const float* x_chw, int64_t HW, double mean0, double mean1, double mean2,
                      float* y_hwc, void* stream);

/* ---- Halo rows of a row-band sharded level through NVLink peer memory (EXPERIMENTAL, off by default) ------
 * The reference has no multi-GPU path (a commented-out device round-robin, neural_style_transfer.py:238-243);
 * this replaces the grouped NCCL send/recv that parallel.halo_exchange issues before every 3x3 convolution of a
 * band (and for the gradient rows in the backward) by ONE launch: every entry pushes `src` (my edge row) into the
 * neighbour's staging slot with stores over NVLink, publishes an arrival counter (release, system scope), waits for
 * the neighbour's row to land in MY staging slot and copies it into `halo`.  src == NULL: zero `halo` (border).
 *   dst_remote / flag_remote : the NEIGHBOUR's staging slot pair and arrival counter, peer-mapped into this process
 *                              (torch symmetric memory: _SymmetricMemory.buffer_ptrs);
 *   stage / flag_local       : my staging slot pair and arrival counter (the neighbour writes them);
 *   slot_stride              : bytes between the two slots of a pair (double buffering by exchange parity);
 *   state                    : 3 zero-initialised uint32 in LOCAL memory {exchanges done, ticket, ticket};
 *   both sides must run the same sequence of exchanges; counters are never reset by the kernel. */
#define AST_HALO_MAX_ROWS 16
typedef struct ast_halo_row {
  const void* src;
  void*       dst_remote;
  const void* stage;
  void*       halo;
  uint32_t*   flag_remote;
  uint32_t*   flag_local;
  uint32_t*   state;
  int64_t     bytes;
  int64_t     slot_stride;
} ast_halo_row;
int ast_halo_exchange(const ast_halo_row* rows, int n_rows, void* stream);

/* ---- Image-gradient all-gather of a row-band sharded closure through NVLink peer memory ---------------------
 * (no reference counterpart).  The rows of the gradient pyramid owned by different ranks are disjoint, so instead of
 * all-reducing the whole image gradient every rank stores ITS rows (segments of a symmetric buffer: same layout on every
 * rank) into every peer's buffer and waits for theirs — one launch, afterwards all ranks hold identical gradients.
 *   local_base / peer_base[p] : my symmetric buffer and peer p's, peer-mapped; a segment is [seg_off, seg_off + seg_bytes)
 *                               of both (at most AST_GATHER_MAX_SEGS = levels x planes row ranges);
 *   ready_* / arrive_*        : per-peer uint32 sequence counters in symmetric memory (remote = peer p's counter for me,
 *                               local = my counter for peer p), zero-initialised;
 *   state                     : 32 zero-initialised uint32 of LOCAL device memory.
 * ast_band_announce (start of a closure, stream-ordered after the optimizer update that read the previous gradient)
 * tells every peer that my buffer may be overwritten; ast_band_gather waits for the peers' announcements before it
 * stores, so one buffer is enough (the gradient keeps its address across CUDA-graph replays).  Both sides must run
 * the same sequence of announce / gather calls. */
#define AST_GATHER_MAX_PEERS 7
#define AST_GATHER_MAX_SEGS  24
typedef struct ast_band_gather_desc {
  const char* local_base;
  char*       peer_base[AST_GATHER_MAX_PEERS];
  uint32_t*   ready_remote[AST_GATHER_MAX_PEERS];
  const uint32_t* ready_local[AST_GATHER_MAX_PEERS];
  uint32_t*   arrive_remote[AST_GATHER_MAX_PEERS];
  const uint32_t* arrive_local[AST_GATHER_MAX_PEERS];
  uint32_t*   state;
  int64_t     seg_off[AST_GATHER_MAX_SEGS];
  int64_t     seg_bytes[AST_GATHER_MAX_SEGS];
  int32_t     n_peers;
  int32_t     n_segs;
} ast_band_gather_desc;
int ast_band_announce(const ast_band_gather_desc* desc, void* stream);
int ast_band_gather(const ast_band_gather_desc* desc, void* stream);

/* ---- Total variation (math_utils.py:37-41) -------------------------------------------------
 *   sums[0] = sum |y[..., :-1] - y[..., 1:]|,  sums[1] = sum |y[:, :-1, :] - y[:, 1:, :]|
 *   *tv = (sums[0]/nx)^2 + (sums[1]/ny)^2 with nx = C*H*(W-1), ny = C*(H-1)*W  (tv may be NULL)
 *   dY (+)= kx*sums[0]*d|dx|/dy + ky*sums[1]*d|dy|/dy ;  kx = 2*w*upstream/nx^2, ky likewise */
int ast_tv_fwd(const float* Y, int C, int H, int W, float* sums2, float* tv, void* ws,
               size_t ws_bytes, void* stream);
int ast_tv_bwd(const float* Y, int C, int H, int W, const float* sums2, float kx, float ky,
               const float* gscale, float* dY, int accumulate, void* stream);
/* The same for rows [r0, r1) of every plane only (Y and dY are still the whole (C, H, W) tensors; the rows above r0
 * and below r1 - 1 are read as neighbours): a rank of a row-band sharded job adds the TV gradient of the rows it owns
 * before the image-gradient gather instead of every rank adding all of it afterwards. */
int ast_tv_bwd_rows(const float* Y, int C, int H, int W, int r0, int r1, const float* sums2, float kx,
                    float ky, const float* gscale, float* dY, int accumulate, void* stream);

/* ---- Level loss assembly (neural_style_transfer.py:100-110) ---------------------------------
 *   out4 = { total, content, style, tv } with style = mean(style_mse[0..n_style)) and
 *   total = content_weight*content + style_weight*style + tv_weight*tv.   One tiny launch
 *   instead of ~10 scalar torch kernels per level. */
int ast_level_combine(const float* style_mse, int n_style, const float* content, const float* tv,
                      float content_weight, float style_weight, float tv_weight, float* out4,
                      void* stream);

/* ---- Bicubic pyramid of the optimizing image (neural_style_transfer.py:168-176) -------------
 * y = F.interpolate(x, size=(H/2, W/2), mode='bicubic') for even H, W: fixed taps
 * [-3/32, 19/32, 19/32, -3/32] at rows/cols 2d-1..2d+2 with clamped borders; *_adj is its exact
 * transpose in gather form (deterministic, replaces upsample_bicubic2d_backward's atomics). */
int ast_bicubic_down2x(const float* x, int C, int H, int W, float* y, void* stream);
int ast_bicubic_down2x_adj(const float* gy, int C, int H, int W, float* gx, int accumulate,
                           void* stream);
/* The pyramid step AND total_variation(x) (math_utils.py:37-41, evaluated on every level at
 * neural_style_transfer.py:107) in one pass over x: every pixel and its right / lower neighbour are in the tile the
 * resampler stages anyway.  sums2 / tv as ast_tv_fwd writes them (feed sums2 to ast_tv_bwd).  ws:
 * ast_bicubic_down2x_tv_workspace_bytes(C, H, W) bytes, zero-filled once, reusable. */
size_t ast_bicubic_down2x_tv_workspace_bytes(int C, int H, int W);
int ast_bicubic_down2x_tv(const float* x, int C, int H, int W, float* y, float* sums2, float* tv,
                          void* ws, size_t ws_bytes, void* stream);
/* General ratio (odd pyramid sizes; cv2.resize(..., INTER_CUBIC) at
 * neural_style_transfer.py:226, :304, :427).  The adjoint is CHW only. */
int ast_bicubic_resize(const float* x, int C, int Hin, int Win, float* y, int Hout, int Wout,
                       int layout, int coord_mode, void* stream);
int ast_bicubic_resize_adj(const float* gy, int C, int Hin, int Win, int Hout, int Wout,
                           float* gx, int accumulate, int coord_mode, void* stream);

/* ---- Structured-noise init (neural_style_transfer.py:265-362, :396-418) ---------------------
 * One fused pass over the top-level image:
 *   acc   = sum over levels of   kind 0: envelope | kind 1: up(lowres)*envelope | kind 2: up(lowres)
 *           (float32 accumulation per level, as the reference's in-place += on a float32 array)
 *   envelope(y,x) = peripheral + gy[y]*gx[x]/center*(central - peripheral)           (fp64)
 *   mode AST_INIT_RANDOM        : out = acc * 0.5
 *   mode AST_INIT_CONTENT_NOISE : r = 5*noise_factor/(5 + blur(clip(|Sobel5|, 0, 100)));
 *                                 out = (1-r)*content + r*acc   (fp64, cast to fp32)
 *                                 (use_gradient_map = 0 -> r = noise_factor everywhere)
 * Random numbers stay on the host (numpy legacy RNG, for parity); lowres grids are device
 * pointers to (lh, lw, 3) float32 HWC.  gy/gx: device fp64 vectors of length H / W
 * (cv2.getGaussianKernel).  content/out: (H, W, 3) float32 HWC. */
#define AST_INIT_RANDOM        0
#define AST_INIT_CONTENT_NOISE 1
#define AST_NOISE_MAX_LEVELS   16
typedef struct ast_noise_level {
  const float*  lowres;      /* device, (lh, lw, 3) HWC; NULL for kind 0                      */
  int32_t       lh, lw;
  int32_t       kind;        /* 0 envelope only, 1 noise*envelope, 2 noise only               */
  int32_t       pad_;
  const double* gy;          /* device, length H                                              */
  const double* gx;          /* device, length W                                              */
  double        center;      /* gy[H/2] * gx[W/2] (the reference divides by this element)     */
  double        central, peripheral;
} ast_noise_level;
/* blur_w0 / blur_w1: centre and +-1 taps of cv2.getGaussianKernel(101, 0.2) (the remaining taps
 * are < 2e-22 and dropped). */
int ast_noise_init(const float* content_hwc, int H, int W, const ast_noise_level* levels,
                   int n_levels, double noise_factor, int mode, int use_gradient_map,
                   double blur_w0, double blur_w1, float* out_hwc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AST_SM100_H_ */
