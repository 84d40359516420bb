/*
 * spyramid_b200.h -- C-ABI of the B200 (sm_100a) kernels behind the Semantic-Pyramid GAN training step.
 *
 * The reference (ChristophReich1996/Semantic_Pyramid_for_Image_Generation) has no FFI of its own: its hot path
 * is a chain of stock torch ops inside models.py / lossfunction.py / model_wrapper.py.  Each entry point below
 * replaces one family of those call sites (cited per function as reference file:line).  The Python host
 * (semantic_pyramid_for_image_generation_b200/{models,lossfunction,model_wrapper}.py) binds them with ctypes;
 * INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns int: 0 = ok, non-zero = error; spyr_last_error() gives the message (thread local).
 *   - no function synchronises the device, allocates device memory or calls exit(); all work is enqueued on
 *     `stream` (a cudaStream_t passed as void*), so calls are CUDA-graph capturable.
 *   - feature maps are NHWC (channels innermost) BF16; vectors/statistics/weights-masters are FP32.
 *   - "packed" conv weights are BF16 [taps][Cout][Cin] (fprop) or [taps][Cin][Cout] (dgrad, taps flipped).
 */
#ifndef SPYRAMID_B200_H
#define SPYRAMID_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* spyr_last_error(void);
int spyr_version(void);

/* ------------------------------------------------------------------------------------------------
 * Precision mode (process-wide).
 *   0  BF16 operands, FP32 accumulate (default): every feature map / packed weight is ONE plane of BF16 values.
 *   1  split BF16 ("strict"): every feature map / packed weight is TWO planes, hi = bf16(v) and lo = bf16(v - hi), the lo
 *      plane stored directly behind the hi plane (for a map of n elements at p the lo plane starts at p + n).  The tensor
 *      cores compute x_hi*w_hi + x_lo*w_hi + x_hi*w_lo (three BF16 MMAs, FP32 accumulate), the bandwidth-bound passes
 *      read hi + lo and write both planes: ~16 mantissa bits end to end, which is what the rel-L2 <= 5e-3 network-level
 *      parity against the FP32 reference needs.  Callers allocate twice the elements for every BF16 map.
 *      FP32 tensors (images, vectors, statistics, weight masters, gradients of parameters) are unaffected.
 * Determinism: no entry point uses floating-point atomics in either mode; grid-wide sums go through caller-provided
 * scratch (SPYR_REDUCE_SCRATCH_BYTES) and are added in a fixed order, split-K partial sums are separate slices.
 * ------------------------------------------------------------------------------------------------ */
int spyr_set_precision(int mode);
int spyr_get_precision(void);
/* scratch a deterministic reduction over `n_outputs` values needs (partial vectors of up to 296 blocks, 8-byte slots) */
#define SPYR_REDUCE_BLOCKS 296
#define SPYR_REDUCE_SCRATCH_BYTES(n_outputs) ((long long)SPYR_REDUCE_BLOCKS * (long long)(n_outputs) * 8)
/* partial sums of the batch-norm backward: up to 4 * 296 blocks, each (sum gy, sum gy * xhat) for C channels in FP32 */
#define SPYR_BN_BWD_PARTIAL_BYTES(C) ((long long)(4 * SPYR_REDUCE_BLOCKS + 4 * SPYR_REDUCE_BLOCKS) * 2 * (long long)(C) * 4)
/* number of kernels launched through this library by the calling process (bench.py "gpu_launches") */
long long spyr_launch_count(void);
/* name of the tensor-core kernel the calling thread's last spyr_conv2d_fprop / spyr_conv2d_wgrad launched:
 * conv_stack3_kernel (Cout = 64, 3x3, maps >= 128 wide), conv_halo2_kernel (CTA pairs, Cout % 128 == 0 or 64),
 * conv_halo_kernel (other maps >= 16x8), conv_fprop_kernel (small maps, split-K, FC), wgrad_halo_kernel, conv_wgrad_kernel */
const char* spyr_last_conv_kernel(void);
void spyr_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core convolution (tcgen05.mma + TMEM accumulators + TMA operand loads).
 * Replaces nn.Conv2d 3x3/s1/p1 and 1x1 forward and input-gradient:
 *   models.py:34,55-60,232-243,299-315,393-404,438-449 and torchvision VGG features (models.py:201-202),
 *   and nn.Linear of the VGG classifier (models.py:210-211) as a 1x1 conv over a (B,1,1,C) map.
 * out[b,h,w,co] = sum_src sum_tap sum_ci x_src[b,h+dy,w+dx,ci] * w_src[tap][co][ci]   (zero padding)
 * followed by the fused epilogue (in this order):
 *   v += bias[co]; v += mask-stencil term; v *= (dmask>0 ? 1 : dmask_slope); v += residual;
 *   y_raw = bf16(v); y_act = bf16(act(v))
 * With y_f32 != NULL the raw accumulator goes to FP32 memory instead and the caller applies bias / activation itself:
 *   f32_store = 1: y_f32[pixel][co] = accumulator (one CTA per tile, splits <= 1);
 *   splits > 1:    split-K -- CTA z of the reduction writes its partial sum to slice z, y_f32[z][pixel][co] (plain
 *                  stores, every slice fully written, no zero-fill needed); spyr_conv2d_epilogue / spyr_vec_epilogue add
 *                  the slices in split order, so the result is bit-reproducible (no floating-point atomics).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* x;  /* NHWC bf16 [B,H,W,cin] */
  const void* w;  /* bf16 [ksize*ksize][Cout][cin] */
  int cin;        /* reduction channels of this source, multiple of 8 */
  int ksize;      /* 1 or 3 */
  int w_mn_major; /* 0: w is [taps][Cout][cin] (fprop pack).  1: input-gradient mode -- w is the SAME fprop pack of
                     the forward conv, read as [taps][cin][Cout] with Cout (the forward conv's Cin) contiguous
                     (UMMA MN-major B operand) and the taps flipped: out[p] = sum_t x[p + d_t] * w[8 - t] */
  int w_per_image;/* 1: ksize must be 1; image b uses weight slice w[b] (batched GEMM for SAGAN attention,
                     models.py:266-268) */
  long long w_lo_off; /* split-BF16 mode: elements from w to its lo plane; 0 = slices * Cout * cin (the default layout).
                         Needed when the weight tensor has more rows than Cout (fc8 is stored with 368 rows for 365 classes) */
} spyr_conv_src;

#define SPYR_CONV_MAX_SRC 9
typedef struct {
  int B, H, W, Cout;
  int nsrc;                   /* 1..9 accumulation sources (1..3 in split-BF16 mode, where each source becomes the three
                                 products x_hi*w_hi + x_lo*w_hi + x_hi*w_lo inside the library) */
  spyr_conv_src src[SPYR_CONV_MAX_SRC];
  const float* bias;          /* [Cout] or NULL */
  const float* bias2;         /* further bias vectors added in the epilogue (fused residual branches), or NULL */
  const float* bias3;
  const float* stencil_mask;  /* f32 [B,H,W] or NULL: extra 1-channel 3x3 conv input (models.py:94 `cat(.., mask)`) */
  const float* stencil_w;     /* f32 [10][Cout]: 9 taps + row 9 = sum over taps */
  const void* dmask;          /* NHWC bf16 [B,H,W,Cout] or NULL */
  float dmask_slope;
  const void* residual;       /* NHWC bf16 [B,H,W,Cout] or NULL */
  void* y_raw;                /* NHWC bf16 or NULL */
  void* y_act;                /* NHWC bf16 or NULL */
  int act;                    /* 0 none, 1 relu, 2 leaky-relu(act_slope) */
  float act_slope;
  float* y_f32;               /* f32 [splits][B*H*W][Cout] or NULL */
  int f32_store;              /* 1: plain store of the accumulator to y_f32 (splits must be 1) */
  int splits;                 /* >=1; >1 requires y_f32 (one slice per split) */
  int block_n;                /* 0 = auto, else 32..256 multiple of 16 */
  int stages;                 /* 0 = auto */
  int residual_pooled;        /* 1: `residual` is NHWC [B,H/2,W/2,Cout] and enters as 0.25 * residual[h/2][w/2] -- the
                                 backward of an average-pooled skip branch (models.py:418,465) without materialising
                                 its full-resolution gradient.  Halo-tiled kernels, Cout % 32 == 0 */
  int pool;                   /* 1: 2x2 average pool fused into the epilogue (nn.AvgPool2d after the second conv of a
                                 discriminator block, models.py:406,451): y_raw / y_act / residual are NHWC
                                 [B,H/2,W/2,Cout]; the residual is added after pooling.  Needs Cout % 32 == 0, maps of at
                                 least 16x8 and none of dmask / stencil / y_f32 / splits */
} spyr_conv_desc;

int spyr_conv2d_fprop(const spyr_conv_desc* d, void* stream);
/* Fused tail of a split-K run: applies the epilogue fields of `d` (biases, stencil, gate, residual, y_raw / y_act) to the
 * FP32 partial accumulators `acc` [d->splits][B*H*W][Cout] that spyr_conv2d_fprop(y_f32 = acc, splits > 1) produced.  Used for the 4x4 / 8x8
 * maps, where 3..30 output tiles cannot fill 148 SMs and the reduction dimension is split instead. */
int spyr_conv2d_epilogue(const spyr_conv_desc* d, const float* acc, void* stream);

/* Weight gradient of the same convolutions (torch autograd conv backward-weight at the call sites above):
 *   dw[tap][ci][co] += sum_{b,h,w} x[b,h+dy,w+dx,ci] * dy[b,h,w,co]
 * Output layout is [taps][Cin][Cout] FP32 ("dgrad-pack order").  The pixel range is split over CTAs; with more than one
 * split each CTA stores its partial sum to its own slice of `scratch` and a second kernel adds the slices to dw in slice
 * order (no floating-point atomics: bit-reproducible).  spyr_conv2d_wgrad_scratch_floats(d) gives the scratch size
 * (0 = not needed).  In split-BF16 mode x and dy are hi + lo plane pairs and three products are accumulated. */
typedef struct {
  int B, H, W, Cin, Cout, ksize;
  const void* x;   /* NHWC bf16 [B,H,W,Cin]  */
  const void* dy;  /* NHWC bf16 [B,H,W,Cout] */
  float* dw;       /* f32 [taps][Cin][Cout]  */
  int cin_stride;  /* 0 = Cin; else row stride (in input channels) of dw: element (tap, ci, co) lives at
                      dw[(tap*cin_stride + ci)*Cout + co] -- lets the mask channel of `cat(feature*mask, mask)` share the
                      buffer (models.py:94) */
  int splits;      /* 0 = auto */
  int stages;      /* 0 = auto */
  int per_image;   /* 1: one dw slice per image, dw is f32 [B][taps][Cin][Cout] (attention dK/dV, models.py:266-268) */
  /* debug/validation knobs for the UMMA MN-major descriptors; 0 = defaults */
  int dbg_lbo, dbg_sbo;
  float* scratch;            /* partial-sum slices, >= spyr_conv2d_wgrad_scratch_floats(d) floats (may be NULL if 0) */
  long long scratch_floats;
} spyr_wgrad_desc;

long long spyr_conv2d_wgrad_scratch_floats(const spyr_wgrad_desc* d); /* host only; -1 on a bad descriptor */
int spyr_conv2d_wgrad(const spyr_wgrad_desc* d, void* stream);

/* Deferred second stages.  A backward pass launches ~90 weight gradients and ~55 bias / mask-channel sums, each followed
 * by a small kernel that adds its partial results in a fixed order.  The *_deferred entry points launch only the first
 * stage and describe the pending sum in `out`; spyr_reduce_batched then performs up to SPYR_REDUCE_BATCH pending sums in
 * ONE launch (same arithmetic, same order: bit-identical to the immediate form).  out->kind == 0 means nothing is pending
 * (a single slice was accumulated directly). */
typedef struct {
  int kind;                 /* 0 none, 1 weight-gradient slices (float4 lanes), 2 partial vectors (one warp per output) */
  const float* partial;     /* kind 1: [nslices][nimg][taps*cin_stride*Cout];  kind 2: [nb][n] */
  float* out0;              /* kind 1: dw;  kind 2: first destination */
  float* out1;
  float* out2;
  int nslices, nimg, taps, rows, cin_stride, Cout, lanes; /* kind 1 */
  int nb, n, mode, C, sink_stride, sink_row, split;       /* kind 2: SumSink of csrc/common.cuh */
} spyr_reduce_entry;
#define SPYR_REDUCE_BATCH 32
int spyr_conv2d_wgrad_deferred(const spyr_wgrad_desc* d, spyr_reduce_entry* out, void* stream);
int spyr_colsum_deferred(const void* g, long long rows, int C, float* out0, float* out1, float* out2, void* scratch,
                         spyr_reduce_entry* out, void* stream);
int spyr_stencil_wgrad_deferred(const float* mask, const void* g, int B, int H, int W, int C, float* dw, int cin_stride,
                                int ci_row, void* scratch, spyr_reduce_entry* out, void* stream);
int spyr_reduce_batched(const spyr_reduce_entry* entries, int n, void* stream); /* n <= SPYR_REDUCE_BATCH */

/* ------------------------------------------------------------------------------------------------
 * Image <-> first-layer operand.  The 3-channel 3x3 convolutions (VGG features.0, models.py:201; Discriminator
 * layers.0.main_block.0, models.py:393) run as a 1x1 tensor-core conv over 32-wide im2col rows
 * (k = tap*3 + c, columns 27..31 zero).  `mean3/invstd3` (may be NULL) fold kornia.normalize (models.py:195-197).
 * Images are NCHW FP32 (data.py:76-90).
 * ------------------------------------------------------------------------------------------------ */
int spyr_im2col3x3(const float* img, int B, int H, int W, const float* mean3, const float* invstd3, void* out, void* stream);
int spyr_col2im3x3(const void* gcol, int B, int H, int W, const float* invstd3, float* gimg, int accumulate, void* stream);
/* AvgPool2d(2) of the image for the input block's skip conv (models.py:417): (B,3,H,W) f32 -> (B,H/2,W/2,8) bf16 */
int spyr_img_avgpool_pad8(const float* img, int B, int H, int W, void* out, void* stream);
int spyr_img_avgpool_pad8_bwd(const void* g8, int B, int H, int W, float* gimg, int accumulate, void* stream);
/* API-boundary layout conversion; `mask` (f32 [B,HW], may be NULL) gates the feature (models.py:94) */
int spyr_nchw_to_nhwc(const float* src, const float* mask, float slope, void* dst, int B, int C, int HW, void* stream);
/* gate_x (f32 NCHW, may be NULL): dst *= (gate_x > 0 ? 1 : slope) -- backward of the LeakyReLU at models.py:33 */
int spyr_nhwc_to_nchw(const void* src, const float* gate_x, float slope, float* dst, int B, int C, int HW, void* stream);
int spyr_maskgate(const void* f, const float* mask, void* out, long long npix, int C, void* stream);
/* dst[t][n][k] = src[taps-1-t][k][n] (bf16): forward weight pack [tap][Cout=K][Cin=N] -> K-major operand of the input
 * gradient (taps flipped); lets 64-wide input gradients run on the row-stacked / CTA-pair kernels */
int spyr_weight_transpose_flip(const void* src, void* dst, int taps, int K, int N, void* stream);

/* ---- VGG-16 fine-tuning (vgg_16_train.py:134-165): what the frozen-encoder path did not need ----
 * weight gradient from the tensor-core layout gw[(t*cin_stride + ci)*Cout + co] to the parameter's (Cout,Cin,kh,kw) */
int spyr_wgrad_to_oihw(const float* gw, float* out, int taps, int Cin, int Cout, int cin_stride, int accumulate, void* stream);
/* nn.Dropout of the torchvision classifier: y = keep ? x / (1-p) : 0, keep mask (u8) saved for the backward;
 * counter-based generator keyed by (seed, offset + element index) */
int spyr_dropout_fwd(const float* x, long long n, float p, unsigned long long seed, unsigned long long offset, float* y,
                     void* y_bf16, unsigned char* mask, void* stream);
int spyr_dropout_bwd(const float* g, const unsigned char* mask, long long n, float p, float* out, void* stream);

/* ---- input pipeline: what the reference's DataLoader workers compute per sample, on the device ----
 * One pyramid level of the mask set of B samples from their descriptors (misc.py:47-67).  `depth` counts levels from the
 * deepest (0 = logits (365), 1 = fc7 (4096), 2 = 8x8 ... 6 = 128x128); vector levels use H = 1.  stage[b] is the kept
 * level; bitmap_hw[b] > 0 marks a spatial sample whose square uint8 {0,1} bitmap (row-major, bitmap_hw x bitmap_hw,
 * at bitmaps + b * bitmap_stride) shows through every level shallower than stage[b], nearest-neighbour resized.
 * out: f32 [B, H*W], values exactly 0.0 / 1.0. */
int spyr_expand_mask_level(const int* stage, const int* bitmap_hw, const unsigned char* bitmaps, long long bitmap_stride,
                           int B, int depth, int H, int W, float* out, void* stream);
/* data.py:49-53: uint8 planes (B*C of them, `plane` bytes each) -> x/255 -> per-plane min-max to [-1, 1]
 * (kornia.normalize_min_max, eps 1e-6), FP32 out, same rounding as the reference's expression */
int spyr_image_u8_minmax_normalize(const unsigned char* img, int planes, long long plane, float* out, void* stream);

/* pooling (NHWC bf16; H, W are the HIGH-resolution dims everywhere) */
int spyr_avgpool2_fwd(const void* x, const void* residual, void* y_raw, void* y_act, float slope, int B, int H, int W, int C,
                      void* stream); /* models.py:406,415-418,451,465; residual is added after pooling */
int spyr_avgpool2_bwd(const void* g_lo, void* g_hi, int B, int H, int W, int C, void* stream);
int spyr_maxpool2_fwd(const void* x, void* y, int B, int H, int W, int C, void* stream); /* models.py:203,245 */
int spyr_maxpool2_bwd(const void* x, const void* gy, void* gx, int B, int H, int W, int C, int relu_gate, int accumulate,
                      void* stream);
int spyr_adaptive_avgpool_fwd(const void* x, void* y, int B, int H, int W, int OH, int OW, int C, void* stream); /* :206 */
int spyr_adaptive_avgpool_bwd(const void* gy, const void* residual, void* gx, int B, int H, int W, int OH, int OW, int C,
                              void* stream);
int spyr_global_avgpool_lrelu_fwd(const void* x, float slope, float* out, int B, int P, int C, void* stream); /* :125-127 */
int spyr_global_avgpool_lrelu_bwd(const void* x, const float* gfeat, float slope, void* gx, int B, int P, int C, void* stream);

/* out = gamma*t + x (models.py:274) and backward (gt = gamma*g, dgamma += <g,t>) */
int spyr_gamma_residual_fwd(const void* t, const void* x, const float* gamma, void* out, void* out_act, float slope,
                            long long n, void* stream);
int spyr_gamma_residual_bwd(const void* g, const void* t, const float* gamma, void* gt, float* dgamma, long long n,
                            void* scratch /* SPYR_REDUCE_SCRATCH_BYTES(1) */, void* stream);
/* bias gradients: out_i[c] += sum_rows g[row][c] (out1/out2 may be NULL) */
int spyr_colsum(const void* g, long long rows, int C, float* out0, float* out1, float* out2,
                void* scratch /* SPYR_REDUCE_SCRATCH_BYTES(C) */, void* stream);
/* weight gradient of the mask channel of cat(feature*mask, mask): dw[(t*cin_stride+ci_row)*C + co] += ... */
int spyr_stencil_wgrad(const float* mask, const void* g, int B, int H, int W, int C, float* dw, int cin_stride, int ci_row,
                       void* scratch /* SPYR_REDUCE_SCRATCH_BYTES(9 * C) */, void* stream);
int spyr_cast_f32_bf16(const float* src, void* dst, long long n, void* stream);
/* epilogue of the split-K FC layers (VGG classifier, models.py:210-213) and of their input-gradients:
 * v = sum_{s < nsplit} acc[s][b][n] (the split-K slices of spyr_conv2d_fprop, summed in split order) + bias[n] + add[b][n];
 * mode 1: v = relu(v); mode 2: v *= (gate[b][n] > 0); any pointer but acc may be NULL */
int spyr_vec_epilogue(const float* acc, int nsplit, const float* bias, const float* add, const float* gate, int mode,
                      float* out_f32, void* out_bf16, int ld_bf16 /* row stride of out_bf16, >= N */, int B, int N,
                      void* stream);
/* generator tail: img = tanh(conv1x1(a; W/sigma) + b) -> NCHW f32 (models.py:58-61,99) and its backward, which emits
 * the gradient w.r.t. the PRE-LeakyReLU input of the 1x1 conv (gate from a), dW (w.r.t. W/sigma) and db */
int spyr_conv1x1_tanh_fwd(const void* a, const float* w, const float* sigma, const float* bias, float* img, int B, int HW,
                          int C, int Cout, void* stream);
int spyr_conv1x1_tanh_bwd(const float* gimg, const float* img, const void* a, const float* w, const float* sigma, float slope,
                          void* gh, float* dw, float* db, int B, int HW, int C, int Cout,
                          void* scratch /* SPYR_REDUCE_SCRATCH_BYTES(Cout * C + Cout) */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (Conditional) batch norm, train-mode statistics, fused with LeakyReLU and bilinear x2 (align_corners=True).
 * models.py:491-506 (ConditionalBatchNorm), :51-54 (final block), :295-310 (generator block).
 * Affine convention: scale = scale_ptr[row*row_stride + c], shift = shift_ptr[row*row_stride + c], row = cls[b] or 0.
 * mode 0: a = lrelu(aff(x));  mode 1: a = up2(lrelu(aff(x))), xu = up2(x);  mode 2: a = lrelu(aff(up2(x))).
 * ------------------------------------------------------------------------------------------------ */
/* Statistics are deterministic two-stage sums: spyr_bn_stats writes one partial vector (sum, sum of squares; FP64) per
 * block into `partials` (SPYR_REDUCE_SCRATCH_BYTES(2 * C) bytes) and spyr_bn_finalize adds them in block order. */
int spyr_bn_stats(const void* x, int B, int H, int W, int C, int up2, double* partials, void* stream);
/* final block (models.py:52-54, upsample -> BN): writes xu = up2(x) (bf16 [B,2H,2W,C]) once and its statistics;
 * sums may be NULL (plain bilinear x2, align_corners=True: the skip branch of a generator block, models.py:338) */
int spyr_up2_stats(const void* x, int B, int H, int W, int C, void* xu_out, double* partials /* may be NULL */,
                   void* stream);
int spyr_bn_finalize(const double* partials /* of spyr_bn_stats over `count` pixels; NULL in eval mode */, double count,
                     int C, float eps, float momentum, float* running_mean,
                     float* running_var, long long* num_batches_tracked, float* mean_rstd /* [2C] */, int training,
                     void* stream);
int spyr_bn_act(const void* x, const float* mean_rstd, const float* scale_ptr, const float* shift_ptr, int row_stride,
                const int* cls, float slope, int mode, void* out_a, void* out_xu, int B, int H, int W, int C, void* stream);
int spyr_bn_bwd_reduce(const void* g, const void* x, const float* mean_rstd, const float* scale_ptr, const float* shift_ptr,
                       int row_stride, const int* cls, float slope, int mode, void* gy_out,
                       float* S /* out: [B][2][C] sums, followed by the block partials they are added from (fixed order);
                                   SPYR_BN_BWD_PARTIAL_BYTES(C) bytes in all */,
                       int B, int H, int W, int C, void* stream);
int spyr_bn_bwd_finalize(const float* S, int B, int C, float count, const float* scale_ptr, int row_stride, const int* cls,
                         float* M /* [2C] */, float* d_scale, float* d_shift, void* stream);
/* the parameter-gradient half of spyr_bn_bwd_finalize on its own (a leaf of the backward pass: d_scale[row(b)][c] += S2[b,c],
 * d_shift likewise with S1, samples in order); spyr_bn_bwd_finalize with d_scale == NULL computes the channel means only */
int spyr_bn_bwd_params(const float* S, int B, int C, int row_stride, const int* cls, float* d_scale, float* d_shift,
                       void* stream);
int spyr_bn_bwd_apply(const void* gy, const void* x, const float* mean_rstd, const float* scale_ptr, int row_stride,
                      const int* cls, const float* M, const void* residual, void* gx, int B, int H, int W, int C, int x_up2,
                      void* stream); /* x_up2: gy/gx at 2H x 2W, xhat from up2(x); H, W are x's dims */
int spyr_up2_bwd(const void* g_hi, void* g_lo, int B, int H, int W, int C, void* stream); /* H, W = LOW-res dims */
int spyr_argmax_rows(const void* onehot, int is_int64, int B, int n, int* out, void* stream); /* models.py:151,501 */

/* ------------------------------------------------------------------------------------------------
 * Spectral normalisation of every layer of a model in one batched pass (torch spectral_norm.py:92-114).
 * The table is static per model (pointers to weight_orig / weight_u / weight_v); per-forward outputs go to
 * caller-provided arenas at the offsets recorded in the table.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const float* w;        /* weight_orig viewed as (rows, cols); conv element (co, ci, t) at co*cols + ci*taps + t */
  float* u;              /* weight_u [rows], updated in place in training mode */
  float* v;              /* weight_v [cols] */
  int rows, cols, taps, cin; /* cols = cin * taps */
  int pack_cin;          /* >0: emit bf16 packed[t][rows][pack_cin] at pack_off (split-BF16 mode: followed by its lo plane,
                            so slices must be spaced by twice the elements); 0: FP32 consumers only need sigma */
  int pack_mode;         /* 0: as above.  1: im2col rows packed[rows][pack_cin], k = t*cin + ci, zero padded (3-channel
                            first layers, models.py:393,403) */
  long long pack_off;    /* element offset into the bf16 arena */
  long long stencil_off; /* >=0: cin == pack_cin + 1, emit f32 stencil[10][rows] for channel pack_cin; else -1 */
  long long gw_off;      /* backward: float offset of G = dL/d(W/sigma) in gw_arena, or -1 (no gradient this pass) */
  int gw_layout;         /* 0: same layout as w;  1: [tap][cin][rows] (spyr_conv2d_wgrad order) */
  long long grad_off;    /* backward: float offset of dL/dweight_orig in grad_arena (layout of w) */
  /* filled by spyr_sn_plan: */
  int index, tile0_wtu, tile0_wv, tile0_pack, tile0_bwd, tile0_tsum;
  long long scratch_off, saved_off; /* saved arena per layer: sigma, u[rows], v[cols] (the clones torch keeps) */
  long long part_off;               /* scratch: row-tile partial sums of W^T u (added in tile order: no atomics) */
} spyr_sn_layer;
typedef struct {
  int tiles_wtu, tiles_wv, tiles_pack, tiles_bwd, tiles_tsum;
  long long scratch_floats, saved_floats;
} spyr_sn_plan_out;
int spyr_sn_plan(spyr_sn_layer* host_tab, int n, spyr_sn_plan_out* out); /* host only; fills the planning fields */
int spyr_sn_forward(const spyr_sn_layer* dev_tab, int n, const spyr_sn_plan_out* plan, int training, float eps,
                    float* scratch, void* packed, float* stencil, float* saved, void* stream);
int spyr_sn_backward(const spyr_sn_layer* dev_tab, int n, const spyr_sn_plan_out* plan, const float* gw_arena,
                     const float* saved, float* dots /* [plan->tiles_bwd] per-tile partial dot products */, float* grad_arena,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Small-batch FP32 linear layers (models.py:28-31,356,360,128,132): y = lrelu_out((f(x) W^T)/sigma + b + y_add),
 * f(x) = lrelu_in(x * xmask).  slopes of 1 disable the activations; sigma/bias/xmask/y_add may be NULL.
 * ------------------------------------------------------------------------------------------------ */
int spyr_linear_fwd(const float* x, const float* xmask, float in_slope, const float* w, const float* sigma,
                    const float* bias, const float* y_add, float out_slope, float* y, int B, int K, int O, void* stream);
int spyr_linear_bwd_x(const float* gy, const float* y, float out_slope, const float* w, const float* sigma, const float* x,
                      float in_slope, float* gx, int accumulate, int B, int K, int O, void* stream);
int spyr_linear_bwd_w(const float* gy, const float* y, float out_slope, const float* x, const float* xmask, float in_slope,
                      float* gw, float* gb, int B, int K, int O, void* stream);
/* Discriminator output (B,B,E): out[i][j][k] = cls[j] + feat[j][k]*emb_w[idx[i]][k]/sigma (models.py:151-155) */
int spyr_dhead_out_fwd(const float* cls, const float* feat, const float* emb_w, const float* sigma, const int* idx,
                       float* out, int B, int E, void* stream);
int spyr_dhead_out_bwd(const float* g, const float* feat, const float* emb_w, const float* sigma, const int* idx,
                       float* g_cls, float* g_feat, float* g_embw, int B, int E, void* stream);

/* Fused SAGAN attention forward (models.py:262-270): o[b,q,:] = softmax_k(q[b,q,:] . k[b,k,:]) @ v[b,k,:], no 1/sqrt(d).
 * q [B,HW,d], k [B,nk,d], v [B,nk,dv], o [B,HW,dv] BF16 row-major (the NHWC maps of the 1x1 convolutions);
 * p_out (may be NULL) receives the normalised attention map [B,HW,nk] in BF16 for the backward pass.
 * Limits: HW % 128 == 0, 8 <= d <= 64, nk in {64,128,192,256}, dv in {64,128}. */
int spyr_sagan_attention_fwd(const void* q, const void* k, const void* v, void* o, void* p_out, int B, int HW, int d, int nk,
                             int dv, void* stream);
/* attention-map softmax over keys (models.py:266) and its backward (unfused path / backward pass) */
int spyr_softmax_rows_fwd(const float* s, void* p, long long rows, int n, void* stream);
int spyr_softmax_rows_bwd(const void* p, const float* dp, void* ds, long long rows, int n, void* stream);

/* losses (lossfunction.py) */
int spyr_lsgan_fwd(const float* p, long long n, float target, float* out, void* stream);           /* :137,:164 */
int spyr_lsgan_bwd(const float* p, long long n, float target, const float* gout, float* gp, void* stream);
int spyr_rec_level_fwd(const void* fr, const void* ff, const float* mask, int B, int H, int W, int C, float* loss,
                       void* scratch /* SPYR_REDUCE_SCRATCH_BYTES(1) */, void* stream);             /* :45-49,67 (loss +=) */
int spyr_rec_level_bwd(const void* fr, const void* ff, const float* mask, int B, int H, int W, int C, const float* gout,
                       void* gff, void* stream);
int spyr_rec_vec_fwd(const float* fr, const float* ff, const float* mask, int B, int N, float* loss,
                     void* scratch /* SPYR_REDUCE_SCRATCH_BYTES(1) */, void* stream);               /* :57-59 (loss +=) */
int spyr_rec_vec_bwd(const float* fr, const float* ff, const float* mask, int B, int N, const float* gout, float* gff,
                     void* stream);
int spyr_diversity_fwd(const float* img, long long img_half, const float* z, long long z_half, float* work /* [2] */,
                       float* loss, void* scratch /* SPYR_REDUCE_SCRATCH_BYTES(2) */, void* stream); /* :102-110 */
int spyr_diversity_bwd(const float* img, long long img_half, const float* work, const float* gout, float* gimg, void* stream);

/* fused multi-tensor Adam (torch.optim.Adam defaults, main.py:64-65); the step counter lives on the device */
#define SPYR_ADAM_MAX_TENSORS 48
typedef struct {
  int count;
  float* p[SPYR_ADAM_MAX_TENSORS];
  const float* g[SPYR_ADAM_MAX_TENSORS];
  float* m[SPYR_ADAM_MAX_TENSORS];
  float* v[SPYR_ADAM_MAX_TENSORS];
  long long n[SPYR_ADAM_MAX_TENSORS];
} spyr_adam_chunk;
/* dst += src over a flat FP32 buffer (accumulating a second backward pass into the gradient arena of the first) */
int spyr_add_inplace(float* dst, const float* src, long long n, void* stream);
int spyr_adam_tick(int* step, void* stream);
int spyr_adam_step(const spyr_adam_chunk* chunk, const int* step, float lr, float beta1, float beta2, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPYRAMID_B200_H */
