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
/* number of kernels launched through this library by the calling process (bench.py "gpu_launches") */
long long spyr_launch_count(void);
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
 * With y_f32 != NULL the raw accumulator is instead added (red.global.add.f32) to y_f32[pixel][co]
 * (split-K over `splits` CTAs along the reduction; caller zero-fills y_f32 and applies bias/act itself).
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
} spyr_conv_src;

typedef struct {
  int B, H, W, Cout;
  int nsrc;
  spyr_conv_src src[3];
  const float* bias;          /* [Cout] or NULL */
  const float* stencil_mask;  /* f32 [B,H,W] or NULL: extra 1-channel 3x3 conv input (models.py:94 `cat(.., mask)`) */
  const float* stencil_w;     /* f32 [10][Cout]: 9 taps + row 9 = sum over taps */
  const void* dmask;          /* NHWC bf16 [B,H,W,Cout] or NULL */
  float dmask_slope;
  const void* residual;       /* NHWC bf16 [B,H,W,Cout] or NULL */
  void* y_raw;                /* NHWC bf16 or NULL */
  void* y_act;                /* NHWC bf16 or NULL */
  int act;                    /* 0 none, 1 relu, 2 leaky-relu(act_slope) */
  float act_slope;
  float* y_f32;               /* f32 [B*H*W][Cout] accumulate target or NULL */
  int f32_store;              /* 1: plain store of the accumulator to y_f32 (splits must be 1; no zero-fill needed) */
  int splits;                 /* >=1; >1 requires y_f32 */
  int block_n;                /* 0 = auto, else 32..256 multiple of 16 */
  int stages;                 /* 0 = auto */
} spyr_conv_desc;

int spyr_conv2d_fprop(const spyr_conv_desc* d, void* stream);

/* Weight gradient of the same convolutions (torch autograd conv backward-weight at the call sites above):
 *   dw[tap][ci][co] += sum_{b,h,w} x[b,h+dy,w+dx,ci] * dy[b,h,w,co]     (fp32 red.add; caller zero-fills)
 * Output layout is [taps][Cin][Cout] FP32 ("dgrad-pack order"). */
typedef struct {
  int B, H, W, Cin, Cout, ksize;
  const void* x;   /* NHWC bf16 [B,H,W,Cin]  */
  const void* dy;  /* NHWC bf16 [B,H,W,Cout] */
  float* dw;       /* f32 [taps][Cin][Cout]  */
  int splits;      /* 0 = auto */
  int stages;      /* 0 = auto */
  int per_image;   /* 1: one dw slice per image, dw is f32 [B][taps][Cin][Cout] (attention dK/dV, models.py:266-268) */
  /* debug/validation knobs for the UMMA MN-major descriptors; 0 = defaults */
  int dbg_lbo, dbg_sbo;
} spyr_wgrad_desc;

int spyr_conv2d_wgrad(const spyr_wgrad_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPYRAMID_B200_H */
