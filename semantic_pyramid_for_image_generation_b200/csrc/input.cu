// Device side of the input pipeline: per-sample mask descriptors -> the seven mask tensors, and the
// loader's image normalisation.  Integer / boolean work: results are bit-exact with the host code they replace.
//
// Replaces: misc.get_masks_for_training's tensor construction (misc.py:47-67: ones / zeros / nearest-resized shape image,
// the random draws stay on the host) and Places365.__getitem__'s to_tensor + kornia.normalize_min_max (data.py:49-53).
#include "common.cuh"
#include "../../include/spyramid_b200.h"

extern void spyr_count_launch();

namespace {

// out[b, h, w] for one pyramid level.  depth counts levels from the deepest (0 = logits); a sample keeps level `stage`
// entirely, and -- when it carries a bitmap -- every shallower level shows the bitmap, nearest-neighbour resized with
// src = dst * src_size / dst_size (integer), which is what F.interpolate(mode='nearest') computes for these sizes.
__global__ void expand_mask_level_kernel(const int* __restrict__ stage, const int* __restrict__ bitmap_hw,
                                         const unsigned char* __restrict__ bitmaps, long long bitmap_stride, int depth, int H,
                                         int W, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int st = stage[b], bhw = bitmap_hw[b];
  const long long n = (long long)H * W;
  float* o = out + (long long)b * n;
  const int mode = depth == st ? 1 : ((bhw > 0 && depth > st) ? 2 : 0);
  const unsigned char* bm = bitmaps + (long long)b * bitmap_stride;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = mode == 1 ? 1.f : 0.f;
    if (mode == 2) {
      const int h = (int)(i / W), w = (int)(i % W);
      int sh = (int)(((long long)h * bhw) / H), sw = (int)(((long long)w * bhw) / W);
      if (sh > bhw - 1) sh = bhw - 1;
      if (sw > bhw - 1) sw = bhw - 1;
      v = bm[sh * bhw + sw] ? 1.f : 0.f;
    }
    o[i] = v;
  }
}

// one block per (sample, channel) plane: min / max of the bytes, then out = (1 - -1) * (x/255 - min) / (max - min + 1e-6) + -1
// evaluated with the rounding of the reference's FP32 expression (no FMA contraction)
__global__ void image_minmax_kernel(const unsigned char* __restrict__ img, long long plane, float* __restrict__ out) {
  __shared__ int smin[32], smax[32];
  const unsigned char* src = img + (long long)blockIdx.x * plane;
  float* dst = out + (long long)blockIdx.x * plane;
  int lo = 255, hi = 0;
  const long long n16 = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? plane / 16 : 0;
  for (long long i = threadIdx.x; i < n16; i += blockDim.x) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
    const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int s = 0; s < 32; s += 8) {
        const int v = (int)((wv[k] >> s) & 0xFFu);
        lo = min(lo, v);
        hi = max(hi, v);
      }
  }
  for (long long i = n16 * 16 + threadIdx.x; i < plane; i += blockDim.x) {
    const int v = src[i];
    lo = min(lo, v);
    hi = max(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    smin[threadIdx.x >> 5] = lo;
    smax[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  lo = 255;
  hi = 0;
  for (int i = 0; i < nw; ++i) {
    lo = min(lo, smin[i]);
    hi = max(hi, smax[i]);
  }
  const float fmin_ = __fdiv_rn((float)lo, 255.f), fmax_ = __fdiv_rn((float)hi, 255.f);
  const float denom = __fadd_rn(__fsub_rn(fmax_, fmin_), 1e-6f);
  for (long long i = threadIdx.x; i < plane; i += blockDim.x) {
    const float x = __fdiv_rn((float)src[i], 255.f);
    dst[i] = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, __fsub_rn(x, fmin_)), denom), -1.f);
  }
}

}  // namespace

extern "C" int spyr_expand_mask_level(const int* stage, const int* bitmap_hw, const unsigned char* bitmaps,
                                      long long bitmap_stride, int B, int depth, int H, int W, float* out, void* stream) {
  SPYR_REQUIRE(stage && bitmap_hw && out && B > 0 && H > 0 && W > 0 && depth >= 0, "expand_mask_level: bad arguments");
  SPYR_REQUIRE(bitmaps != nullptr || bitmap_stride == 0, "expand_mask_level: bitmap_stride without bitmaps");
  const long long n = (long long)H * W;
  int gx = (int)((n + 255) / 256);
  if (gx > 64) gx = 64;
  dim3 grid(gx, B);
  expand_mask_level_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(stage, bitmap_hw, bitmaps, bitmap_stride, depth, H, W, out);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}

extern "C" int spyr_image_u8_minmax_normalize(const unsigned char* img, int planes, long long plane, float* out, void* stream) {
  SPYR_REQUIRE(img && out && planes > 0 && plane > 0, "image_u8_minmax_normalize: bad arguments");
  image_minmax_kernel<<<planes, 512, 0, (cudaStream_t)stream>>>(img, plane, out);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
