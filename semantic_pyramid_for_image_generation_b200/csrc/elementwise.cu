// Bandwidth-bound glue of the training step: one coalesced, 16-byte-vectorised pass each.
// Feature maps are NHWC BF16 (8 channels = one uint4 per thread); images are NCHW FP32 as the reference's
// data loader delivers them (data.py:76-90).
//
// Replaces the ATen elementwise / pooling / interpolation calls of the reference:
//   kornia.normalize (models.py:195-197), nn.MaxPool2d (models.py:203,245), nn.AvgPool2d (models.py:406,451),
//   nn.AdaptiveAvgPool2d (models.py:126,206), feature*mask and torch.cat (models.py:78-94), gamma*o+x (models.py:274),
//   the final 1x1 conv + tanh (models.py:58-61,99).
#include "common.cuh"
#include "../../include/spyramid_b200.h"

extern void spyr_count_launch();

namespace {


inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------------------------
// 3-channel image <-> 32-wide im2col rows (k = tap*3 + c, taps row-major over (dy,dx), 27..31 zero)
// ---------------------------------------------------------------------------------------------
template <bool S>
__global__ void im2col3x3_kernel(const float* __restrict__ img, int B, int H, int W, const float* __restrict__ mean,
                                 const float* __restrict__ invstd, const ActT<S> out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long npix = (long long)B * H * W;
  if (idx >= npix) return;
  const int w = (int)(idx % W);
  const int h = (int)((idx / W) % H);
  const int b = (int)(idx / ((long long)W * H));
  float m[3] = {0.f, 0.f, 0.f}, is[3] = {1.f, 1.f, 1.f};
  if (mean != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      m[c] = mean[c];
      is[c] = invstd[c];
    }
  }
  float v[32];
#pragma unroll
  for (int i = 27; i < 32; ++i) v[i] = 0.f;
  const float* base = img + (size_t)b * 3 * H * W;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
    const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[t * 3 + c] = in ? (__ldg(base + ((size_t)c * H + hh) * W + ww) - m[c]) * is[c] : 0.f;
  }
  const ActT<S> dst = out + idx * 32;
#pragma unroll
  for (int g = 0; g < 4; ++g) st8(dst + g * 8, v + g * 8);
}

template <bool S>
__global__ void col2im3x3_kernel(const ActT<S> gcol, int B, int H, int W, const float* __restrict__ invstd,
                                 float* __restrict__ gimg, int accumulate) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long npix = (long long)B * H * W;
  if (idx >= npix) return;
  const int w = (int)(idx % W);
  const int h = (int)((idx / W) % H);
  const int b = (int)(idx / ((long long)W * H));
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    // col[p][t] = img[p + d_t]  =>  gimg[q] += gcol[q - d_t][t]
    const int hh = h - (t / 3 - 1), ww = w - (t % 3 - 1);
    if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
    const ActT<S> src = gcol + ((((size_t)b * H + hh) * W + ww) * 32 + t * 3);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += ldf(src, c);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = acc[c];
    if (invstd != nullptr) v *= invstd[c];
    float* dst = gimg + (((size_t)b * 3 + c) * H + h) * W + w;
    *dst = accumulate ? (*dst + v) : v;
  }
}

// (B,3,H,W) f32 -> 2x2 average -> (B,H/2,W/2,8) bf16 (channels 3..7 zero): input of the D input block's skip conv
template <bool S>
__global__ void img_avgpool_pad8_kernel(const float* __restrict__ img, int B, int H, int W, const ActT<S> out) {
  const int OH = H / 2, OW = W / 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * OH * OW) return;
  const int w = (int)(idx % OW);
  const int h = (int)((idx / OW) % OH);
  const int b = (int)(idx / ((long long)OW * OH));
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* p = img + (((size_t)b * 3 + c) * H + 2 * h) * W + 2 * w;
    v[c] = 0.25f * (p[0] + p[1] + p[W] + p[W + 1]);
  }
  st8(out + idx * 8, v);
}
template <bool S>
__global__ void img_avgpool_pad8_bwd_kernel(const ActT<S> g8, int B, int H, int W, float* __restrict__ gimg,
                                            int accumulate) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int w = (int)(idx % W);
  const int h = (int)((idx / W) % H);
  const int b = (int)(idx / ((long long)W * H));
  const ActT<S> src = g8 + (((size_t)b * (H / 2) + h / 2) * (W / 2) + w / 2) * 8;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = 0.25f * ldf(src, c);
    float* dst = gimg + (((size_t)b * 3 + c) * H + h) * W + w;
    *dst = accumulate ? (*dst + v) : v;
  }
}

// ---------------------------------------------------------------------------------------------
// NCHW f32 <-> NHWC bf16 (API boundary only), optional per-pixel mask gate
// ---------------------------------------------------------------------------------------------
template <bool S>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, const float* __restrict__ mask, float slope,
                                    const ActT<S> dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? src[((size_t)b * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) {
      float v = tile[threadIdx.x][i];
      if (mask != nullptr) v *= mask[(size_t)b * HW + p];
      stf(dst, ((size_t)b * HW + p) * C + c, lrelu_f(v, slope));
    }
  }
}
template <bool S>
__global__ void nhwc_to_nchw_kernel(const ActT<S> src, const float* __restrict__ gate_x, float slope,
                                    float* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? ldf(src, ((size_t)b * HW + p) * C + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) {
      const size_t o = ((size_t)b * C + c) * HW + p;
      float v = tile[threadIdx.x][i];
      if (gate_x != nullptr && !(gate_x[o] > 0.f)) v *= slope;
      dst[o] = v;
    }
  }
}

template <bool S>
__global__ void maskgate_kernel(const ActT<S> f, const float* __restrict__ mask, const ActT<S> out,
                                long long npix, int cg) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix * cg) return;
  const long long pix = idx / cg;
  float v[8];
  ld8(f + idx * 8, v);
  const float m = __ldg(mask + pix);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= m;
  st8(out + idx * 8, v);
}

// ---------------------------------------------------------------------------------------------
// pooling
// ---------------------------------------------------------------------------------------------
template <bool S>
__global__ void avgpool2_fwd_kernel(const ActT<S> x, const ActT<S> residual, const ActT<S> y_raw,
                                    const ActT<S> y_act, float slope, int B, int H, int W, int cg) {
  const int OH = H / 2, OW = W / 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * OH * OW * cg) return;
  const int c = (int)((unsigned)idx % (unsigned)cg);  // 32-bit index split: element counts are < 2^31 (checked at launch)
  int t = (int)((unsigned)idx / (unsigned)cg);
  const int w = (int)(t % OW);
  t /= OW;
  const int h = (int)(t % OH);
  const int b = (int)(t / OH);
  const size_t C = (size_t)cg * 8;
  const ActT<S> p = x + ((((size_t)b * H + 2 * h) * W + 2 * w) * C + (size_t)c * 8);
  float a[8], v[8];
  ld8(p, v);
  ld8(p + C, a);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] += a[j];
  ld8(p + (size_t)W * C, a);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] += a[j];
  ld8(p + (size_t)W * C + C, a);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.25f * (v[j] + a[j]);
  if (!residual.null()) {
    ld8(residual + idx * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += a[j];
  }
  if (!y_raw.null()) st8(y_raw + idx * 8, v);
  if (!y_act.null()) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = lrelu_f(v[j], slope);
    st8(y_act + idx * 8, v);
  }
}
// g_hi[b,h,w,:] = 0.25 * g_lo[b,h/2,w/2,:]   (H, W are the high-resolution dims)
template <bool S>
__global__ void avgpool2_bwd_kernel(const ActT<S> g_lo, const ActT<S> g_hi, int B, int H, int W, int cg) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W * cg) return;
  const int c = (int)((unsigned)idx % (unsigned)cg);  // 32-bit index split: element counts are < 2^31 (checked at launch)
  int t = (int)((unsigned)idx / (unsigned)cg);
  const int w = (int)(t % W);
  t /= W;
  const int h = (int)(t % H);
  const int b = (int)(t / H);
  float v[8];
  ld8(g_lo + ((((size_t)b * (H / 2) + h / 2) * (W / 2) + w / 2) * cg + c) * 8, v);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= 0.25f;
  st8(g_hi + idx * 8, v);
}

template <bool S>
__global__ void maxpool2_fwd_kernel(const ActT<S> x, const ActT<S> y, int B, int H, int W, int cg) {
  const int OH = H / 2, OW = W / 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * OH * OW * cg) return;
  const int c = (int)((unsigned)idx % (unsigned)cg);  // 32-bit index split: element counts are < 2^31 (checked at launch)
  int t = (int)((unsigned)idx / (unsigned)cg);
  const int w = (int)(t % OW);
  t /= OW;
  const int h = (int)(t % OH);
  const int b = (int)(t / OH);
  const size_t C = (size_t)cg * 8;
  const ActT<S> p = x + ((((size_t)b * H + 2 * h) * W + 2 * w) * C + (size_t)c * 8);
  float a[8], v[8];
  ld8(p, v);
  ld8(p + C, a);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], a[j]);
  ld8(p + (size_t)W * C, a);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], a[j]);
  ld8(p + (size_t)W * C + C, a);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], a[j]);
  st8(y + idx * 8, v);
}
// routes gy to the FIRST maximum of each 2x2 window in (row, col) scan order (ATen max_pool2d backward),
// optionally gated by x > 0 (the ReLU that produced x).
template <bool S>
__global__ void maxpool2_bwd_kernel(const ActT<S> x, const ActT<S> gy, const ActT<S> gx, int B,
                                    int H, int W, int cg, int relu_gate, int accumulate) {
  const int OH = H / 2, OW = W / 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * OH * OW * cg) return;
  const int c = (int)((unsigned)idx % (unsigned)cg);  // 32-bit index split: element counts are < 2^31 (checked at launch)
  int t = (int)((unsigned)idx / (unsigned)cg);
  const int w = (int)(t % OW);
  t /= OW;
  const int h = (int)(t % OH);
  const int b = (int)(t / OH);
  const size_t C = (size_t)cg * 8;
  const size_t off = (((size_t)b * H + 2 * h) * W + 2 * w) * C + (size_t)c * 8;
  float v[4][8], g[8], o[4][8];
  ld8(x + off, v[0]);
  ld8(x + off + C, v[1]);
  ld8(x + off + (size_t)W * C, v[2]);
  ld8(x + off + (size_t)W * C + C, v[3]);
  ld8(gy + idx * 8, g);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int best = 0;
    float m = v[0][j];
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (v[k][j] > m) {
        m = v[k][j];
        best = k;
      }
    const float gv = (relu_gate && !(m > 0.f)) ? 0.f : g[j];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k][j] = (k == best) ? gv : 0.f;
  }
  if (accumulate) {
    const size_t offs[4] = {off, off + C, off + (size_t)W * C, off + (size_t)W * C + C};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float prev[8];
      ld8(gx + offs[k], prev);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[k][j] += prev[j];
    }
  }
  st8(gx + off, o[0]);
  st8(gx + off + C, o[1]);
  st8(gx + off + (size_t)W * C, o[2]);
  st8(gx + off + (size_t)W * C + C, o[3]);
}

// torch adaptive_avg_pool2d windows: [floor(i*H/OH), ceil((i+1)*H/OH))
__device__ __forceinline__ void adaptive_win(int i, int in, int out, int* s, int* e) {
  *s = (i * in) / out;
  *e = ((i + 1) * in + out - 1) / out;
}
template <bool S>
__global__ void adaptive_avgpool_fwd_kernel(const ActT<S> x, const ActT<S> y, int B, int H, int W, int OH,
                                            int OW, int cg) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * OH * OW * cg) return;
  const int c = (int)((unsigned)idx % (unsigned)cg);  // 32-bit index split: element counts are < 2^31 (checked at launch)
  int t = (int)((unsigned)idx / (unsigned)cg);
  const int ow = (int)(t % OW);
  t /= OW;
  const int oh = (int)(t % OH);
  const int b = (int)(t / OH);
  int hs, he, ws, we;
  adaptive_win(oh, H, OH, &hs, &he);
  adaptive_win(ow, W, OW, &ws, &we);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a[8];
  for (int h = hs; h < he; ++h)
    for (int w = ws; w < we; ++w) {
      ld8(x + ((((size_t)b * H + h) * W + w) * cg + c) * 8, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j];
    }
  const float inv = 1.f / (float)((he - hs) * (we - ws));
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] *= inv;
  st8(y + idx * 8, acc);
}
template <bool S>
__global__ void adaptive_avgpool_bwd_kernel(const ActT<S> gy, const ActT<S> residual,
                                            const ActT<S> gx, int B, int H, int W, int OH, int OW, int cg) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W * cg) return;
  const int c = (int)((unsigned)idx % (unsigned)cg);  // 32-bit index split: element counts are < 2^31 (checked at launch)
  int t = (int)((unsigned)idx / (unsigned)cg);
  const int w = (int)(t % W);
  t /= W;
  const int h = (int)(t % H);
  const int b = (int)(t / H);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a[8];
  for (int oh = 0; oh < OH; ++oh) {
    int hs, he;
    adaptive_win(oh, H, OH, &hs, &he);
    if (h < hs || h >= he) continue;
    for (int ow = 0; ow < OW; ++ow) {
      int ws, we;
      adaptive_win(ow, W, OW, &ws, &we);
      if (w < ws || w >= we) continue;
      const float inv = 1.f / (float)((he - hs) * (we - ws));
      ld8(gy + ((((size_t)b * OH + oh) * OW + ow) * cg + c) * 8, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j] * inv;
    }
  }
  if (!residual.null()) {
    ld8(residual + idx * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += a[j];
  }
  st8(gx + idx * 8, acc);
}

// feat[b][c] = mean_p lrelu(x[b,p,c])  (models.py:125-127); one block per image, threads over channels
template <bool S>
__global__ void global_avgpool_lrelu_fwd_kernel(const ActT<S> x, float slope, float* __restrict__ out, int P,
                                                int C) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < P; ++p) acc += lrelu_f(ldf(x, ((size_t)b * P + p) * C + c), slope);
    out[(size_t)b * C + c] = acc / (float)P;
  }
}
template <bool S>
__global__ void global_avgpool_lrelu_bwd_kernel(const ActT<S> x, const float* __restrict__ gfeat, float slope,
                                                const ActT<S> gx, int P, int C) {
  // one element per thread over a (sample, chunk) grid: a 256-thread loop per sample was a 12 us serial chain at the head
  // of every discriminator backward
  const int b = blockIdx.x;
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= P * C) return;
  const int c = i % C;
  const float xv = ldf(x, (size_t)b * P * C + i);
  const float g = gfeat[(size_t)b * C + c] / (float)P;
  stf(gx, (size_t)b * P * C + i, xv > 0.f ? g : g * slope);
}

// ---------------------------------------------------------------------------------------------
// out = gamma * t + x   and its backward  (SelfAttention tail, models.py:274)
// ---------------------------------------------------------------------------------------------
template <bool S>
__global__ void gamma_residual_fwd_kernel(const ActT<S> t, const ActT<S> x,
                                          const float* __restrict__ gamma, const ActT<S> out,
                                          const ActT<S> out_act, float slope, long long n8) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  const float g = __ldg(gamma);
  float a[8], b[8];
  ld8(t + idx * 8, a);
  ld8(x + idx * 8, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = g * a[j] + b[j];
  st8(out + idx * 8, a);
  if (!out_act.null()) {
    // activate the value the raw output holds (BF16-rounded, or hi + lo in split mode)
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = lrelu_f(stored_value(a[j], out), slope);
    st8(out_act + idx * 8, a);
  }
}
// gt = gamma * g ; dgamma += sum(g * t)
template <bool S>
__global__ void gamma_residual_bwd_kernel(const ActT<S> g, const ActT<S> t,
                                          const float* __restrict__ gamma, const ActT<S> gt,
                                          float* __restrict__ dgamma, long long n8, float* __restrict__ scratch,
                                          unsigned int* ticket) {
  const float gm = __ldg(gamma);
  float acc = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n8; idx += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    ld8(g + idx * 8, a);
    ld8(t + idx * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc += a[j] * b[j];
      a[j] *= gm;
    }
    st8(gt + idx * 8, a);
  }
  acc = warp_sum(acc);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) scratch[blockIdx.x] = v;
  }
  if (spyr_last_block(ticket, gridDim.x))
    spyr_sum_partials<float>(scratch, (int)gridDim.x, 1, [&](int, float total) { *dgamma += total; });
}

// out_i[c] += sum_rows g[row][c]   (bias gradients; up to three identical destinations)
template <bool S>
__global__ void colsum_kernel(const ActT<S> g, long long rows, int cg, float* __restrict__ scratch) {
  extern __shared__ float sh[];  // [prows][cg*8]
  const int c = threadIdx.x % cg;
  const int pr = threadIdx.x / cg;
  const int prows = blockDim.x / cg;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a[8];
  if (pr < prows)
#pragma unroll 8  // eight independent 16-byte loads in flight per thread (a rolled loop has one)
    for (long long r = (long long)blockIdx.x * prows + pr; r < rows; r += (long long)gridDim.x * prows) {
      ld8_nc(g + (r * cg + c) * 8, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j];
    }
  if (pr < prows) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[(pr * cg + c) * 8 + j] = acc[j];
  }
  __syncthreads();
  const int C = cg * 8;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < prows; ++r) s += sh[r * C + i];
    scratch[(size_t)blockIdx.x * C + i] = s;
  }
}

// dw[(t*cin_stride + ci_row)*Cout + co] += sum_{b,h,w} mask[b,h+dy,w+dx] * g[b,h,w,co]
// (weight gradient of the mask channel of `cat(feature*mask, mask)`, models.py:94,336)
template <bool S>
__global__ void stencil_wgrad_kernel(const float* __restrict__ mask, const ActT<S> g, int B, int H, int W,
                                     int cg, float* __restrict__ scratch) {
  extern __shared__ float sh[];  // [prows][9][C]
  const int C = cg * 8;
  const int c = threadIdx.x % cg;
  const int pr = threadIdx.x / cg;
  const int prows = blockDim.x / cg;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  float acc_all[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long npix = (long long)B * H * W;
  if (pr < prows)
    for (long long p = (long long)blockIdx.x * prows + pr; p < npix; p += (long long)gridDim.x * prows) {
      const int w = (int)(p % W);
      const int h = (int)((p / W) % H);
      const long long bbase = p - (long long)h * W - w;
      float mk[9];
      bool any = false, all = true;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
        mk[t] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(mask + bbase + (long long)hh * W + ww) : 0.f;
        any = any || (mk[t] != 0.f);
        all = all && (mk[t] == 1.f);
      }
      if (!any) continue;
      float a[8];
      ld8(g + (p * cg + c) * 8, a);
      if (all) {
        // interior of a kept region (masks are 0/1 and mostly constant per sample): one sum serves all nine taps
#pragma unroll
        for (int j = 0; j < 8; ++j) acc_all[j] += a[j];
        continue;
      }
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] += mk[t] * a[j];
    }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] += acc_all[j];
  if (pr < prows) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) sh[(pr * 9 + t) * C + c * 8 + j] = acc[t][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < prows; ++r) s += sh[r * 9 * C + i];
    scratch[(size_t)blockIdx.x * 9 * C + i] = s;
  }
}

// dst[t][n][k] = src[taps-1-t][k][n]: the packed forward weights [tap][Cout][Cin] re-laid out as the K-major operand of
// the input-gradient convolution (taps flipped, channel roles swapped).  Only used for 64-wide outputs, where the
// K-major form lets the CTA-pair kernel run the layer (a few hundred KB at most).
__global__ void weight_transpose_flip_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int taps, int K, int N) {
  __shared__ bf16 tile[32][33];
  const int t = blockIdx.z;
  const bf16* s = src + (size_t)(taps - 1 - t) * K * N;
  bf16* d = dst + (size_t)t * N * K;
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, n = n0 + threadIdx.x;
    if (k < K && n < N) tile[r][threadIdx.x] = s[(size_t)k * N + n];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, k = k0 + threadIdx.x;
    if (n < N && k < K) d[(size_t)n * K + k] = tile[threadIdx.x][r];
  }
}

// out[co][ci][t] = gw[(t * cin_stride + ci) * Cout + co]: tensor-core weight gradient ([tap][Cin][Cout], FP32) -> the
// parameter's own (Cout, Cin, kh, kw) layout.  Spectral-normed layers get this from sn_backward; plain layers (VGG
// fine-tuning, vgg_16_train.py:164) from here.  ACC: out += (gradient accumulation over batch chunks).
__global__ void wgrad_to_oihw_kernel(const float* __restrict__ gw, float* __restrict__ out, int taps, int Cin, int Cout,
                                     int cin_stride, int accumulate) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    if (ci < Cin && co < Cout) tile[r][threadIdx.x] = gw[((size_t)t * cin_stride + ci) * Cout + co];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    if (co < Cout && ci < Cin) {
      float* o = out + ((size_t)co * Cin + ci) * taps + t;
      *o = accumulate ? *o + tile[threadIdx.x][r] : tile[threadIdx.x][r];
    }
  }
}

// inverted dropout (nn.Dropout, torchvision VGG classifier[2], [5]): keep with probability 1 - p, scale by 1 / (1 - p).
// Counter-based generator: element i of call (seed, offset) always draws the same number, independent of the launch shape.
__device__ __forceinline__ uint32_t mix32(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}
template <bool S>
__global__ void dropout_fwd_kernel(const float* __restrict__ x, long long n, float p, unsigned long long seed,
                                   unsigned long long offset, float* __restrict__ y, const ActT<S> y_bf16,
                                   unsigned char* __restrict__ mask) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float u = (float)(mix32(seed * 0x100000001B3ull + offset + (unsigned long long)i) >> 8) * (1.0f / 16777216.0f);
  const bool keep = u >= p;
  const float v = keep ? x[i] / (1.f - p) : 0.f;
  mask[i] = keep ? 1 : 0;
  if (y != nullptr) y[i] = v;
  if (!y_bf16.null()) stf(y_bf16, (size_t)i, v);
}
__global__ void dropout_bwd_kernel(const float* __restrict__ g, const unsigned char* __restrict__ mask, long long n, float p,
                                   float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = mask[i] ? g[i] / (1.f - p) : 0.f;
}

template <bool S>
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, const ActT<S> dst, long long n) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) stf(dst, (size_t)idx, src[idx]);
}

template <bool S>
__global__ void vec_epilogue_kernel(const float* __restrict__ acc, int nsplit, const float* __restrict__ bias,
                                    const float* __restrict__ add, const float* __restrict__ gate, int mode,
                                    float* __restrict__ out_f32, const ActT<S> out_bf16, int ld_bf16, int B, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  float v = acc[i];
  for (int sp = 1; sp < nsplit; ++sp) v += acc[(size_t)sp * B * N + i];  // split-K partial sums, fixed order
  if (bias != nullptr) v += bias[i % N];
  if (add != nullptr) v += add[i];
  if (mode == 1) v = fmaxf(v, 0.f);
  if (mode == 2 && !(gate[i] > 0.f)) v = 0.f;
  if (out_f32 != nullptr) out_f32[i] = v;
  if (!out_bf16.null()) stf(out_bf16, (size_t)(i / N) * ld_bf16 + i % N, v);
}

// ---------------------------------------------------------------------------------------------
// Epilogue of a split-K convolution on a small map: the FP32 partial accumulators of the K splits (one slice per split,
// summed here in split order -- no atomics, bit-reproducible) get
// the same fused tail as the in-kernel epilogue of spyr_conv2d_fprop: biases, mask-channel stencil, gate, residual,
// raw + activated BF16 outputs.
// ---------------------------------------------------------------------------------------------
template <bool S>
__global__ void conv_epilogue_kernel(const float* __restrict__ acc, int nsplit, int B, int H, int W, int cg,
                                     const float* __restrict__ bias, const float* __restrict__ bias2,
                                     const float* __restrict__ bias3, const float* __restrict__ stencil_mask,
                                     const float* __restrict__ stencil_w, const ActT<S> dmask, float dmask_slope,
                                     const ActT<S> residual, const ActT<S> y_raw, const ActT<S> y_act,
                                     int act, float act_slope) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W * cg) return;
  const int c = (int)((unsigned)idx % (unsigned)cg);  // 32-bit index split: element counts are < 2^31 (checked at launch)
  const long long pix = idx / cg;
  const int C = cg * 8;
  float v[8];
  const float4 a0 = *reinterpret_cast<const float4*>(acc + idx * 8);
  const float4 a1 = *reinterpret_cast<const float4*>(acc + idx * 8 + 4);
  v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
  const size_t slice = (size_t)B * H * W * C;
  for (int sp = 1; sp < nsplit; ++sp) {
    const float4 b0 = *reinterpret_cast<const float4*>(acc + sp * slice + idx * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(acc + sp * slice + idx * 8 + 4);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = c * 8 + j;
    if (bias != nullptr) v[j] += __ldg(&bias[col]);
    if (bias2 != nullptr) v[j] += __ldg(&bias2[col]);
    if (bias3 != nullptr) v[j] += __ldg(&bias3[col]);
  }
  if (stencil_mask != nullptr) {
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const long long bbase = pix - (long long)h * W - w;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const float m = __ldg(&stencil_mask[bbase + (long long)hh * W + ww]);
      if (m != 0.f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += m * __ldg(&stencil_w[t * C + c * 8 + j]);
      }
    }
  }
  if (!dmask.null()) {
    float d[8];
    ld8(dmask.p + idx * 8, d);  // the sign of a value is the sign of its hi part
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(d[j] > 0.f)) v[j] *= dmask_slope;
  }
  if (!residual.null()) {
    float r[8];
    ld8(residual + idx * 8, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += r[j];
  }
  if (!y_raw.null()) st8(y_raw + idx * 8, v);
  if (!y_act.null()) {
    const float sl = (act == 1) ? 0.f : ((act == 2) ? act_slope : 1.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * sl;
    st8(y_act + idx * 8, v);
  }
}

// ---------------------------------------------------------------------------------------------
// generator tail: img = tanh(conv1x1(a; W/sigma) + b), C -> 3 channels, NCHW f32 out  (models.py:58-61,99)
// ---------------------------------------------------------------------------------------------
template <int CO, bool S>
__global__ void conv1x1_tanh_fwd_kernel(const ActT<S> a, const float* __restrict__ w,
                                        const float* __restrict__ sigma, const float* __restrict__ bias,
                                        float* __restrict__ img, int HW, int C, long long npix) {
  extern __shared__ float ws[];  // [CO][C]
  const float inv = 1.f / __ldg(sigma);
  for (int i = threadIdx.x; i < CO * C; i += blockDim.x) ws[i] = w[i] * inv;
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = bias[o];
  for (int k = 0; k < C; k += 8) {
    float v[8];
    ld8(a + p * C + k, v);
#pragma unroll
    for (int o = 0; o < CO; ++o)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[o] += v[j] * ws[o * C + k + j];
  }
  const long long b = p / HW, q = p % HW;
#pragma unroll
  for (int o = 0; o < CO; ++o) img[((size_t)b * CO + o) * HW + q] = tanhf(acc[o]);
}
// gh[p][k] = (sum_o gpre[o] W[o][k]/sigma) * lrelu'(a[p][k]);  dW_sn[o][k] += sum_p gpre[o] a[p][k];  db[o] += sum_p gpre[o]
template <int CO, bool S>
__global__ void conv1x1_tanh_bwd_kernel(const float* __restrict__ gimg, const float* __restrict__ img,
                                        const ActT<S> a, const float* __restrict__ w,
                                        const float* __restrict__ sigma, float slope, const ActT<S> gh,
                                        int HW, int C, long long npix, float* __restrict__ scratch) {
  extern __shared__ float sh[];  // ws[CO][C] then red[prows][CO][C] then redb[prows][CO]
  const int cg = C / 8;
  const int c = threadIdx.x % cg;
  const int pr = threadIdx.x / cg;
  const int prows = blockDim.x / cg;
  float* ws = sh;
  float* red = sh + CO * C;
  float* redb = red + prows * CO * C;
  const float inv = 1.f / __ldg(sigma);
  for (int i = threadIdx.x; i < CO * C; i += blockDim.x) ws[i] = w[i] * inv;
  __syncthreads();
  float accw[CO][8], accb[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) {
    accb[o] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) accw[o][j] = 0.f;
  }
  if (pr < prows)
    for (long long p = (long long)blockIdx.x * prows + pr; p < npix; p += (long long)gridDim.x * prows) {
      const long long b = p / HW, q = p % HW;
      float gp[CO];
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        const float y = __ldg(img + ((size_t)b * CO + o) * HW + q);
        gp[o] = __ldg(gimg + ((size_t)b * CO + o) * HW + q) * (1.f - y * y);
      }
      float v[8], o8[8];
      ld8(a + p * C + c * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          s += gp[o] * ws[o * C + c * 8 + j];
          accw[o][j] += gp[o] * v[j];
        }
        o8[j] = v[j] > 0.f ? s : s * slope;
      }
      st8(gh + p * C + c * 8, o8);
      if (c == 0) {
#pragma unroll
        for (int o = 0; o < CO; ++o) accb[o] += gp[o];
      }
    }
  if (pr < prows) {
#pragma unroll
    for (int o = 0; o < CO; ++o) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(pr * CO + o) * C + c * 8 + j] = accw[o][j];
      if (c == 0) redb[pr * CO + o] = accb[o];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CO * C; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < prows; ++r) s += red[r * CO * C + i];
    scratch[(size_t)blockIdx.x * (CO * C + CO) + i] = s;
  }
  if (threadIdx.x < CO) {
    float s = 0.f;
    for (int r = 0; r < prows; ++r) s += redb[r * CO + threadIdx.x];
    scratch[(size_t)blockIdx.x * (CO * C + CO) + CO * C + threadIdx.x] = s;
  }
}

}  // namespace

#define SPYR_C8(C) SPYR_REQUIRE((C) > 0 && (C) % 8 == 0, "%s: channel count %d must be a multiple of 8", __func__, (int)(C))

extern "C" int spyr_im2col3x3(const float* img, int B, int H, int W, const float* mean3, const float* invstd3, void* out,
                              void* stream) {
  SPYR_REQUIRE(img && out && B > 0 && H > 0 && W > 0, "im2col3x3: bad arguments");
  const long long n = (long long)B * H * W;
  SPYR_WITH_SPLIT(im2col3x3_kernel<kS><<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(img, B, H, W, mean3, invstd3, make_act(out, n * 32)));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_col2im3x3(const void* gcol, int B, int H, int W, const float* invstd3, float* gimg, int accumulate,
                              void* stream) {
  SPYR_REQUIRE(gcol && gimg && B > 0, "col2im3x3: bad arguments");
  const long long n = (long long)B * H * W;
  SPYR_WITH_SPLIT(col2im3x3_kernel<kS><<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(make_act(gcol, n * 32), B, H, W, invstd3, gimg,
                                                                       accumulate));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_img_avgpool_pad8(const float* img, int B, int H, int W, void* out, void* stream) {
  SPYR_REQUIRE(img && out && H % 2 == 0 && W % 2 == 0, "img_avgpool_pad8: bad arguments");
  const long long n = (long long)B * (H / 2) * (W / 2);
  SPYR_WITH_SPLIT(img_avgpool_pad8_kernel<kS><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(img, B, H, W, make_act(out, n * 8)));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_img_avgpool_pad8_bwd(const void* g8, int B, int H, int W, float* gimg, int accumulate, void* stream) {
  SPYR_REQUIRE(g8 && gimg && H % 2 == 0 && W % 2 == 0, "img_avgpool_pad8_bwd: bad arguments");
  const long long n = (long long)B * H * W;
  SPYR_WITH_SPLIT(img_avgpool_pad8_bwd_kernel<kS><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(make_act(g8, n * 2), B, H, W, gimg,
                                                                                  accumulate));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_nchw_to_nhwc(const float* src, const float* mask, float slope, void* dst, int B, int C, int HW,
                                 void* stream) {
  SPYR_REQUIRE(src && dst && B > 0 && C > 0 && HW > 0, "nchw_to_nhwc: bad arguments");
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B), block(32, 8);
  SPYR_WITH_SPLIT(nchw_to_nhwc_kernel<kS><<<grid, block, 0, (cudaStream_t)stream>>>(src, mask, slope, make_act(dst, (long long)B * C * HW), C, HW));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_nhwc_to_nchw(const void* src, const float* gate_x, float slope, float* dst, int B, int C, int HW,
                                 void* stream) {
  SPYR_REQUIRE(src && dst && B > 0 && C > 0 && HW > 0, "nhwc_to_nchw: bad arguments");
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B), block(32, 8);
  SPYR_WITH_SPLIT(nhwc_to_nchw_kernel<kS><<<grid, block, 0, (cudaStream_t)stream>>>(make_act(src, (long long)B * C * HW), gate_x, slope, dst, C, HW));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_maskgate(const void* f, const float* mask, void* out, long long npix, int C, void* stream) {
  SPYR_C8(C);
  SPYR_WITH_SPLIT(maskgate_kernel<kS><<<grid_for(npix * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(make_act(f, npix * C), mask, make_act(out, npix * C), npix,
                                                                                   C / 8));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_avgpool2_fwd(const void* x, const void* residual, void* y_raw, void* y_act, float slope, int B, int H,
                                 int W, int C, void* stream) {
  SPYR_C8(C);
  SPYR_REQUIRE(H % 2 == 0 && W % 2 == 0, "avgpool2_fwd: odd size");
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
  SPYR_N32(n);
  SPYR_WITH_SPLIT(avgpool2_fwd_kernel<kS><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(make_act(x, n * 32), make_act(residual, n * 8),
                                                                          make_act(y_raw, n * 8), make_act(y_act, n * 8), slope, B, H, W, C / 8));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_avgpool2_bwd(const void* g_lo, void* g_hi, int B, int H, int W, int C, void* stream) {
  SPYR_C8(C);
  const long long n = (long long)B * H * W * (C / 8);
  SPYR_N32(n);
  SPYR_WITH_SPLIT(avgpool2_bwd_kernel<kS><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(make_act(g_lo, n * 2), make_act(g_hi, n * 8), B, H, W, C / 8));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_maxpool2_fwd(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  SPYR_C8(C);
  SPYR_REQUIRE(H % 2 == 0 && W % 2 == 0, "maxpool2_fwd: odd size");
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
  SPYR_N32(n);
  SPYR_WITH_SPLIT(maxpool2_fwd_kernel<kS><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(make_act(x, n * 32), make_act(y, n * 8), B, H, W, C / 8));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_maxpool2_bwd(const void* x, const void* gy, void* gx, int B, int H, int W, int C, int relu_gate,
                                 int accumulate, void* stream) {
  SPYR_C8(C);
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
  SPYR_N32(n);
  SPYR_WITH_SPLIT(maxpool2_bwd_kernel<kS><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(make_act(x, n * 32), make_act(gy, n * 8), make_act(gx, n * 32), B, H,
                                                                          W, C / 8, relu_gate, accumulate));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_adaptive_avgpool_fwd(const void* x, void* y, int B, int H, int W, int OH, int OW, int C, void* stream) {
  SPYR_C8(C);
  const long long n = (long long)B * OH * OW * (C / 8);
  SPYR_N32(n);
  SPYR_WITH_SPLIT(adaptive_avgpool_fwd_kernel<kS><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(make_act(x, (long long)B * H * W * C), make_act(y, n * 8), B, H, W, OH, OW,
                                                                                  C / 8));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_adaptive_avgpool_bwd(const void* gy, const void* residual, void* gx, int B, int H, int W, int OH,
                                         int OW, int C, void* stream) {
  SPYR_C8(C);
  const long long n = (long long)B * H * W * (C / 8);
  SPYR_N32(n);
  SPYR_WITH_SPLIT(adaptive_avgpool_bwd_kernel<kS><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      make_act(gy, (long long)B * OH * OW * C), make_act(residual, n * 8), make_act(gx, n * 8), B, H, W, OH, OW, C / 8));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_global_avgpool_lrelu_fwd(const void* x, float slope, float* out, int B, int P, int C, void* stream) {
  SPYR_WITH_SPLIT(global_avgpool_lrelu_fwd_kernel<kS><<<B, 256, 0, (cudaStream_t)stream>>>(make_act(x, (long long)B * P * C), slope, out, P, C));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_global_avgpool_lrelu_bwd(const void* x, const float* gfeat, float slope, void* gx, int B, int P, int C,
                                             void* stream) {
  SPYR_WITH_SPLIT(global_avgpool_lrelu_bwd_kernel<kS><<<dim3(B, ceil_div(P * C, 256)), 256, 0, (cudaStream_t)stream>>>(
      make_act(x, (long long)B * P * C), gfeat, slope, make_act(gx, (long long)B * P * C), P, C));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_gamma_residual_fwd(const void* t, const void* x, const float* gamma, void* out, void* out_act,
                                       float slope, long long n, void* stream) {
  SPYR_REQUIRE(n % 8 == 0, "gamma_residual_fwd: n must be a multiple of 8");
  SPYR_WITH_SPLIT(gamma_residual_fwd_kernel<kS><<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
      make_act(t, n), make_act(x, n), gamma, make_act(out, n), make_act(out_act, n), slope, n / 8));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_gamma_residual_bwd(const void* g, const void* t, const float* gamma, void* gt, float* dgamma,
                                       long long n, void* scratch, void* stream) {
  SPYR_REQUIRE(n % 8 == 0, "gamma_residual_bwd: n must be a multiple of 8");
  SPYR_REQUIRE(scratch != nullptr, "gamma_residual_bwd: scratch is NULL");
  int grid = grid_for(n / 8, 256);
  if (grid > SPYR_REDUCE_BLOCKS) grid = SPYR_REDUCE_BLOCKS;
  SPYR_WITH_SPLIT(gamma_residual_bwd_kernel<kS><<<grid, 256, 0, (cudaStream_t)stream>>>(make_act(g, n), make_act(t, n), gamma, make_act(gt, n),
                                                                    dgamma, n / 8, (float*)scratch, spyr_next_ticket()));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
static int colsum_impl(const void* g, long long rows, int C, float* out0, float* out1, float* out2, void* scratch,
                       spyr_reduce_entry* deferred, void* stream);
extern "C" int spyr_colsum(const void* g, long long rows, int C, float* out0, float* out1, float* out2, void* scratch,
                           void* stream) {
  return colsum_impl(g, rows, C, out0, out1, out2, scratch, nullptr, stream);
}
extern "C" int spyr_colsum_deferred(const void* g, long long rows, int C, float* out0, float* out1, float* out2,
                                    void* scratch, spyr_reduce_entry* out, void* stream) {
  SPYR_REQUIRE(out != nullptr, "colsum_deferred: out is NULL");
  memset(out, 0, sizeof(*out));
  return colsum_impl(g, rows, C, out0, out1, out2, scratch, out, stream);
}
static void defer_sum(spyr_reduce_entry* e, const float* scratch, int nb, int n, const SumSink& sk) {
  e->kind = 2;
  e->partial = scratch;
  e->out0 = sk.out0; e->out1 = sk.out1; e->out2 = sk.out2;
  e->nb = nb; e->n = n; e->mode = sk.mode; e->C = sk.C; e->sink_stride = sk.cin_stride; e->sink_row = sk.ci_row;
  e->split = sk.split;
}
static int colsum_impl(const void* g, long long rows, int C, float* out0, float* out1, float* out2, void* scratch,
                       spyr_reduce_entry* deferred, void* stream) {
  SPYR_REQUIRE(scratch != nullptr, "colsum: scratch is NULL");
  SPYR_C8(C);
  const int cg = C / 8;
  SPYR_REQUIRE(cg <= 256, "colsum: C=%d too large", C);
  const int prows = 256 / cg;
  const int threads = prows * cg;
  long long want = (rows + prows * 8 - 1) / (prows * 8);
  // 4-byte partials: twice SPYR_REDUCE_BLOCKS vectors fit the 8-byte-slot scratch
  int grid = (int)(want < 1 ? 1 : (want > 2 * SPYR_REDUCE_BLOCKS ? 2 * SPYR_REDUCE_BLOCKS : want));
  SPYR_WITH_SPLIT(colsum_kernel<kS><<<grid, threads, (size_t)prows * C * 4, (cudaStream_t)stream>>>(make_act(g, rows * C), rows, cg,
                                                                                (float*)scratch));
  spyr_count_launch();
  SumSink sk = {out0, out1, out2, 0, C, 0, 0, 0};
  if (deferred != nullptr) {
    defer_sum(deferred, (const float*)scratch, grid, C, sk);
    SPYR_LAUNCH_CHECK();
    return 0;
  }
  SPYR_CHECK_CUDA(spyr_launch_sum_partials((const float*)scratch, grid, C, sk, (cudaStream_t)stream));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
static int stencil_impl(const float* mask, const void* g, int B, int H, int W, int C, float* dw, int cin_stride, int ci_row,
                        void* scratch, spyr_reduce_entry* deferred, void* stream);
extern "C" int spyr_stencil_wgrad(const float* mask, const void* g, int B, int H, int W, int C, float* dw, int cin_stride,
                                  int ci_row, void* scratch, void* stream) {
  return stencil_impl(mask, g, B, H, W, C, dw, cin_stride, ci_row, scratch, nullptr, stream);
}
extern "C" int spyr_stencil_wgrad_deferred(const float* mask, const void* g, int B, int H, int W, int C, float* dw,
                                           int cin_stride, int ci_row, void* scratch, spyr_reduce_entry* out, void* stream) {
  SPYR_REQUIRE(out != nullptr, "stencil_wgrad_deferred: out is NULL");
  memset(out, 0, sizeof(*out));
  return stencil_impl(mask, g, B, H, W, C, dw, cin_stride, ci_row, scratch, out, stream);
}
static int stencil_impl(const float* mask, const void* g, int B, int H, int W, int C, float* dw, int cin_stride, int ci_row,
                        void* scratch, spyr_reduce_entry* deferred, void* stream) {
  SPYR_REQUIRE(scratch != nullptr, "stencil_wgrad: scratch is NULL");
  SPYR_C8(C);
  const int cg = C / 8;
  SPYR_REQUIRE(cg <= 128, "stencil_wgrad: C=%d too large", C);
  int prows = 128 / cg;
  if (prows > 4) prows = 4;
  const int threads = prows * cg;
  const size_t smem = (size_t)prows * 9 * C * 4;
  static bool configured = false;
  if (!configured) {
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(stencil_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(stencil_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  const long long npix = (long long)B * H * W;
  long long want = (npix + prows * 16 - 1) / (prows * 16);
  int grid = (int)(want < 1 ? 1 : (want > 2 * SPYR_REDUCE_BLOCKS ? 2 * SPYR_REDUCE_BLOCKS : want));
  SPYR_WITH_SPLIT(stencil_wgrad_kernel<kS><<<grid, threads, smem, (cudaStream_t)stream>>>(mask, make_act(g, npix * C), B, H, W, cg,
                                                                      (float*)scratch));
  spyr_count_launch();
  SumSink sk = {dw, nullptr, nullptr, 1, C, cin_stride, ci_row, 0};
  if (deferred != nullptr) {
    defer_sum(deferred, (const float*)scratch, grid, 9 * C, sk);
    SPYR_LAUNCH_CHECK();
    return 0;
  }
  SPYR_CHECK_CUDA(spyr_launch_sum_partials((const float*)scratch, grid, 9 * C, sk, (cudaStream_t)stream));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_weight_transpose_flip(const void* src, void* dst, int taps, int K, int N, void* stream) {
  SPYR_REQUIRE(src && dst && taps > 0 && K > 0 && N > 0, "weight_transpose_flip: bad arguments");
  dim3 grid(ceil_div(N, 32), ceil_div(K, 32), taps);
  weight_transpose_flip_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const bf16*)src, (bf16*)dst, taps, K, N);
  spyr_count_launch();
  if (spyr_split()) {  // the lo plane of the pack follows its hi plane
    const size_t n = (size_t)taps * K * N;
    weight_transpose_flip_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const bf16*)src + n, (bf16*)dst + n, taps,
                                                                                 K, N);
    spyr_count_launch();
  }
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_wgrad_to_oihw(const float* gw, float* out, int taps, int Cin, int Cout, int cin_stride, int accumulate,
                                  void* stream) {
  SPYR_REQUIRE(gw && out && taps > 0 && Cin > 0 && Cout > 0 && cin_stride >= Cin, "wgrad_to_oihw: bad arguments");
  dim3 grid(ceil_div(Cout, 32), ceil_div(Cin, 32), taps);
  wgrad_to_oihw_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(gw, out, taps, Cin, Cout, cin_stride, accumulate);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_dropout_fwd(const float* x, long long n, float p, unsigned long long seed, unsigned long long offset,
                                float* y, void* y_bf16, unsigned char* mask, void* stream) {
  SPYR_REQUIRE(x && mask && n > 0 && p >= 0.f && p < 1.f, "dropout_fwd: bad arguments (0 <= p < 1)");
  SPYR_WITH_SPLIT(dropout_fwd_kernel<kS><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, p, seed, offset, y, make_act(y_bf16, n), mask));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_dropout_bwd(const float* g, const unsigned char* mask, long long n, float p, float* out, void* stream) {
  SPYR_REQUIRE(g && mask && out && n > 0 && p >= 0.f && p < 1.f, "dropout_bwd: bad arguments (0 <= p < 1)");
  dropout_bwd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(g, mask, n, p, out);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
  SPYR_WITH_SPLIT(cast_f32_bf16_kernel<kS><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(src, make_act(dst, n), n));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_vec_epilogue(const float* acc, int nsplit, const float* bias, const float* add, const float* gate,
                                 int mode, float* out_f32, void* out_bf16, int ld_bf16, int B, int N, void* stream) {
  SPYR_REQUIRE(acc && nsplit >= 1 && (mode != 2 || gate) && (out_bf16 == nullptr || ld_bf16 >= N),
               "vec_epilogue: bad arguments");
  SPYR_WITH_SPLIT(vec_epilogue_kernel<kS><<<grid_for((long long)B * N, 256), 256, 0, (cudaStream_t)stream>>>(
      acc, nsplit, bias, add, gate, mode, out_f32, make_act(out_bf16, (long long)B * ld_bf16), ld_bf16, B, N));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_conv2d_epilogue(const spyr_conv_desc* d, const float* acc, void* stream) {
  SPYR_REQUIRE(d != nullptr && acc != nullptr, "conv2d_epilogue: bad arguments");
  SPYR_C8(d->Cout);
  const long long n = (long long)d->B * d->H * d->W * (d->Cout / 8);
  SPYR_WITH_SPLIT(conv_epilogue_kernel<kS><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      acc, d->splits < 1 ? 1 : d->splits, d->B, d->H, d->W, d->Cout / 8, d->bias, d->bias2, d->bias3, d->stencil_mask,
      d->stencil_w, make_act(d->dmask, n * 8), d->dmask_slope, make_act(d->residual, n * 8), make_act(d->y_raw, n * 8),
      make_act(d->y_act, n * 8), d->act, d->act_slope));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_conv1x1_tanh_fwd(const void* a, const float* w, const float* sigma, const float* bias, float* img,
                                     int B, int HW, int C, int Cout, void* stream) {
  SPYR_C8(C);
  SPYR_REQUIRE(Cout == 3 || Cout == 1, "conv1x1_tanh_fwd: out_channels must be 1 or 3 (got %d)", Cout);
  const long long npix = (long long)B * HW;
  const size_t smem = (size_t)Cout * C * 4;
  if (Cout == 3)
    SPYR_WITH_SPLIT(conv1x1_tanh_fwd_kernel<3, kS><<<grid_for(npix, 128), 128, smem, (cudaStream_t)stream>>>(make_act(a, npix * C), w, sigma, bias, img,
                                                                                         HW, C, npix));
  else
    SPYR_WITH_SPLIT(conv1x1_tanh_fwd_kernel<1, kS><<<grid_for(npix, 128), 128, smem, (cudaStream_t)stream>>>(make_act(a, npix * C), w, sigma, bias, img,
                                                                                         HW, C, npix));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_conv1x1_tanh_bwd(const float* gimg, const float* img, const void* a, const float* w, const float* sigma,
                                     float slope, void* gh, float* dw, float* db, int B, int HW, int C, int Cout,
                                     void* scratch, void* stream) {
  SPYR_REQUIRE(scratch != nullptr, "conv1x1_tanh_bwd: scratch is NULL");
  SPYR_C8(C);
  SPYR_REQUIRE(Cout == 3 || Cout == 1, "conv1x1_tanh_bwd: out_channels must be 1 or 3 (got %d)", Cout);
  const int cg = C / 8;
  SPYR_REQUIRE(cg <= 64, "conv1x1_tanh_bwd: C=%d too large", C);
  const int prows = 256 / cg;
  const int threads = prows * cg;
  const long long npix = (long long)B * HW;
  const size_t smem = ((size_t)Cout * C + (size_t)prows * Cout * C + (size_t)prows * Cout) * 4;
  SPYR_REQUIRE(smem <= 48 * 1024, "conv1x1_tanh_bwd: smem %zu too large", smem);
  long long want = (npix + prows * 16 - 1) / (prows * 16);
  int grid = (int)(want < 1 ? 1 : (want > 2 * SPYR_REDUCE_BLOCKS ? 2 * SPYR_REDUCE_BLOCKS : want));
  if (Cout == 3)
    SPYR_WITH_SPLIT(conv1x1_tanh_bwd_kernel<3, kS><<<grid, threads, smem, (cudaStream_t)stream>>>(gimg, img, make_act(a, npix * C), w, sigma, slope,
                                                                              make_act(gh, npix * C), HW, C, npix,
                                                                              (float*)scratch));
  else
    SPYR_WITH_SPLIT(conv1x1_tanh_bwd_kernel<1, kS><<<grid, threads, smem, (cudaStream_t)stream>>>(gimg, img, make_act(a, npix * C), w, sigma, slope,
                                                                              make_act(gh, npix * C), HW, C, npix,
                                                                              (float*)scratch));
  spyr_count_launch();
  SumSink sk = {dw, db, nullptr, 2, C, 0, 0, Cout * C};
  SPYR_CHECK_CUDA(spyr_launch_sum_partials((const float*)scratch, grid, Cout * C + Cout, sk, (cudaStream_t)stream));
  SPYR_LAUNCH_CHECK();
  return 0;
}
