// Small-batch ("skinny") fully connected layers in FP32 on the CUDA cores: M = batch <= 32 rows, so the work is
// reading the weight matrix once -- HBM/L2-bound, no tensor-core tile can be filled.
//
// Replaces nn.Linear at models.py:28-31 (Generator.linear_layer), models.py:356,360 (LinearBlock),
// models.py:128,132 (Discriminator head / classification) and the (B,B,128) projection of models.py:151-155.
// Spectral normalisation is folded in: y = (x W^T) / sigma + b with sigma read from device memory.
#include "common.cuh"
#include "../../include/spyramid_b200.h"

extern void spyr_count_launch();

namespace {

constexpr int LB = 8;        // batch rows per CTA
constexpr int LROWS = 32;    // output rows per CTA (8 warps x 4)
constexpr int LKC = 256;     // k chunk staged in shared memory

__device__ __forceinline__ float in_transform(float x, const float* mask, size_t i, float slope) {
  if (mask != nullptr) x *= mask[i];
  return x > 0.f ? x : x * slope;  // slope == 1 -> identity
}

// y[b][o] = (sum_k f(x[b][k]) W[o][k]) * inv_sigma + bias[o]  (+ y_add[b][o]) ; optional LeakyReLU on the output
__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ xmask, float in_slope,
                                  const float* __restrict__ w, const float* __restrict__ sigma,
                                  const float* __restrict__ bias, const float* __restrict__ y_add, float out_slope,
                                  float* __restrict__ y, int B, int K, int O) {
  __shared__ float xs[LB][LKC];
  const int o0 = blockIdx.x * LROWS, b0 = blockIdx.y * LB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[4][LB];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int b = 0; b < LB; ++b) acc[r][b] = 0.f;
  for (int k0 = 0; k0 < K; k0 += LKC) {
    __syncthreads();
    for (int i = threadIdx.x; i < LB * LKC; i += blockDim.x) {
      const int b = i / LKC, k = k0 + i % LKC;
      float v = 0.f;
      if (b0 + b < B && k < K) v = in_transform(x[(size_t)(b0 + b) * K + k], xmask, (size_t)(b0 + b) * K + k, in_slope);
      xs[b][i % LKC] = v;
    }
    __syncthreads();
    // k outer, rows inner: the LB activations of a k are read from shared memory once and serve the warp's four rows
#pragma unroll
    for (int j = 0; j < LKC / 32; ++j) {
      const int kk = j * 32 + lane;
      const bool ok = k0 + kk < K;
      float wv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int o = o0 + warp * 4 + r;
        wv[r] = (ok && o < O) ? __ldg(w + (size_t)o * K + k0 + kk) : 0.f;
      }
#pragma unroll
      for (int b = 0; b < LB; ++b) {
        const float xv = xs[b][kk];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][b] += wv[r] * xv;
      }
    }
  }
  const float inv = sigma != nullptr ? 1.f / __ldg(sigma) : 1.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int o = o0 + warp * 4 + r;
#pragma unroll
    for (int b = 0; b < LB; ++b) {
      const float s = warp_sum(acc[r][b]);
      if (lane == 0 && o < O && b0 + b < B) {
        float v = s * inv + (bias != nullptr ? bias[o] : 0.f);
        if (y_add != nullptr) v += y_add[(size_t)(b0 + b) * O + o];
        y[(size_t)(b0 + b) * O + o] = v > 0.f ? v : v * out_slope;
      }
    }
  }
}

// gx[b][k] = (sum_o gy'[b][o] W[o][k]) * inv_sigma * f'(x[b][k]),  gy' = gy * out_lrelu'(y)
// CTA = 32 k-columns x 8 slices of the output dimension (one warp per slice), reduced through shared memory.
__global__ void linear_bwd_x_kernel(const float* __restrict__ gy, const float* __restrict__ y, float out_slope,
                                    const float* __restrict__ w, const float* __restrict__ sigma,
                                    const float* __restrict__ x, float in_slope, float* __restrict__ gx, int accumulate,
                                    int B, int K, int O) {
  __shared__ float gs[LB][256];
  __shared__ float part[8][LB][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + tx;
  const int b0 = blockIdx.y * LB;
  float acc[LB];
#pragma unroll
  for (int b = 0; b < LB; ++b) acc[b] = 0.f;
  for (int o0 = 0; o0 < O; o0 += 256) {
    __syncthreads();
    for (int i = threadIdx.x; i < LB * 256; i += blockDim.x) {
      const int b = i / 256, o = o0 + i % 256;
      float v = 0.f;
      if (b0 + b < B && o < O) {
        v = gy[(size_t)(b0 + b) * O + o];
        if (y != nullptr && !(y[(size_t)(b0 + b) * O + o] > 0.f)) v *= out_slope;
      }
      gs[b][i % 256] = v;
    }
    __syncthreads();
    if (k < K) {
      const int no = min(256, O - o0);
#pragma unroll 4
      for (int oo = ty; oo < no; oo += 8) {
        const float wv = __ldg(&w[(size_t)(o0 + oo) * K + k]);
#pragma unroll
        for (int b = 0; b < LB; ++b) acc[b] += wv * gs[b][oo];
      }
    }
  }
#pragma unroll
  for (int b = 0; b < LB; ++b) part[ty][b][tx] = acc[b];
  __syncthreads();
  if (ty == 0 && k < K) {
    const float inv = sigma != nullptr ? 1.f / __ldg(sigma) : 1.f;
#pragma unroll
    for (int b = 0; b < LB; ++b) {
      if (b0 + b >= B) break;
      float v = 0.f;
#pragma unroll
      for (int s = 0; s < 8; ++s) v += part[s][b][tx];
      v *= inv;
      const size_t i = (size_t)(b0 + b) * K + k;
      if (x != nullptr && !(x[i] > 0.f)) v *= in_slope;
      gx[i] = accumulate ? gx[i] + v : v;
    }
  }
}

// gw[o][k] += sum_b gy'[b][o] f(x[b][k]);  gb[o] += sum_b gy'[b][o]      (gradient w.r.t. W / sigma)
// A thread owns one input column k: its B activations stay in registers while it walks over the WROWS output rows of
// the CTA, whose gy' values sit in shared memory (broadcast reads).  One global load per B outputs instead of two per
// FMA (the per-(o, k) version spent 100 us on the 4096 -> 2048 layer, 20x its store time).
constexpr int WROWS = 16;
template <int BMAX>  // batch rows held in registers (multiple of 4, >= B)
__global__ void __launch_bounds__(128) linear_bwd_w_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                                                           float out_slope, const float* __restrict__ x,
                                                           const float* __restrict__ xmask, float in_slope,
                                                           float* __restrict__ gw, float* __restrict__ gb, int B, int K,
                                                           int O) {
  __shared__ __align__(16) float gs[WROWS][BMAX];
  const int o0 = blockIdx.y * WROWS;
  for (int i = threadIdx.x; i < WROWS * BMAX; i += blockDim.x) {
    const int r = i / BMAX, b = i % BMAX, o = o0 + r;
    float v = 0.f;
    if (b < B && o < O) {
      v = gy[(size_t)b * O + o];
      if (y != nullptr && !(y[(size_t)b * O + o] > 0.f)) v *= out_slope;
    }
    gs[r][b] = v;
  }
  __syncthreads();
  if (blockIdx.x == 0 && gb != nullptr && threadIdx.x < WROWS && o0 + threadIdx.x < O) {
    float sum = 0.f;
    for (int b = 0; b < B; ++b) sum += gs[threadIdx.x][b];
    gb[o0 + threadIdx.x] += sum;  // one thread of one CTA owns this output: plain read-modify-write
  }
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float xk[BMAX];
#pragma unroll
  for (int b = 0; b < BMAX; ++b)
    xk[b] = b < B ? in_transform(x[(size_t)b * K + k], xmask, (size_t)b * K + k, in_slope) : 0.f;
  const int rows = min(WROWS, O - o0);
  for (int r = 0; r < rows; ++r) {
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < BMAX / 4; ++q) {
      const float4 g4 = *reinterpret_cast<const float4*>(&gs[r][4 * q]);  // broadcast read
      acc += g4.x * xk[4 * q] + g4.y * xk[4 * q + 1] + g4.z * xk[4 * q + 2] + g4.w * xk[4 * q + 3];
    }
    gw[(size_t)(o0 + r) * K + k] += acc;
  }
}

// out[i][j][k] = cls[j] + feat[j][k] * emb_w[idx[i]][k] / sigma          (models.py:151-155)
__global__ void dhead_out_fwd_kernel(const float* __restrict__ cls, const float* __restrict__ feat,
                                     const float* __restrict__ emb_w, const float* __restrict__ sigma,
                                     const int* __restrict__ idx, float* __restrict__ out, int B, int E) {
  const int n = B * B * E;
  const float inv = 1.f / __ldg(sigma);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int k = t % E, j = (t / E) % B, i = t / (E * B);
    out[t] = cls[j] + feat[(size_t)j * E + k] * emb_w[(size_t)idx[i] * E + k] * inv;
  }
}
// g_cls[j] = sum_{i,k} g[i,j,k];  g_feat[j,k] = sum_i g[i,j,k] emb[i,k];  g_embw[idx[i]][k] += sum_j g[i,j,k] feat[j,k]
// one CTA per batch row r (it plays j for g_cls / g_feat and i for g_embw), threads over k
__global__ void dhead_out_bwd_kernel(const float* __restrict__ g, const float* __restrict__ feat,
                                     const float* __restrict__ emb_w, const float* __restrict__ sigma,
                                     const int* __restrict__ idx, float* __restrict__ g_cls, float* __restrict__ g_feat,
                                     float* __restrict__ g_embw, int B, int E) {
  __shared__ float red[32];
  const float inv = 1.f / __ldg(sigma);
  const int r = blockIdx.x;
  float cls_acc = 0.f;
  for (int k = threadIdx.x; k < E; k += blockDim.x) {
    float gf = 0.f;
#pragma unroll 4  // read-only: four samples' loads (g, idx -> embedding row) in flight, sums stay in batch order
    for (int q = 0; q < B; ++q) {
      const float gj = g[((size_t)q * B + r) * E + k];  // i = q, j = r
      gf += gj * emb_w[(size_t)idx[q] * E + k] * inv;
      cls_acc += gj;
    }
    g_feat[(size_t)r * E + k] = gf;
  }
  if (g_embw != nullptr) {
    // samples of the same class share an embedding row: the CTA of the FIRST such sample adds all of them, in batch
    // order (no atomics, bit-reproducible)
    const int cls_r = idx[r];
    bool first = true;
    for (int q = 0; q < r; ++q) first = first && (idx[q] != cls_r);
    if (first) {
      for (int k = threadIdx.x; k < E; k += blockDim.x) {
        float ge = 0.f;
        for (int r2 = r; r2 < B; ++r2) {
          if (idx[r2] != cls_r) continue;
#pragma unroll 4
          for (int q = 0; q < B; ++q) ge += g[((size_t)r2 * B + q) * E + k] * feat[(size_t)q * E + k];  // i = r2, j = q
        }
        g_embw[(size_t)cls_r * E + k] += ge;
      }
    }
  }
  cls_acc = warp_sum(cls_acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cls_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    g_cls[r] = t;
  }
}

}  // namespace

extern "C" int spyr_linear_fwd(const float* x, const float* xmask, float in_slope, const float* w, const float* sigma,
                               const float* bias, const float* y_add, float out_slope, float* y, int B, int K, int O,
                               void* stream) {
  SPYR_REQUIRE(x && w && y && B > 0 && K > 0 && O > 0, "linear_fwd: bad arguments");
  dim3 grid(ceil_div(O, LROWS), ceil_div(B, LB));
  linear_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, xmask, in_slope, w, sigma, bias, y_add, out_slope, y, B, K, O);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_linear_bwd_x(const float* gy, const float* y, float out_slope, const float* w, const float* sigma,
                                 const float* x, float in_slope, float* gx, int accumulate, int B, int K, int O,
                                 void* stream) {
  SPYR_REQUIRE(gy && w && gx, "linear_bwd_x: bad arguments");
  dim3 grid(ceil_div(K, 32), ceil_div(B, LB));
  linear_bwd_x_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gy, y, out_slope, w, sigma, x, in_slope, gx, accumulate, B, K,
                                                              O);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_linear_bwd_w(const float* gy, const float* y, float out_slope, const float* x, const float* xmask,
                                 float in_slope, float* gw, float* gb, int B, int K, int O, void* stream) {
  SPYR_REQUIRE(gy && x && gw && B <= 32, "linear_bwd_w: bad arguments (batch must be <= 32)");
  dim3 grid(ceil_div(K, 128), ceil_div(O, WROWS));
  if (B <= 8)
    linear_bwd_w_kernel<8><<<grid, 128, 0, (cudaStream_t)stream>>>(gy, y, out_slope, x, xmask, in_slope, gw, gb, B, K, O);
  else if (B <= 20)
    linear_bwd_w_kernel<20><<<grid, 128, 0, (cudaStream_t)stream>>>(gy, y, out_slope, x, xmask, in_slope, gw, gb, B, K, O);
  else
    linear_bwd_w_kernel<32><<<grid, 128, 0, (cudaStream_t)stream>>>(gy, y, out_slope, x, xmask, in_slope, gw, gb, B, K, O);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_dhead_out_fwd(const float* cls, const float* feat, const float* emb_w, const float* sigma,
                                  const int* idx, float* out, int B, int E, void* stream) {
  const int n = B * B * E;
  dhead_out_fwd_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(cls, feat, emb_w, sigma, idx, out, B, E);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_dhead_out_bwd(const float* g, const float* feat, const float* emb_w, const float* sigma, const int* idx,
                                  float* g_cls, float* g_feat, float* g_embw, int B, int E, void* stream) {
  dhead_out_bwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(g, feat, emb_w, sigma, idx, g_cls, g_feat, g_embw, B, E);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
