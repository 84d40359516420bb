// Row softmax of the SAGAN attention map, the loss reductions and the fused Adam update.
//
// Replaces: `.softmax(dim=-1)` between the two bmm of SelfAttention (models.py:266); LSGAN losses
// (lossfunction.py:131-137,156-164); SemanticReconstructionLoss (lossfunction.py:31-68); DiversityLoss
// (lossfunction.py:92-110); torch.optim.Adam.step (main.py:64-65, model_wrapper.py:162,190).
#include "common.cuh"
#include "../../include/spyramid_b200.h"

extern void spyr_count_launch();

namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// ---- attention softmax: one warp per query row ----
template <bool S>
__global__ void softmax_rows_fwd_kernel(const float* __restrict__ s, const ActT<S> p, long long rows, int n) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* sr = s + row * n;
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, sr[j]);
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) sum += __expf(sr[j] - m);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int j = lane; j < n; j += 32) stf(p, (size_t)(row * n + j), __expf(sr[j] - m) * inv);
}
// dS = P * (dP - sum_k dP*P)
template <bool S>
__global__ void softmax_rows_bwd_kernel(const ActT<S> p, const float* __restrict__ dp, const ActT<S> ds,
                                        long long rows, int n) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float d = 0.f;
  for (int j = lane; j < n; j += 32) d += ldf(p, (size_t)(row * n + j)) * dp[row * n + j];
  d = warp_sum(d);
  for (int j = lane; j < n; j += 32) {
    const float pv = ldf(p, (size_t)(row * n + j));
    stf(ds, (size_t)(row * n + j), pv * (dp[row * n + j] - d));
  }
}

// ---- LSGAN: out = 0.5 * mean((p - target)^2); single CTA => deterministic ----
__global__ void lsgan_fwd_kernel(const float* __restrict__ p, long long n, float target, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = p[i] - target;
    acc += d * d;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) *out = 0.5f * acc / (float)n;
}
__global__ void lsgan_bwd_kernel(const float* __restrict__ p, long long n, float target, const float* __restrict__ gout,
                                 float* __restrict__ gp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gp[i] = __ldg(gout) * (p[i] - target) / (float)n;
}

// ---- semantic reconstruction, one pyramid level: loss += mean(|maxpool2(fr) - maxpool2(ff)| * maxpool2(mask)) ----
template <bool S>
__device__ __forceinline__ void pool4(const ActT<S>& x, size_t off, size_t C, size_t WC, float v[4][8]) {
  ld8(x + off, v[0]);
  ld8(x + off + C, v[1]);
  ld8(x + off + WC, v[2]);
  ld8(x + off + WC + C, v[3]);
}
template <bool S>
__global__ void rec_level_fwd_kernel(const ActT<S> fr, const ActT<S> ff,
                                     const float* __restrict__ mask, int B, int H, int W, int cg, float inv_numel,
                                     float* __restrict__ loss, float* __restrict__ scratch, unsigned int* ticket) {
  __shared__ float red[32];
  const int OH = H / 2, OW = W / 2;
  const size_t C = (size_t)cg * 8;
  const long long n = (long long)B * OH * OW * cg;
  float acc = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned)idx % (unsigned)cg);  // 32-bit index split: element counts are < 2^31 (checked at launch)
    int t = (int)((unsigned)idx / (unsigned)cg);
    const int w = (int)(t % OW);
    t /= OW;
    const int h = (int)(t % OH);
    const int b = (int)(t / OH);
    const float* mp = mask + ((size_t)b * H + 2 * h) * W + 2 * w;
    const float pm = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[W], mp[W + 1]));
    if (pm == 0.f) continue;
    const size_t off = (((size_t)b * H + 2 * h) * W + 2 * w) * C + (size_t)c * 8;
    float a[4][8], q[4][8];
    pool4(fr, off, C, (size_t)W * C, a);
    pool4(ff, off, C, (size_t)W * C, q);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float pr = fmaxf(fmaxf(a[0][j], a[1][j]), fmaxf(a[2][j], a[3][j]));
      const float pf = fmaxf(fmaxf(q[0][j], q[1][j]), fmaxf(q[2][j], q[3][j]));
      acc += fabsf((pr - pf) * pm);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) scratch[blockIdx.x] = acc;
  if (spyr_last_block(ticket, gridDim.x))
    spyr_sum_partials<float>(scratch, (int)gridDim.x, 1, [&](int, float total) { *loss += total * inv_numel; });
}
// g_ff = d loss / d ff : -sign((pr - pf) * pm) * pm / numel * gout, routed to the first maximum of the 2x2 window
template <bool S>
__global__ void rec_level_bwd_kernel(const ActT<S> fr, const ActT<S> ff,
                                     const float* __restrict__ mask, int B, int H, int W, int cg, float inv_numel,
                                     const float* __restrict__ gout, const ActT<S> gff) {
  const int OH = H / 2, OW = W / 2;
  const size_t C = (size_t)cg * 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * OH * OW * cg) return;
  const int c = (int)((unsigned)idx % (unsigned)cg);  // 32-bit index split: element counts are < 2^31 (checked at launch)
  int t = (int)((unsigned)idx / (unsigned)cg);
  const int w = (int)(t % OW);
  t /= OW;
  const int h = (int)(t % OH);
  const int b = (int)(t / OH);
  const float* mp = mask + ((size_t)b * H + 2 * h) * W + 2 * w;
  const float pm = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[W], mp[W + 1]));
  const size_t off = (((size_t)b * H + 2 * h) * W + 2 * w) * C + (size_t)c * 8;
  const size_t WC = (size_t)W * C;
  float o[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) o[k][j] = 0.f;
  if (pm != 0.f) {
    const float coef = __ldg(gout) * inv_numel * pm;
    float a[4][8], q[4][8];
    pool4(fr, off, C, WC, a);
    pool4(ff, off, C, WC, q);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float pr = fmaxf(fmaxf(a[0][j], a[1][j]), fmaxf(a[2][j], a[3][j]));
      int best = 0;
      float pf = q[0][j];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (q[k][j] > pf) {
          pf = q[k][j];
          best = k;
        }
      const float d = (pr - pf) * pm;
      const float g = d > 0.f ? -coef : (d < 0.f ? coef : 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k == best) o[k][j] = g;
    }
  }
  st8(gff + off, o[0]);
  st8(gff + off + C, o[1]);
  st8(gff + off + WC, o[2]);
  st8(gff + off + WC + C, o[3]);
}
// vector levels (fc7, logits): MaxPool1d(2) over consecutive pairs, FP32 (B, N)
__global__ void rec_vec_fwd_kernel(const float* __restrict__ fr, const float* __restrict__ ff,
                                   const float* __restrict__ mask, int B, int N, float inv_numel, float* __restrict__ loss,
                                   float* __restrict__ scratch, unsigned int* ticket) {
  __shared__ float red[32];
  const int P = N / 2;
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * P; i += gridDim.x * blockDim.x) {
    const int b = i / P, k = i % P;
    const size_t o = (size_t)b * N + 2 * k;
    const float pm = fmaxf(mask[o], mask[o + 1]);
    acc += fabsf((fmaxf(fr[o], fr[o + 1]) - fmaxf(ff[o], ff[o + 1])) * pm);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) scratch[blockIdx.x] = acc;
  if (spyr_last_block(ticket, gridDim.x))
    spyr_sum_partials<float>(scratch, (int)gridDim.x, 1, [&](int, float total) { *loss += total * inv_numel; });
}
__global__ void rec_vec_bwd_kernel(const float* __restrict__ fr, const float* __restrict__ ff,
                                   const float* __restrict__ mask, int B, int N, float inv_numel,
                                   const float* __restrict__ gout, float* __restrict__ gff) {
  const int P = N / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * P) {
    // odd tail element never enters a pooling window
    if ((N & 1) && i < B * P + B) gff[(size_t)(i - B * P) * N + N - 1] = 0.f;
    return;
  }
  const int b = i / P, k = i % P;
  const size_t o = (size_t)b * N + 2 * k;
  const float pm = fmaxf(mask[o], mask[o + 1]);
  const int best = ff[o + 1] > ff[o] ? 1 : 0;
  const float d = (fmaxf(fr[o], fr[o + 1]) - ff[o + best]) * pm;
  const float coef = __ldg(gout) * inv_numel * pm;
  const float g = d > 0.f ? -coef : (d < 0.f ? coef : 0.f);
  gff[o + best] = g;
  gff[o + 1 - best] = 0.f;
}

// ---- diversity: work[0] = mean|z_a - z_b|, work[1] = mean|img_a - img_b|; loss = work[0] / (work[1] + 1e-8) ----
__global__ void absdiff_mean_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float inv_n,
                                    float* __restrict__ out, float* __restrict__ scratch, unsigned int* ticket) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += fabsf(a[i] - b[i]);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) scratch[blockIdx.x] = acc;
  if (spyr_last_block(ticket, gridDim.x))
    spyr_sum_partials<float>(scratch, (int)gridDim.x, 1, [&](int, float total) { *out = total * inv_n; });
}
__global__ void diversity_finalize_kernel(const float* __restrict__ work, float* __restrict__ loss) {
  *loss = work[0] / (work[1] + 1e-08f);
}
__global__ void diversity_bwd_kernel(const float* __restrict__ img, long long half, const float* __restrict__ work,
                                     const float* __restrict__ gout, float* __restrict__ gimg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  const float den = work[1] + 1e-08f;
  const float coef = -__ldg(gout) * work[0] / (den * den) / (float)half;
  const float d = img[i] - img[half + i];
  const float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  gimg[i] = coef * s;
  gimg[half + i] = -coef * s;
}

// ---- Adam ----
__global__ void adam_tick_kernel(int* step) { *step += 1; }
__global__ void adam_kernel(const spyr_adam_chunk args, const int* __restrict__ step, float lr, float beta1, float beta2,
                            float eps) {
  const int t = blockIdx.y;
  const long long n = args.n[t];
  float* __restrict__ p = args.p[t];
  const float* __restrict__ g = args.g[t];
  float* __restrict__ m = args.m[t];
  float* __restrict__ v = args.v[t];
  const float st = (float)(*step);
  const float bc1 = 1.f - powf(beta1, st);
  const float bc2s = sqrtf(1.f - powf(beta2, st));
  const float step_size = lr / bc1;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long long n4 = vec ? n >> 2 : 0;
  for (long long i = tid; i < n4; i += nthreads) {
    const float4 gi = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i], pi = reinterpret_cast<float4*>(p)[i];
#define SPYR_ADAM1(c)                                        \
  mi.c = beta1 * mi.c + (1.f - beta1) * gi.c;                \
  vi.c = beta2 * vi.c + (1.f - beta2) * gi.c * gi.c;         \
  pi.c -= step_size * mi.c / (sqrtf(vi.c) / bc2s + eps);
    SPYR_ADAM1(x) SPYR_ADAM1(y) SPYR_ADAM1(z) SPYR_ADAM1(w)
#undef SPYR_ADAM1
    reinterpret_cast<float4*>(m)[i] = mi;
    reinterpret_cast<float4*>(v)[i] = vi;
    reinterpret_cast<float4*>(p)[i] = pi;
  }
  for (long long i = (n4 << 2) + tid; i < n; i += nthreads) {
    const float gi = g[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) / bc2s + eps);
  }
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + i);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
}

}  // namespace

extern "C" int spyr_softmax_rows_fwd(const float* s, void* p, long long rows, int n, void* stream) {
  SPYR_WITH_SPLIT(softmax_rows_fwd_kernel<kS><<<(int)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(s, make_act(p, rows * n), rows, n));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_softmax_rows_bwd(const void* p, const float* dp, void* ds, long long rows, int n, void* stream) {
  SPYR_WITH_SPLIT(softmax_rows_bwd_kernel<kS><<<(int)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(make_act(p, rows * n), dp, make_act(ds, rows * n), rows, n));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_lsgan_fwd(const float* p, long long n, float target, float* out, void* stream) {
  lsgan_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p, n, target, out);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_lsgan_bwd(const float* p, long long n, float target, const float* gout, float* gp, void* stream) {
  lsgan_bwd_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, n, target, gout, gp);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_rec_level_fwd(const void* fr, const void* ff, const float* mask, int B, int H, int W, int C, float* loss,
                                  void* scratch, void* stream) {
  SPYR_REQUIRE(scratch != nullptr, "rec_level_fwd: scratch is NULL");
  SPYR_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "rec_level_fwd: bad shape");
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
  SPYR_N32(n);
  const float inv = 1.f / ((float)B * (float)C * (float)(H / 2) * (float)(W / 2));
  long long want = (n + 1023) / 1024;
  const int grid = (int)(want < 1 ? 1 : (want > SPYR_REDUCE_BLOCKS ? SPYR_REDUCE_BLOCKS : want));
  SPYR_WITH_SPLIT(rec_level_fwd_kernel<kS><<<grid, 256, 0, (cudaStream_t)stream>>>(make_act(fr, n * 32), make_act(ff, n * 32), mask, B, H, W,
                                                              C / 8, inv, loss, (float*)scratch, spyr_next_ticket()));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_rec_level_bwd(const void* fr, const void* ff, const float* mask, int B, int H, int W, int C,
                                  const float* gout, void* gff, void* stream) {
  SPYR_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "rec_level_bwd: bad shape");
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
  SPYR_N32(n);
  const float inv = 1.f / ((float)B * (float)C * (float)(H / 2) * (float)(W / 2));
  SPYR_WITH_SPLIT(rec_level_bwd_kernel<kS><<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(make_act(fr, n * 32), make_act(ff, n * 32), mask, B,
                                                                                 H, W, C / 8, inv, gout, make_act(gff, n * 32)));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_rec_vec_fwd(const float* fr, const float* ff, const float* mask, int B, int N, float* loss,
                                void* scratch, void* stream) {
  SPYR_REQUIRE(scratch != nullptr, "rec_vec_fwd: scratch is NULL");
  const float inv = 1.f / ((float)B * (float)(N / 2));
  rec_vec_fwd_kernel<<<8, 256, 0, (cudaStream_t)stream>>>(fr, ff, mask, B, N, inv, loss, (float*)scratch, spyr_next_ticket());
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_rec_vec_bwd(const float* fr, const float* ff, const float* mask, int B, int N, const float* gout,
                                float* gff, void* stream) {
  const float inv = 1.f / ((float)B * (float)(N / 2));
  const int n = B * (N / 2) + B;
  rec_vec_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(fr, ff, mask, B, N, inv, gout, gff);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_diversity_fwd(const float* img, long long img_half, const float* z, long long z_half, float* work,
                                  float* loss, void* scratch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPYR_REQUIRE(img_half > 0 && z_half > 0, "diversity_fwd: batch must be > 1 (lossfunction.py:100)");
  SPYR_REQUIRE(scratch != nullptr, "diversity_fwd: scratch is NULL");
  float* sc = (float*)scratch;
  absdiff_mean_kernel<<<4, 256, 0, stream>>>(z, z + z_half, z_half, 1.f / (float)z_half, work, sc, spyr_next_ticket());
  spyr_count_launch();
  long long want = (img_half + 2047) / 2048;
  const int grid = (int)(want < 1 ? 1 : (want > SPYR_REDUCE_BLOCKS ? SPYR_REDUCE_BLOCKS : want));
  absdiff_mean_kernel<<<grid, 256, 0, stream>>>(img, img + img_half, img_half, 1.f / (float)img_half, work + 1,
                                                sc + SPYR_REDUCE_BLOCKS, spyr_next_ticket());
  spyr_count_launch();
  diversity_finalize_kernel<<<1, 1, 0, stream>>>(work, loss);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_diversity_bwd(const float* img, long long img_half, const float* work, const float* gout, float* gimg,
                                  void* stream) {
  diversity_bwd_kernel<<<(int)((img_half + 255) / 256), 256, 0, (cudaStream_t)stream>>>(img, img_half, work, gout, gimg);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_adam_tick(int* step, void* stream) {
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_adam_step(const spyr_adam_chunk* chunk, const int* step, float lr, float beta1, float beta2, float eps,
                              void* stream) {
  SPYR_REQUIRE(chunk && chunk->count > 0 && chunk->count <= SPYR_ADAM_MAX_TENSORS, "adam_step: bad chunk");
  long long maxn = 0;
  for (int i = 0; i < chunk->count; ++i) maxn = chunk->n[i] > maxn ? chunk->n[i] : maxn;
  long long want = (maxn + 2047) / 2048;
  const int gx = (int)(want < 1 ? 1 : (want > 592 ? 592 : want));
  dim3 grid(gx, chunk->count);
  adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*chunk, step, lr, beta1, beta2, eps);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_add_inplace(float* dst, const float* src, long long n, void* stream) {
  SPYR_REQUIRE(dst && src && n % 4 == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0,
               "add_inplace: needs 16-byte aligned buffers and n %% 4 == 0");
  long long want = (n / 4 + 1023) / 1024;
  const int grid = (int)(want < 1 ? 1 : (want > 1184 ? 1184 : want));
  add_inplace_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dst, src, n / 4);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
