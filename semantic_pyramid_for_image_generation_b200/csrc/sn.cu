// Spectral normalisation for ALL layers of a model in three launches per forward (plus two per backward).
//
// Replaces torch.nn.utils.spectral_norm (torch/nn/utils/spectral_norm.py:92-114) at every call site in the
// reference (models.py:28,34,55,58,128,132,135,232-243,299-313,356-360,393-403,438-448):
//   train : v <- normalize(W^T u), u <- normalize(W v)   (one power iteration, in place, eps 1e-12)
//           sigma = u . (W v);  the layer then uses W / sigma
//   eval  : sigma from the stored u, v (no iteration)
// and emits, in the same pass, the BF16 operand the tensor-core convolutions consume:
//   packed[tap][cout][cin] = bf16(W[cout][cin][tap] / sigma)      (see spyr_conv_src.w)
// plus, for the `cat(feature*mask, mask)` convolutions (models.py:94,312-315), the FP32 taps of the extra mask
// channel as stencil[10][cout] (9 taps + their sum).
// Backward: dL/dW = (G - <G, W/sigma> u v^T) / sigma with u, v treated as constants (they are detached clones in
// torch), G given either in W's own layout or in the wgrad kernel's [tap][cin][cout] layout.
#include "common.cuh"
#include "../../include/spyramid_b200.h"

extern void spyr_count_launch();

namespace {

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

__device__ __forceinline__ int find_layer(const spyr_sn_layer* tab, int n, int tile, int which, int* sh_idx) {
  if (threadIdx.x == 0) {
    int l = 0;
    for (int i = 0; i < n; ++i) {
      const int t0 = which == 0 ? tab[i].tile0_wtu
                     : (which == 1 ? tab[i].tile0_wv
                                   : (which == 2 ? tab[i].tile0_pack : (which == 3 ? tab[i].tile0_bwd : tab[i].tile0_tsum)));
      if (t0 <= tile) l = i;
    }
    *sh_idx = l;
  }
  __syncthreads();
  return *sh_idx;
}

constexpr int WTU_ROWS = 64, WTU_COLS = 256;

// partial[rt][j] = sum_{i in row tile rt} W[i][j] u[i]: one CTA per (row tile, 256 columns), plain stores into the layer's
// partial slices; sn_tsum_kernel adds the row tiles in order (no atomics: bit-reproducible)
__global__ void sn_wtu_kernel(const spyr_sn_layer* __restrict__ tab, int n, float* __restrict__ scratch) {
  __shared__ int sh_idx;
  __shared__ float su[WTU_ROWS];
  const int l = find_layer(tab, n, blockIdx.x, 0, &sh_idx);
  const spyr_sn_layer L = tab[l];
  const int tile = blockIdx.x - L.tile0_wtu;
  const int ctiles = (L.cols + WTU_COLS - 1) / WTU_COLS;
  const int rt = tile / ctiles;
  const int r0 = rt * WTU_ROWS, c0 = (tile % ctiles) * WTU_COLS;
  const int nr = min(WTU_ROWS, L.rows - r0);
  if (threadIdx.x < nr) su[threadIdx.x] = L.u[r0 + threadIdx.x];
  __syncthreads();
  const int j = c0 + threadIdx.x;
  if (j >= L.cols) return;
  const float* wp = L.w + (size_t)r0 * L.cols + j;
  float acc = 0.f;
#pragma unroll 8
  for (int i = 0; i < nr; ++i) acc += wp[(size_t)i * L.cols] * su[i];
  scratch[L.part_off + (size_t)rt * L.cols + j] = acc;
}
// t[j] = sum over the row tiles (in order) of partial[rt][j]
__global__ void sn_tsum_kernel(const spyr_sn_layer* __restrict__ tab, int n, float* __restrict__ scratch) {
  __shared__ int sh_idx;
  const int l = find_layer(tab, n, blockIdx.x, 4, &sh_idx);
  const spyr_sn_layer L = tab[l];
  const int j = (blockIdx.x - L.tile0_tsum) * blockDim.x + threadIdx.x;
  if (j >= L.cols) return;
  const int rtiles = (L.rows + WTU_ROWS - 1) / WTU_ROWS;
  float acc = 0.f;
  for (int rt = 0; rt < rtiles; ++rt) acc += scratch[L.part_off + (size_t)rt * L.cols + j];
  scratch[L.scratch_off + j] = acc;
}

constexpr int WV_ROWS = 8;  // one weight row per warp
constexpr int PACK_ROWS = 1;  // weight rows packed per CTA (4 measured slower: 70 vs 52 us, fewer CTAs in flight)

// s[i] = sum_j W[i][j] v[j], v = t / max(|t|, eps) (train) or the stored v (eval); one CTA also stores v
__global__ void sn_wv_kernel(const spyr_sn_layer* __restrict__ tab, int n, float* __restrict__ scratch, int training,
                             float eps) {
  __shared__ int sh_idx;
  __shared__ float red[32];
  const int l = find_layer(tab, n, blockIdx.x, 1, &sh_idx);
  const spyr_sn_layer L = tab[l];
  const int tile = blockIdx.x - L.tile0_wv;
  const float* vec;
  float vscale = 1.f;
  if (training) {
    const float* t = scratch + L.scratch_off;
    float ss = 0.f;
    for (int j = threadIdx.x; j < L.cols; j += blockDim.x) ss += t[j] * t[j];
    ss = block_sum(ss, red);
    vscale = 1.f / fmaxf(sqrtf(ss), eps);
    vec = t;
    if (tile == 0)
      for (int j = threadIdx.x; j < L.cols; j += blockDim.x) L.v[j] = t[j] * vscale;
  } else {
    vec = L.v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = tile * WV_ROWS + warp;
  if (i >= L.rows) return;
  const float* wp = L.w + (size_t)i * L.cols;
  float acc = 0.f;
  if ((L.cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(wp) | reinterpret_cast<uintptr_t>(vec)) & 15) == 0) {
    const float4* w4 = reinterpret_cast<const float4*>(wp);
    const float4* v4 = reinterpret_cast<const float4*>(vec);
    const int n4 = L.cols >> 2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int j = lane; j < n4; j += 32) {
      const float4 a = __ldg(w4 + j), b = v4[j];
      a0 += a.x * b.x;
      a1 += a.y * b.y;
      a2 += a.z * b.z;
      a3 += a.w * b.w;
    }
    acc = (a0 + a1) + (a2 + a3);
  } else {
#pragma unroll 4
    for (int j = lane; j < L.cols; j += 32) acc += wp[j] * vec[j];
  }
  acc = warp_sum(acc) * vscale;
  if (lane == 0) scratch[L.scratch_off + L.cols + i] = acc;
}

// sigma, u update, saved copies, BF16 pack.  One CTA per weight row (pack layers) or one CTA per layer.
__global__ void sn_pack_kernel(const spyr_sn_layer* __restrict__ tab, int n, const float* __restrict__ scratch,
                               int training, float eps, bf16* __restrict__ packed, float* __restrict__ stencil,
                               float* __restrict__ saved, int split) {
  __shared__ int sh_idx;
  __shared__ float red[32];
  const int l = find_layer(tab, n, blockIdx.x, 2, &sh_idx);
  const spyr_sn_layer L = tab[l];
  const int tile = blockIdx.x - L.tile0_pack;
  const float* s = scratch + L.scratch_off + L.cols;
  float sigma, uscale = 1.f;
  if (training) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < L.rows; i += blockDim.x) ss += s[i] * s[i];
    ss = block_sum(ss, red);
    uscale = 1.f / fmaxf(sqrtf(ss), eps);
    sigma = ss * uscale;  // u . s with u = s * uscale
  } else {
    float d = 0.f;
    for (int i = threadIdx.x; i < L.rows; i += blockDim.x) d += L.u[i] * s[i];
    sigma = block_sum(d, red);
  }
  float* sv = saved + L.saved_off;
  if (tile == 0) {
    // after the block_sum barrier every thread of this CTA has finished reading L.u
    for (int i = threadIdx.x; i < L.rows; i += blockDim.x) {
      const float un = training ? s[i] * uscale : L.u[i];
      if (training) L.u[i] = un;
      sv[1 + i] = un;
    }
    for (int j = threadIdx.x; j < L.cols; j += blockDim.x) sv[1 + L.rows + j] = L.v[j];
    if (threadIdx.x == 0) sv[0] = sigma;
  }
  if (L.pack_cin <= 0) return;
  const float inv = 1.f / sigma;
  bf16* dst = packed + L.pack_off;
  const int taps = L.taps, pc = L.pack_cin;
  // split mode: the lo plane (bf16 of the rounding residual) follows the hi plane of the layer's pack
  const size_t lo = split ? (size_t)L.rows * pc * (L.pack_mode == 1 ? 1 : taps) : 0;
  for (int co = tile * PACK_ROWS; co < min(L.rows, (tile + 1) * PACK_ROWS); ++co) {
    const float* wrow = L.w + (size_t)co * L.cols;
    if (L.pack_mode == 1) {
      // im2col rows: packed[co][k], k = t*cin + ci, zero padded to pack_cin columns
      for (int k = threadIdx.x; k < pc; k += blockDim.x) {
        float v = 0.f;
        if (k < L.cols) v = wrow[(k % L.cin) * taps + k / L.cin] * inv;
        const bf16 h = __float2bfloat16(v);
        dst[(size_t)co * pc + k] = h;
        if (lo) dst[lo + (size_t)co * pc + k] = __float2bfloat16(v - __bfloat162float(h));
      }
      continue;
    }
    // packed[t][co][ci], ci < pack_cin
#pragma unroll 4
    for (int i = threadIdx.x; i < taps * pc; i += blockDim.x) {
      const int t = i / pc, ci = i % pc;
      const float v = __ldg(wrow + ci * taps + t) * inv;
      const bf16 h = __float2bfloat16(v);
      dst[((size_t)t * L.rows + co) * pc + ci] = h;
      if (lo) dst[lo + ((size_t)t * L.rows + co) * pc + ci] = __float2bfloat16(v - __bfloat162float(h));
    }
    if (L.stencil_off >= 0 && threadIdx.x < 32) {
      // extra (mask) input channel pack_cin: FP32 taps + their sum
      float* st = stencil + L.stencil_off;
      float v = 0.f;
      if (threadIdx.x < taps) {
        v = wrow[pc * taps + threadIdx.x] * inv;
        st[(size_t)threadIdx.x * L.rows + co] = v;
      }
      v = warp_sum(v);
      if (threadIdx.x == 0) st[(size_t)9 * L.rows + co] = v;
    }
  }
}

constexpr int BT = 32;  // backward tile: 32 rows (cout) x BTC input channels x taps
// single-tap layers (linear, 1x1, embedding) take 9x wider tiles so that every CTA moves the same ~9K elements
__host__ __device__ inline int sn_btc(int taps) { return taps == 1 ? 9 * BT : BT; }

// pass 1: dots[tile] = sum over the tile of G * W (one slot per CTA)
// pass 2: out = (G - <G, W>/sigma * u v^T) / sigma, <G, W> = the layer's tile slots summed in a fixed order
// TAPS is a compile-time copy of L.taps for the two shapes that carry all the bytes (1 and 9): the index split
// e -> (ci, t) is then a multiply-shift instead of an integer division per element.
template <int PASS, int TAPS>
__device__ __forceinline__ void sn_bwd_tile(const spyr_sn_layer& L, int tile, const float* __restrict__ gw_arena,
                                            const float* __restrict__ saved, float* __restrict__ dots,
                                            float* __restrict__ grad_arena, float* gsh, float* red) {
  const int taps = TAPS > 0 ? TAPS : L.taps;
  const int btc = sn_btc(taps);
  // tap stride == 4 (mod 32 banks): the 32 lanes of a row sweep (ci, t) pairs, and with a stride == 0 the nine taps
  // of one ci would share a bank
  const int tstride = btc * (BT + 1) + 4;
  const int ctiles = (L.cin + btc - 1) / btc;
  const int co0 = (tile / ctiles) * BT, ci0 = (tile % ctiles) * btc;
  const int nco = min(BT, L.rows - co0), nci = min(btc, L.cin - ci0);
  const float* gw = gw_arena + L.gw_off;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (L.gw_layout == 1) {
    // coalesced along cout.  Each warp owns rows r = wid, wid + nw, ...; four independent 128-byte loads in flight per
    // warp (a rolled loop has one, which made this kernel latency-bound: 36 dependent round trips per warp).
    const int nrow = taps * nci;
    for (int r0 = wid; r0 < nrow; r0 += 4 * nw) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = r0 + q * nw;
        const int t = r / nci, ci = r - t * nci;
        v[q] = (r < nrow && lane < nco) ? __ldg(gw + ((size_t)t * L.cin + ci0 + ci) * L.rows + co0 + lane) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = r0 + q * nw;
        const int t = r / nci, ci = r - t * nci;
        if (r < nrow && lane < nco) gsh[t * tstride + ci * (BT + 1) + lane] = v[q];
      }
    }
  } else {
    for (int r = wid; r < nco; r += nw)
      for (int e = lane; e < nci * taps; e += 32) {
        const int ci = e / taps, t = e - ci * taps;
        gsh[t * tstride + ci * (BT + 1) + r] = gw[(size_t)(co0 + r) * L.cols + (size_t)(ci0 + ci) * taps + t];
      }
  }
  __syncthreads();
  const float* sv = saved + L.saved_off;
  const float sigma = sv[0];
  if (PASS == 1) {
    float acc = 0.f;
    for (int r = wid; r < nco; r += nw) {
      const float* wrow = L.w + (size_t)(co0 + r) * L.cols + (size_t)ci0 * taps;
#pragma unroll 3
      for (int e = lane; e < nci * taps; e += 32) {
        const int ci = e / taps, t = e - ci * taps;
        acc += gsh[t * tstride + ci * (BT + 1) + r] * __ldg(wrow + e);
      }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) dots[L.tile0_bwd + tile] = acc;
  } else {
    const float inv = 1.f / sigma;
    const int ltiles = ((L.rows + BT - 1) / BT) * ctiles;
    float dsum = 0.f;
    for (int i = threadIdx.x; i < ltiles; i += blockDim.x) dsum += dots[L.tile0_bwd + i];
    const float coef = block_sum(dsum, red) * inv;  // <G, W/sigma>
    float* out = grad_arena + L.grad_off;
    const float* vv = sv + 1 + L.rows + (size_t)ci0 * taps;
    for (int r = wid; r < nco; r += nw) {
      const float ur = sv[1 + co0 + r] * coef;
      float* orow = out + (size_t)(co0 + r) * L.cols + (size_t)ci0 * taps;
#pragma unroll 3
      for (int e = lane; e < nci * taps; e += 32) {
        const int ci = e / taps, t = e - ci * taps;
        orow[e] = (gsh[t * tstride + ci * (BT + 1) + r] - ur * __ldg(vv + e)) * inv;
      }
    }
  }
}

template <int PASS>
__global__ void sn_bwd_kernel(const spyr_sn_layer* __restrict__ tab, int n, const float* __restrict__ gw_arena,
                              const float* __restrict__ saved, float* __restrict__ dots, float* __restrict__ grad_arena) {
  extern __shared__ float gsh[];  // [taps][BT ci][BT+1 co]
  __shared__ int sh_idx;
  __shared__ float red[32];
  const int l = find_layer(tab, n, blockIdx.x, 3, &sh_idx);
  const spyr_sn_layer L = tab[l];
  if (L.gw_off < 0) return;
  const int tile = blockIdx.x - L.tile0_bwd;
  if (L.taps == 9)
    sn_bwd_tile<PASS, 9>(L, tile, gw_arena, saved, dots, grad_arena, gsh, red);
  else if (L.taps == 1)
    sn_bwd_tile<PASS, 1>(L, tile, gw_arena, saved, dots, grad_arena, gsh, red);
  else
    sn_bwd_tile<PASS, 0>(L, tile, gw_arena, saved, dots, grad_arena, gsh, red);
}

}  // namespace

extern "C" int spyr_sn_plan(spyr_sn_layer* tab, int n, spyr_sn_plan_out* out) {
  SPYR_REQUIRE(tab != nullptr && out != nullptr && n > 0, "sn_plan: bad arguments");
  int t_wtu = 0, t_wv = 0, t_pack = 0, t_bwd = 0, t_tsum = 0;
  long long scratch = 0, saved = 0;
  for (int i = 0; i < n; ++i) {
    spyr_sn_layer& L = tab[i];
    SPYR_REQUIRE(L.rows > 0 && L.cols > 0 && L.taps > 0 && L.cin * L.taps == L.cols, "sn_plan: layer %d has bad shape", i);
    SPYR_REQUIRE(L.taps <= 9, "sn_plan: layer %d has %d taps", i, L.taps);
    L.index = i;
    L.tile0_wtu = t_wtu;
    t_wtu += ceil_div(L.rows, WTU_ROWS) * ceil_div(L.cols, WTU_COLS);
    L.tile0_tsum = t_tsum;
    t_tsum += ceil_div(L.cols, 256);
    L.tile0_wv = t_wv;
    t_wv += ceil_div(L.rows, WV_ROWS);
    L.tile0_pack = t_pack;
    t_pack += L.pack_cin > 0 ? ceil_div(L.rows, PACK_ROWS) : 1;
    L.tile0_bwd = t_bwd;
    t_bwd += ceil_div(L.rows, BT) * ceil_div(L.cin, sn_btc(L.taps));
    L.scratch_off = scratch;  // 16-byte aligned so the power-iteration vector can be read as float4
    scratch += (L.cols + L.rows + 3) & ~3;
    L.part_off = scratch;  // row-tile partial sums of W^T u
    scratch += ((long long)ceil_div(L.rows, WTU_ROWS) * L.cols + 3) & ~3;
    L.saved_off = saved;
    saved += 1 + L.rows + L.cols;
  }
  out->tiles_wtu = t_wtu;
  out->tiles_tsum = t_tsum;
  out->tiles_wv = t_wv;
  out->tiles_pack = t_pack;
  out->tiles_bwd = t_bwd;
  out->scratch_floats = scratch;
  out->saved_floats = saved;
  return 0;
}

extern "C" int spyr_sn_forward(const spyr_sn_layer* dev_tab, int n, const spyr_sn_plan_out* plan, int training, float eps,
                               float* scratch, void* packed, float* stencil, float* saved, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPYR_REQUIRE(dev_tab && plan && scratch && saved && n > 0, "sn_forward: bad arguments");
  if (training) {
    sn_wtu_kernel<<<plan->tiles_wtu, WTU_COLS, 0, stream>>>(dev_tab, n, scratch);
    spyr_count_launch();
    SPYR_LAUNCH_CHECK();
    sn_tsum_kernel<<<plan->tiles_tsum, 256, 0, stream>>>(dev_tab, n, scratch);
    spyr_count_launch();
    SPYR_LAUNCH_CHECK();
  }
  sn_wv_kernel<<<plan->tiles_wv, 256, 0, stream>>>(dev_tab, n, scratch, training, eps);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  sn_pack_kernel<<<plan->tiles_pack, 128, 0, stream>>>(dev_tab, n, scratch, training, eps, (bf16*)packed, stencil, saved,
                                                       spyr_split() ? 1 : 0);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}

extern "C" int spyr_sn_backward(const spyr_sn_layer* dev_tab, int n, const spyr_sn_plan_out* plan, const float* gw_arena,
                                const float* saved, float* dots, float* grad_arena, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPYR_REQUIRE(dev_tab && plan && gw_arena && saved && dots && grad_arena, "sn_backward: bad arguments");
  const size_t smem = (size_t)9 * (BT * (BT + 1) + 4) * sizeof(float);
  sn_bwd_kernel<1><<<plan->tiles_bwd, 256, smem, stream>>>(dev_tab, n, gw_arena, saved, dots, grad_arena);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  sn_bwd_kernel<2><<<plan->tiles_bwd, 256, smem, stream>>>(dev_tab, n, gw_arena, saved, dots, grad_arena);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
