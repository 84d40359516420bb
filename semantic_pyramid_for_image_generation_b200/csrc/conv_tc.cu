// Implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a).
//
//   fprop / dgrad : D[pixel][cout] = sum_{src,tap,cin} X_src[pixel+tap][cin] * W_src[tap][cout][cin]
//                   A = activations, K-major (NHWC: channels contiguous), loaded by TMA as a
//                       {64ch, TW, TH, TN} box at tap-shifted coordinates -- the zero padding of the
//                       convolution is the TMA out-of-bounds fill, so there is no im2col buffer.
//                   B = packed weights [tap][cout][cin], K-major, {64ch, BLOCK_N, 1} box.
//                   128B swizzle on both; tcgen05.mma kind::f16 (BF16 in, FP32 accumulate in TMEM).
//   wgrad         : D[(tap,cin)][cout] = sum_pixel X[pixel+tap][cin] * dY[pixel][cout]
//                   both operands MN-major (the reduction runs over pixels, channels are contiguous).
//
// Warp roles (256 threads): warp0 = TMA producer, warp1 = MMA issuer (one elected lane),
// warp2 = TMEM allocator, warps4-7 = epilogue (TMEM -> registers -> fused epilogue -> global).
//
// Replaces the cuDNN/MKL-DNN calls behind nn.Conv2d at reference models.py:34,55-60,232-243,299-315,
// 393-404,438-449 and torchvision VGG features/classifier (models.py:201-211).
#include "common.cuh"
#include <stdlib.h>
#include "../../include/spyramid_b200.h"

extern void spyr_count_launch();
int spyr_conv_halo_launch(const spyr_conv_desc* d, cudaStream_t stream);
int spyr_conv_halo2_launch(const spyr_conv_desc* d, cudaStream_t stream);
int spyr_wgrad_halo_launch(const spyr_wgrad_desc* d, cudaStream_t stream);

namespace {

constexpr int BLOCK_M = 128;  // pixels per CTA tile == TMEM lanes
constexpr int KC = 64;        // channels per k-step: 64 bf16 = 128 B = one swizzle row
constexpr int A_BYTES = BLOCK_M * KC * 2;

struct FpropParams {
  int B, H, W, Cout;
  int TW, TH, TN;
  int tiles_w, tiles_h;
  int nsrc;
  int ktaps[3];    // taps per source (1 or 9)
  int kchunks[3];  // ceil(cin/64) per source
  int ktotal;      // total k-steps
  int wmn[3];      // source uses the MN-major (input-gradient) weight view
  int wpi[3];      // source uses per-image weights
  int b_bytes;     // bytes reserved for the B tile per stage
  int f32_store;
  int block_n;
  int stages;
  int splits;
  uint32_t tmem_cols;
  const float* bias;
  const float* bias2;
  const float* bias3;
  const float* stencil_mask;
  const float* stencil_w;
  const bf16* dmask;
  float dmask_slope;
  const bf16* residual;
  bf16* y_raw;
  bf16* y_act;
  int act;
  float act_slope;
  float* y_f32;
};

struct TmapPack {
  CUtensorMap x[3];
  CUtensorMap w[3];
};

__global__ void __launch_bounds__(256, 1)
conv_fprop_kernel(const __grid_constant__ TmapPack maps, const FpropParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB | B block_n*128] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = A_BYTES + p.b_bytes;
  const int b_chunks = (p.block_n + 63) >> 6;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  int mt = blockIdx.x;
  const int tw_i = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int th_i = mt % p.tiles_h;
  const int tn_i = mt / p.tiles_h;
  const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;
  const int n_off = blockIdx.y * p.block_n;

  // split-K range
  const int per = (p.ktotal + p.splits - 1) / p.splits;
  const int k_begin = blockIdx.z * per;
  const int k_end = min(p.ktotal, k_begin + per);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) {
      tma_prefetch_desc(&maps.x[s]);
      tma_prefetch_desc(&maps.w[s]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (k_begin < k_end) {
    if (warp == 0) {
      if (elect_one()) {
        // ===== TMA producer =====
        // decode k_begin -> (src, tap, chunk)
        int s = 0, rem = k_begin;
        while (s < p.nsrc - 1 && rem >= p.ktaps[s] * p.kchunks[s]) {
          rem -= p.ktaps[s] * p.kchunks[s];
          ++s;
        }
        int tap = rem / p.kchunks[s];
        int chunk = rem % p.kchunks[s];
        int stage = 0;
        uint32_t phase = 0;
        for (int it = k_begin; it < k_end; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * stage_bytes;
          uint8_t* b_dst = a_dst + A_BYTES;
          const uint32_t b_tx = p.wmn[s] ? (uint32_t)(b_chunks * 8192) : (uint32_t)(p.block_n * KC * 2);
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)A_BYTES + b_tx);
          int dy = 0, dx = 0;
          if (p.ktaps[s] == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          tma_load_4d(a_dst, &maps.x[s], &full_bar[stage], chunk * KC, w0 + dx, h0 + dy, n0);
          if (p.wmn[s]) {
            // input-gradient view: B[k = forward cout][n = forward cin], n contiguous, taps flipped
            const int wtap = p.wpi[s] ? n0 : (p.ktaps[s] == 9 ? 8 - tap : 0);
            for (int j = 0; j < b_chunks; ++j)
              tma_load_3d(b_dst + j * 8192, &maps.w[s], &full_bar[stage], n_off + j * 64, chunk * KC, wtap);
          } else {
            tma_load_3d(b_dst, &maps.w[s], &full_bar[stage], chunk * KC, n_off, p.wpi[s] ? n0 : tap);
          }
          if (++chunk == p.kchunks[s]) {
            chunk = 0;
            if (++tap == p.ktaps[s]) {
              tap = 0;
              ++s;
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer =====
      const uint32_t idesc_k = umma_idesc_bf16(BLOCK_M, p.block_n, 0, 0);
      const uint32_t idesc_mn = umma_idesc_bf16(BLOCK_M, p.block_n, 0, 1);
      const bool issue = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0;
      uint32_t phase = 0;
      int s = 0, s_end = p.ktaps[0] * p.kchunks[0];
      for (int it = k_begin; it < k_end; ++it) {
        while (it >= s_end && s < p.nsrc - 1) {
          ++s;
          s_end += p.ktaps[s] * p.kchunks[s];
        }
        const bool mn = p.wmn[s] != 0;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        // warp-uniform operands, elected issue (a single-lane loop costs an R2UR/ELECT waterfall per MMA)
        const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
        const uint32_t b_addr = a_addr + A_BYTES;
        const uint64_t da0 = umma_smem_desc_sw128(a_addr, 0, 1024);
        // K-major B: 16 channels = 32 B along the swizzled row.  MN-major B: 16 k-rows of 128 B = 2 KB,
        // 64-column chunks 8 KB apart (LBO), 8-row swizzle atoms 1 KB apart (SBO).
        const uint64_t db0 = mn ? umma_smem_desc_sw128(b_addr, 8192, 1024) : umma_smem_desc_sw128(b_addr, 0, 1024);
        const uint32_t bk16 = mn ? (2048u >> 4) : 2u;
        const uint32_t idesc = mn ? idesc_mn : idesc_k;
        if (issue) {
#pragma unroll
          for (int k = 0; k < KC / 16; ++k)
            umma_bf16(tmem_u, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(bk16 * k), idesc, (it > k_begin || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (it == k_end - 1) umma_commit(tmem_full_bar);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else if (warp >= 4) {
      // ===== epilogue =====
      const int q = warp & 3;  // TMEM lane quarter this warp may read
      const int m = q * 32 + lane;
      const int tw = m % p.TW;
      const int th = (m / p.TW) % p.TH;
      const int tn = m / (p.TW * p.TH);
      const int n = n0 + tn, h = h0 + th, w = w0 + tw;
      const bool valid = (n < p.B) && (h < p.H) && (w < p.W);
      const size_t pix = ((size_t)n * p.H + h) * p.W + w;

      // 3x3 neighbourhood of the 1-channel stencil mask (zero padded)
      float mk[9];
      int mk_mode = 0;  // 0: all zero, 1: all one, 2: general
      if (p.stencil_mask != nullptr && valid) {
        bool all0 = true, all1 = true;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
          float v = 0.f;
          if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) v = __ldg(&p.stencil_mask[((size_t)n * p.H + hh) * p.W + ww]);
          mk[t] = v;
          all0 = all0 && (v == 0.f);
          all1 = all1 && (v == 1.f);
        }
        mk_mode = all0 ? 0 : (all1 ? 1 : 2);
      }

      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        if (!valid) continue;
        const int col0 = n_off + c0;
        if (col0 >= p.Cout) continue;
        if (p.y_f32 != nullptr) {
          float* dst = p.y_f32 + pix * p.Cout + col0;
          if (p.f32_store) {
            if (col0 + 32 <= p.Cout && (p.Cout & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                  __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.Cout) dst[j] = __uint_as_float(r[j]);
            }
            continue;
          }
          if (col0 + 32 <= p.Cout && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              atomicAdd(reinterpret_cast<float4*>(dst + j),
                        make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                    __uint_as_float(r[j + 3])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.Cout) atomicAdd(dst + j, __uint_as_float(r[j]));
          }
          continue;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // groups of 8 channels = one 16-byte vector
          const int col = col0 + g * 8;
          if (col + 8 > p.Cout) break;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __ldg(&p.bias[col + j]);
          }
          if (p.bias2 != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __ldg(&p.bias2[col + j]);
          }
          if (p.bias3 != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __ldg(&p.bias3[col + j]);
          }
          if (mk_mode == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __ldg(&p.stencil_w[9 * p.Cout + col + j]);
          } else if (mk_mode == 2) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (mk[t] != 0.f) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += mk[t] * __ldg(&p.stencil_w[t * p.Cout + col + j]);
              }
            }
          }
          const size_t off = pix * p.Cout + col;
          if (p.dmask != nullptr) {
            const uint4 mv = *reinterpret_cast<const uint4*>(p.dmask + off);
            const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack_bf16x2(mw[j]);
              if (!(f.x > 0.f)) v[2 * j] *= p.dmask_slope;
              if (!(f.y > 0.f)) v[2 * j + 1] *= p.dmask_slope;
            }
          }
          if (p.residual != nullptr) {
            const uint4 rv = *reinterpret_cast<const uint4*>(p.residual + off);
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack_bf16x2(rw[j]);
              v[2 * j] += f.x;
              v[2 * j + 1] += f.y;
            }
          }
          if (p.y_raw != nullptr) {
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]);
            o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]);
            o.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(p.y_raw + off) = o;
          }
          if (p.y_act != nullptr) {
            const float sl = (p.act == 1) ? 0.f : ((p.act == 2) ? p.act_slope : 1.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * sl;
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]);
            o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]);
            o.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(p.y_act + off) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: D[(tap,cin) 128][cout block_n] over a split of the pixel tiles; fp32 red.add to dw[tap][cin][cout]
// ------------------------------------------------------------------------------------------------
constexpr int KP = 64;  // pixels per k-step

struct WgradParams {
  int B, H, W, Cin, Cout, taps;
  int TW, TH, TN;
  int tiles_w, tiles_h, tiles_n;
  int ptiles;       // total pixel tiles (k-steps over the whole problem)
  int cin_chunks;   // ceil(Cin/64)
  int mchunks;      // taps * cin_chunks (64-row chunks of M)
  int block_n;      // multiple of 64, <= 256
  int stages;
  int splits;
  uint32_t tmem_cols;
  uint32_t lbo, sbo;
  int per_image;    // splits per image when > 0 (dw gets one slice per image)
  int cin_stride;   // row stride of dw in input channels
  float* dw;
};

struct WgradMaps {
  CUtensorMap x;
  CUtensorMap dy;
};

__global__ void __launch_bounds__(256, 1)
conv_wgrad_kernel(const __grid_constant__ WgradMaps maps, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int CH_BYTES = KP * 128;  // one 64-channel x 64-pixel chunk = 8 KB
  const int a_bytes = 2 * CH_BYTES;
  const int b_chunks = p.block_n / 64;
  const int stage_bytes = a_bytes + b_chunks * CH_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mtile = blockIdx.x;           // 128 rows = 2 chunks of (tap, 64 cin)
  const int n_off = blockIdx.y * p.block_n;
  const int per = (p.ptiles + p.splits - 1) / p.splits;
  const int k_begin = blockIdx.z * per;
  const int k_end = min(p.ptiles, k_begin + per);
  float* dw_out = p.dw;
  if (p.per_image > 0) dw_out += (size_t)(blockIdx.z / p.per_image) * p.taps * p.Cin * p.Cout;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.dy);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (k_begin < k_end) {
    if (warp == 0) {
      if (elect_one()) {
        int ctap[2], cc0[2], cdy[2], cdx[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int g = mtile * 2 + j;
          if (g < p.mchunks) {
            ctap[j] = g / p.cin_chunks;
            cc0[j] = (g % p.cin_chunks) * 64;
          } else {
            ctap[j] = 0;
            cc0[j] = p.cin_chunks * 64 + 64;  // fully out of bounds -> TMA zero fill
          }
          cdy[j] = (p.taps == 9) ? ctap[j] / 3 - 1 : 0;
          cdx[j] = (p.taps == 9) ? ctap[j] % 3 - 1 : 0;
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int it = k_begin; it < k_end; ++it) {
          int t = it;
          const int tw_i = t % p.tiles_w;
          t /= p.tiles_w;
          const int th_i = t % p.tiles_h;
          const int tn_i = t / p.tiles_h;
          const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * stage_bytes;
          uint8_t* b_dst = a_dst + a_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d(a_dst + j * CH_BYTES, &maps.x, &full_bar[stage], cc0[j], w0 + cdx[j], h0 + cdy[j], n0);
          for (int j = 0; j < b_chunks; ++j)
            tma_load_4d(b_dst + j * CH_BYTES, &maps.dy, &full_bar[stage], n_off + j * 64, w0, h0, n0);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      // warp-uniform operands, one elected lane issues; descriptors advance by integer adds on the start-address
      // field (16-byte units)
      {
        const bool issue = elect_one();
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t idesc = umma_idesc_bf16(BLOCK_M, p.block_n, 1, 1);
        const uint64_t hi = ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)((p.lbo >> 4) & 0x3FFF) << 16) |
                            ((uint64_t)((p.sbo >> 4) & 0x3FFF) << 32);
        const uint32_t base16 = (smem_u32(smem) & 0x3FFFF) >> 4;
        const uint32_t stage16 = (uint32_t)stage_bytes >> 4, a16 = (uint32_t)a_bytes >> 4;
        int stage = 0;
        uint32_t phase = 0;
        for (int it = k_begin; it < k_end; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da0 = hi | (uint64_t)(base16 + (uint32_t)stage * stage16);
          const uint64_t db0 = da0 + a16;
          if (issue) {
#pragma unroll
            for (int k = 0; k < KP / 16; ++k)  // 16 pixels (K) per MMA = 16 rows of 128 B = 2 KB
              umma_bf16(tmem_u, da0 + (uint64_t)(k * 128), db0 + (uint64_t)(k * 128), idesc, (it > k_begin || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (it == k_end - 1) umma_commit(tmem_full_bar);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        __syncwarp();
      }
    } else if (warp >= 4) {
      const int q = warp & 3;
      const int m = q * 32 + lane;
      const int g = mtile * 2 + (m >> 6);
      const bool valid_chunk = g < p.mchunks;
      const int tap = valid_chunk ? g / p.cin_chunks : 0;
      const int ci = valid_chunk ? (g % p.cin_chunks) * 64 + (m & 63) : 0;
      const bool valid = valid_chunk && ci < p.Cin;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        if (!valid) continue;
        const int col0 = n_off + c0;
        float* dst = dw_out + ((size_t)tap * p.cin_stride + ci) * p.Cout + col0;
        if (col0 + 32 <= p.Cout && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            atomicAdd(reinterpret_cast<float4*>(dst + j),
                      make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                  __uint_as_float(r[j + 3])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.Cout) atomicAdd(dst + j, __uint_as_float(r[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

void pick_pixel_tile(int H, int W, int pixels, int* TW, int* TH, int* TN) {
  int tw = W < 16 ? W : 16;
  int th = pixels / tw;
  if (th > H) th = H;
  int tn = pixels / (tw * th);
  *TW = tw;
  *TH = th;
  *TN = tn;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

extern "C" int spyr_conv2d_fprop(const spyr_conv_desc* d, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPYR_REQUIRE(d != nullptr, "conv2d_fprop: null descriptor");
  SPYR_REQUIRE(d->nsrc >= 1 && d->nsrc <= 3, "conv2d_fprop: nsrc=%d out of range", d->nsrc);
  SPYR_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cout > 0, "conv2d_fprop: bad shape");
  SPYR_REQUIRE(is_pow2(d->H) && is_pow2(d->W) || (d->H == 1 && d->W == 1), "conv2d_fprop: H,W must be powers of two");
  {
    // maps of 16x8 pixels and larger run on the persistent halo-tiled kernel (conv_halo.cu); SPYR_CONV_LEGACY=1 forces
    // the per-tap kernel below (A/B measurements)
    static const bool legacy = getenv("SPYR_CONV_LEGACY") != nullptr;
    static const bool no_pair = getenv("SPYR_CONV_NO_PAIR") != nullptr;
    if (!legacy) {
      if (!no_pair) {
        // >= 128 output channels: CTA pairs (tcgen05 cta_group::2) halve the weight operand fetched per SM
        const int rc2 = spyr_conv_halo2_launch(d, stream);
        if (rc2 >= 0) return rc2;
      }
      const int rc = spyr_conv_halo_launch(d, stream);
      if (rc >= 0) return rc;
    }
  }
  SPYR_REQUIRE(!d->pool && !d->residual_pooled,
               "conv2d_fprop: the pooled epilogue / pooled residual exist on the halo-tiled kernels only (maps >= 16x8, no split-K)");
  FpropParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cout = d->Cout;
  pick_pixel_tile(d->H, d->W, BLOCK_M, &p.TW, &p.TH, &p.TN);
  SPYR_REQUIRE(p.TW * p.TH * p.TN == BLOCK_M && p.TN <= 256, "conv2d_fprop: cannot tile %dx%d", d->H, d->W);
  p.tiles_w = d->W / p.TW;
  p.tiles_h = d->H / p.TH;
  const int tiles_n = ceil_div(d->B, p.TN);
  p.nsrc = d->nsrc;
  int bn = d->block_n;
  if (bn == 0) {
    bn = d->Cout >= 256 ? 256 : ((d->Cout + 15) / 16) * 16;
    if (bn < 32) bn = 32;
  }
  SPYR_REQUIRE(bn >= 16 && bn <= 256 && bn % 16 == 0, "conv2d_fprop: block_n=%d invalid", bn);
  p.block_n = bn;
  const bool f32_out = d->y_f32 != nullptr;
  SPYR_REQUIRE(f32_out || (d->Cout % 8 == 0), "conv2d_fprop: bf16 output needs Cout %% 8 == 0 (Cout=%d)", d->Cout);
  SPYR_REQUIRE(d->splits <= 1 || f32_out, "conv2d_fprop: split-K needs y_f32");
  SPYR_REQUIRE(!d->f32_store || (f32_out && d->splits <= 1), "conv2d_fprop: f32_store needs y_f32 and splits<=1");
  p.splits = d->splits < 1 ? 1 : d->splits;
  p.f32_store = d->f32_store;
  p.b_bytes = bn * KC * 2;

  TmapPack maps;
  p.ktotal = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    const spyr_conv_src& src = d->src[s];
    SPYR_REQUIRE(src.ksize == 1 || src.ksize == 3, "conv2d_fprop: ksize must be 1 or 3");
    SPYR_REQUIRE(src.cin % 8 == 0 && src.cin > 0, "conv2d_fprop: cin=%d must be a multiple of 8", src.cin);
    SPYR_REQUIRE(((uintptr_t)src.x & 15) == 0 && ((uintptr_t)src.w & 15) == 0, "conv2d_fprop: unaligned pointer");
    p.ktaps[s] = src.ksize * src.ksize;
    p.kchunks[s] = ceil_div(src.cin, KC);
    p.ktotal += p.ktaps[s] * p.kchunks[s];
    p.wmn[s] = src.w_mn_major ? 1 : 0;
    p.wpi[s] = src.w_per_image ? 1 : 0;
    SPYR_REQUIRE(!src.w_per_image || (src.ksize == 1 && p.TN == 1),
                 "conv2d_fprop: per-image weights need ksize 1 and H*W >= 128");
    SPYR_REQUIRE(!src.w_mn_major || (d->Cout % 8 == 0), "conv2d_fprop: MN-major weights need Cout %% 8 == 0");
    if (src.w_mn_major && ceil_div(bn, 64) * 8192 > p.b_bytes) p.b_bytes = ceil_div(bn, 64) * 8192;
    {
      uint64_t dims[4] = {(uint64_t)src.cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
      uint64_t strides[3] = {(uint64_t)src.cin * 2, (uint64_t)d->W * src.cin * 2, (uint64_t)d->H * d->W * src.cin * 2};
      uint32_t box[4] = {KC, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
      if (spyr_tmap_encode(&maps.x[s], src.x, 4, dims, strides, box, 1)) return 3;
    }
    const uint64_t wslices = src.w_per_image ? (uint64_t)d->B : (uint64_t)p.ktaps[s];
    if (src.w_mn_major) {
      uint64_t dims[3] = {(uint64_t)d->Cout, (uint64_t)src.cin, wslices};
      uint64_t strides[2] = {(uint64_t)d->Cout * 2, (uint64_t)d->Cout * src.cin * 2};
      uint32_t box[3] = {64, KC, 1};
      if (spyr_tmap_encode(&maps.w[s], src.w, 3, dims, strides, box, 1)) return 3;
    } else {
      uint64_t dims[3] = {(uint64_t)src.cin, (uint64_t)d->Cout, wslices};
      uint64_t strides[2] = {(uint64_t)src.cin * 2, (uint64_t)d->Cout * src.cin * 2};
      uint32_t box[3] = {KC, (uint32_t)bn, 1};
      if (spyr_tmap_encode(&maps.w[s], src.w, 3, dims, strides, box, 1)) return 3;
    }
  }
  for (int s = d->nsrc; s < 3; ++s) {
    maps.x[s] = maps.x[0];
    maps.w[s] = maps.w[0];
  }
  if (p.splits > p.ktotal) p.splits = p.ktotal;
  const int stage_bytes = A_BYTES + p.b_bytes;
  int stages = d->stages;
  if (stages == 0) {
    stages = (bn > 128) ? 4 : (96 * 1024) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
  }
  p.stages = stages;
  p.tmem_cols = pow2_cols(bn);
  p.bias = d->bias;
  p.bias2 = d->bias2;
  p.bias3 = d->bias3;
  p.stencil_mask = d->stencil_mask;
  p.stencil_w = d->stencil_w;
  p.dmask = (const bf16*)d->dmask;
  p.dmask_slope = d->dmask_slope;
  p.residual = (const bf16*)d->residual;
  p.y_raw = (bf16*)d->y_raw;
  p.y_act = (bf16*)d->y_act;
  p.act = d->act;
  p.act_slope = d->act_slope;
  p.y_f32 = d->y_f32;

  const size_t smem_bytes = (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
  SPYR_REQUIRE(smem_bytes <= 227 * 1024, "conv2d_fprop: smem %zu too large", smem_bytes);
  static size_t configured = 0;
  if (smem_bytes > configured) {
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(conv_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  dim3 grid(p.tiles_w * p.tiles_h * tiles_n, ceil_div(d->Cout, bn), p.splits);
  conv_fprop_kernel<<<grid, 256, smem_bytes, stream>>>(maps, p);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}

extern "C" int spyr_conv2d_wgrad(const spyr_wgrad_desc* d, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPYR_REQUIRE(d != nullptr, "conv2d_wgrad: null descriptor");
  SPYR_REQUIRE(d->ksize == 1 || d->ksize == 3, "conv2d_wgrad: ksize must be 1 or 3");
  SPYR_REQUIRE(d->Cin % 8 == 0 && d->Cout % 8 == 0, "conv2d_wgrad: Cin/Cout must be multiples of 8");
  SPYR_REQUIRE((is_pow2(d->H) && is_pow2(d->W)) || (d->H == 1 && d->W == 1), "conv2d_wgrad: H,W must be powers of two");
  {
    // 3x3 gradients on maps of 16x8 pixels and larger run on the halo-tiled kernel (wgrad_halo.cu)
    static const bool legacy = getenv("SPYR_CONV_LEGACY") != nullptr;
    if (!legacy) {
      const int rc = spyr_wgrad_halo_launch(d, stream);
      if (rc >= 0) return rc;
    }
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
  p.taps = d->ksize * d->ksize;
  pick_pixel_tile(d->H, d->W, KP, &p.TW, &p.TH, &p.TN);
  SPYR_REQUIRE(p.TW * p.TH * p.TN == KP, "conv2d_wgrad: cannot tile %dx%d", d->H, d->W);
  p.tiles_w = d->W / p.TW;
  p.tiles_h = d->H / p.TH;
  p.tiles_n = ceil_div(d->B, p.TN);
  p.ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.cin_chunks = ceil_div(d->Cin, 64);
  p.mchunks = p.taps * p.cin_chunks;
  int bn = d->Cout >= 256 ? 256 : ceil_div(d->Cout, 64) * 64;
  p.block_n = bn;
  const int mtiles = ceil_div(p.mchunks, 2);
  const int ntiles = ceil_div(d->Cout, bn);
  int splits = d->splits;
  if (splits <= 0) {
    splits = (2 * 148) / (mtiles * ntiles);
    if (splits < 1) splits = 1;
  }
  if (splits > p.ptiles) splits = p.ptiles;
  if (d->per_image) {
    const int tpi = p.tiles_w * p.tiles_h;  // pixel tiles per image
    SPYR_REQUIRE(p.TN == 1 && p.tiles_n == d->B, "conv2d_wgrad: per_image needs H*W >= 64");
    int s = 1;
    while (s * 2 <= tpi && tpi % (s * 2) == 0 && mtiles * ntiles * d->B * s * 2 <= 2 * 148) s *= 2;
    p.per_image = s;
    splits = d->B * s;
  }
  p.splits = splits;
  const int stage_bytes = (2 + bn / 64) * KP * 128;
  int stages = d->stages;
  if (stages == 0) {
    stages = (192 * 1024) / stage_bytes;
    if (stages > 6) stages = 6;
  }
  p.stages = stages;
  p.tmem_cols = pow2_cols(bn);
  p.lbo = d->dbg_lbo ? (uint32_t)d->dbg_lbo : (uint32_t)(KP * 128);
  p.sbo = d->dbg_sbo ? (uint32_t)d->dbg_sbo : 1024u;
  p.dw = d->dw;
  p.cin_stride = d->cin_stride > 0 ? d->cin_stride : d->Cin;
  SPYR_REQUIRE(!d->per_image || p.cin_stride == d->Cin, "conv2d_wgrad: per_image cannot use cin_stride");

  WgradMaps maps;
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t strides[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    if (spyr_tmap_encode(&maps.x, d->x, 4, dims, strides, box, 1)) return 3;
  }
  {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t strides[3] = {(uint64_t)d->Cout * 2, (uint64_t)d->W * d->Cout * 2, (uint64_t)d->H * d->W * d->Cout * 2};
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    if (spyr_tmap_encode(&maps.dy, d->dy, 4, dims, strides, box, 1)) return 3;
  }
  const size_t smem_bytes = (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
  SPYR_REQUIRE(smem_bytes <= 227 * 1024, "conv2d_wgrad: smem %zu too large", smem_bytes);
  static bool configured = false;
  if (!configured) {
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  dim3 grid(mtiles, ntiles, splits);
  conv_wgrad_kernel<<<grid, 256, smem_bytes, stream>>>(maps, p);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
