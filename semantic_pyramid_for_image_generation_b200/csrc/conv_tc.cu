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
void spyr_note_kernel(int id);
int spyr_conv_halo_launch(const spyr_conv_desc* d, cudaStream_t stream);
int spyr_conv_halo2_launch(const spyr_conv_desc* d, cudaStream_t stream);
int spyr_conv_stack3_launch(const spyr_conv_desc* d, cudaStream_t stream);
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
  int ktaps[SPYR_CONV_MAX_SRC];    // taps per source (1 or 9)
  int kchunks[SPYR_CONV_MAX_SRC];  // ceil(cin/64) per source
  int ktotal;      // total k-steps
  int wmn[SPYR_CONV_MAX_SRC];      // source uses the MN-major (input-gradient) weight view
  int wpi[SPYR_CONV_MAX_SRC];      // source uses per-image weights
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
  int split;                     // split-BF16 mode: outputs / residual are hi + lo plane pairs
  long long plane;               // elements from the hi to the lo plane (B*H*W*Cout)
  float acc_scale;               // split mode: 1 + (hi*hi accumulations) * 2^-25, undoes the accumulator's truncation bias
};

struct TmapPack {
  CUtensorMap x[SPYR_CONV_MAX_SRC];
  CUtensorMap w[SPYR_CONV_MAX_SRC];
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

  // a split of the reduction that got no k-steps still owns a slice of y_f32: its epilogue warps store zeros
  const bool has_work = k_begin < k_end;
  if (has_work || p.splits > 1) {
    if (warp == 0) {
      if (has_work && elect_one()) {
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
    } else if (warp == 1 && has_work) {
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

      if (has_work) {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
      }
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        uint32_t r[32];
        if (has_work) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (!valid) continue;
        const int col0 = n_off + c0;
        if (col0 >= p.Cout) continue;
        if (p.y_f32 != nullptr) {
          // slice blockIdx.z of the split-K partial sums (f32_store: the only slice), plain stores
          float* dst = p.y_f32 + ((size_t)blockIdx.z * ((size_t)p.B * p.H * p.W) + pix) * p.Cout + col0;
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
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // groups of 8 channels = one 16-byte vector
          const int col = col0 + g * 8;
          if (col + 8 > p.Cout) break;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]) * p.acc_scale;
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
          const long long lo = p.split ? p.plane : 0;
          if (p.residual != nullptr) {
            float rv[8];
            ld8(Act(const_cast<bf16*>(p.residual) + off, lo), rv);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += rv[j];
          }
          if (p.y_raw != nullptr) st8(Act(p.y_raw + off, lo), v);
          if (p.y_act != nullptr) {
            const float sl = (p.act == 1) ? 0.f : ((p.act == 2) ? p.act_slope : 1.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * sl;
            st8(Act(p.y_act + off, lo), v);
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
// wgrad: D[(tap,cin) 128][cout block_n] over a split of the pixel tiles.  With one split the accumulator is added to
// dw[tap][cin][cout] directly (each element has one owner); with several, split z stores its partial sum to slice z of a
// scratch buffer and wgrad_reduce_kernel adds the slices in order -- no floating-point atomics, bit-reproducible.
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
  float* partial;   // scratch [slices][nimg][taps*cin_stride*Cout] or nullptr (direct accumulation into dw)
  int slice0;       // first slice of this launch (split-BF16 mode runs three launches into one slice set)
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
  const size_t slice_floats = (size_t)p.taps * p.cin_stride * p.Cout;
  const int img = p.per_image > 0 ? (int)blockIdx.z / p.per_image : 0;
  const int zz = p.per_image > 0 ? (int)blockIdx.z % p.per_image : (int)blockIdx.z;
  const int nimg = p.per_image > 0 ? p.B : 1;
  float* dw_out = p.partial != nullptr ? p.partial + ((size_t)(p.slice0 + zz) * nimg + img) * slice_floats
                                       : p.dw + (size_t)img * slice_floats;

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
      // rows past the destination's row stride are padding channels of the operand (im2col rows 27..31, the 3 -> 8 channel
      // pad of the input block's skip path): they have no slot in dw
      const bool valid = valid_chunk && ci < p.Cin && ci < p.cin_stride;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        if (!valid) continue;
        const int col0 = n_off + c0;
        float* dst = dw_out + ((size_t)tap * p.cin_stride + ci) * p.Cout + col0;
        wgrad_store32(dst, r, p.Cout - col0, p.partial == nullptr);
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

// dw[img][tap][ci][co] += sum over slices of partial[slice][img][tap][ci][co].  `lanes` (a power of two <= 32, chosen from
// the problem shape only) adjacent threads share one float4 of dw: each adds the slices lane, lane + lanes, ... and a fixed
// xor tree combines them, so the order of the additions -- and the result, bit for bit -- does not depend on scheduling.
// Rows ci >= Cin of a strided dw (the mask-channel row of cat(feature*mask, mask)) belong to another kernel: untouched.
__device__ __forceinline__ void wgrad_reduce_body(const float* __restrict__ partial, int nslices, int nimg,
                                                  float* __restrict__ dw, int taps, int Cin, int cin_stride, int Cout,
                                                  int lanes, long long block) {
  const int c4 = Cout >> 2;
  const long long per_img = (long long)taps * Cin * c4;
  const long long total = per_img * nimg;
  const size_t slice_floats = (size_t)taps * cin_stride * Cout;
  const int lane = (int)(threadIdx.x % (unsigned)lanes);
  const long long i = (block * blockDim.x + threadIdx.x) / lanes;
  const bool live = i < total;  // whole lane groups are live or not: the shuffles below stay within a group
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  size_t off = 0;
  if (live) {
    const int img = (int)(i / per_img);
    long long r = i - (long long)img * per_img;
    const int co = (int)(r % c4) * 4;
    r /= c4;
    const int ci = (int)(r % Cin), tap = (int)(r / Cin);
    off = (size_t)img * slice_floats + ((size_t)tap * cin_stride + ci) * Cout + co;
#pragma unroll 4
    for (int sl = lane; sl < nslices; sl += lanes) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(partial + (size_t)sl * nimg * slice_floats + off));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  for (int o = lanes >> 1; o > 0; o >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  if (live && lane == 0) {
    float4 d = *reinterpret_cast<float4*>(dw + off);
    d.x += acc.x; d.y += acc.y; d.z += acc.z; d.w += acc.w;
    *reinterpret_cast<float4*>(dw + off) = d;
  }
}
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nslices, int nimg, float* __restrict__ dw,
                                    int taps, int Cin, int cin_stride, int Cout, int lanes, int) {
  wgrad_reduce_body(partial, nslices, nimg, dw, taps, Cin, cin_stride, Cout, lanes, (long long)blockIdx.x);
}

// Up to SPYR_REDUCE_BATCH pending second stages in one launch: block ranges per entry, same code per entry as the
// immediate kernels (wgrad_reduce_kernel / spyr_sum_partials_kernel), so the results are bit-identical.
struct ReduceBatch {
  spyr_reduce_entry e[SPYR_REDUCE_BATCH];
  int block0[SPYR_REDUCE_BATCH + 1];
  int n;
};
__global__ void reduce_batched_kernel(const __grid_constant__ ReduceBatch b) {
  int k = 0;
  while (k + 1 < b.n && (int)blockIdx.x >= b.block0[k + 1]) ++k;
  const spyr_reduce_entry& e = b.e[k];
  const long long block = (long long)blockIdx.x - b.block0[k];
  if (e.kind == 1) {
    wgrad_reduce_body(e.partial, e.nslices, e.nimg, e.out0, e.taps, e.rows, e.cin_stride, e.Cout, e.lanes, block);
  } else {
    SumSink sk = {e.out0, e.out1, e.out2, e.mode, e.C, e.sink_stride, e.sink_row, e.split};
    spyr_sum_partials_body(e.partial, e.nb, e.n, sk, block);
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

static int conv2d_fprop_impl(const spyr_conv_desc* d, cudaStream_t stream);

extern "C" int spyr_conv2d_fprop(const spyr_conv_desc* d, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPYR_REQUIRE(d != nullptr, "conv2d_fprop: null descriptor");
  SPYR_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cout > 0, "conv2d_fprop: bad shape");
  if (!spyr_split()) {
    SPYR_REQUIRE(d->nsrc >= 1 && d->nsrc <= SPYR_CONV_MAX_SRC, "conv2d_fprop: nsrc=%d out of range", d->nsrc);
    return conv2d_fprop_impl(d, stream);
  }
  // split-BF16 mode: x = x_hi + x_lo, w = w_hi + w_lo (lo planes behind the hi planes); every source becomes the three
  // products x_lo*w_hi + x_hi*w_lo + x_hi*w_hi accumulated in FP32 (x_lo*w_lo is below 2^-17 of the result).
  // Order matters: the tensor pipe's FP32 accumulator truncates (round toward zero) on every accumulation, a relative
  // 2^-24 of the CURRENT accumulator magnitude per tcgen05.mma (tools/probe_split_accum.py: a 4608-long reduction with
  // the large products first shrinks by 2.2e-5).  All small (lo) products therefore go first, while the accumulator is
  // still 2^-9 of its final size, and the hi*hi products of all sources last: sources [0, 2n) = lo products,
  // [2n, 3n) = hi*hi.  The launchers compensate the remaining, predictable shrink of the hi*hi chain (acc_scale).
  SPYR_REQUIRE(d->nsrc >= 1 && 3 * d->nsrc <= SPYR_CONV_MAX_SRC, "conv2d_fprop: nsrc=%d out of range in split-BF16 mode",
               d->nsrc);
  spyr_conv_desc e = *d;
  e.nsrc = 3 * d->nsrc;
  for (int s = 0; s < d->nsrc; ++s) {
    const spyr_conv_src& src = d->src[s];
    const size_t xn = (size_t)d->B * d->H * d->W * src.cin;
    const size_t wslices = src.w_per_image ? (size_t)d->B : (size_t)(src.ksize * src.ksize);
    const size_t wn = src.w_lo_off > 0 ? (size_t)src.w_lo_off : wslices * (size_t)d->Cout * src.cin;
    const bf16* x_hi = (const bf16*)src.x;
    const bf16* w_hi = (const bf16*)src.w;
    const int n = d->nsrc;
    e.src[2 * s] = src;          // x_lo * w_hi
    e.src[2 * s].x = x_hi + xn;
    e.src[2 * s + 1] = src;      // x_hi * w_lo
    e.src[2 * s + 1].w = w_hi + wn;
    e.src[2 * n + s] = src;      // x_hi * w_hi
  }
  return conv2d_fprop_impl(&e, stream);
}

static int conv2d_fprop_impl(const spyr_conv_desc* d, cudaStream_t stream) {
  SPYR_REQUIRE(is_pow2(d->H) && is_pow2(d->W) || (d->H == 1 && d->W == 1), "conv2d_fprop: H,W must be powers of two");
  {
    // maps of 16x8 pixels and larger run on the persistent halo-tiled kernel (conv_halo.cu); SPYR_CONV_LEGACY=1 forces
    // the per-tap kernel below (A/B measurements)
    static const bool legacy = getenv("SPYR_CONV_LEGACY") != nullptr;
    static const bool no_pair = getenv("SPYR_CONV_NO_PAIR") != nullptr;
    if (!legacy) {
      {
        // 64 output channels, 3x3, maps >= 128 wide: the three taps of a kernel row stacked along N (conv_stack3.cu)
        const int rc3 = spyr_conv_stack3_launch(d, stream);
        if (rc3 >= 0) return rc3;
      }
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
  for (int s = d->nsrc; s < SPYR_CONV_MAX_SRC; ++s) {
    maps.x[s] = maps.x[0];
    maps.w[s] = maps.w[0];
  }
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
  p.split = (spyr_split() && d->y_f32 == nullptr) ? 1 : 0;
  p.plane = (long long)d->B * d->H * d->W * d->Cout;
  p.acc_scale = 1.f;
  if (p.split) {
    int hh_steps = 0;  // k-steps of the hi*hi sources (the last third, see spyr_conv2d_fprop), 4 accumulations each
    for (int s = 2 * (d->nsrc / 3); s < d->nsrc; ++s) hh_steps += p.ktaps[s] * p.kchunks[s];
    p.acc_scale = 1.f + (float)(4 * hh_steps) * 2.9802322e-8f;
  }

  const size_t smem_bytes = (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
  SPYR_REQUIRE(smem_bytes <= 227 * 1024, "conv2d_fprop: smem %zu too large", smem_bytes);
  static size_t configured = 0;
  if (smem_bytes > configured) {
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(conv_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  dim3 grid(p.tiles_w * p.tiles_h * tiles_n, ceil_div(d->Cout, bn), p.splits);
  conv_fprop_kernel<<<grid, 256, smem_bytes, stream>>>(maps, p);
  spyr_note_kernel(2);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}

// Planning shared by spyr_conv2d_wgrad and spyr_conv2d_wgrad_scratch_floats.
int spyr_wgrad_halo_plan(const spyr_wgrad_desc* d, int* splits);
int spyr_wgrad_halo_launch(const spyr_wgrad_desc* d, const void* x, const void* dy, float* partial, int slice0,
                           cudaStream_t stream);

static int wgrad_tc_plan(const spyr_wgrad_desc* d, WgradParams* pp) {
  WgradParams& p = *pp;
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
    // Every split beyond the first costs a full FP32 copy of the gradient (its slice of partial sums is written, then read
    // by the fixed-order reduction): split only while each CTA keeps >= 16 pixel tiles of reduction and the grid stays
    // within one wave.  (With the round-1 rule, two waves of 4-5 k-step CTAs, the 8x8 / 16x16 layers spent 3/4 of their
    // time moving partial sums.)
    splits = 148 / (mtiles * ntiles);
    const int by_work = p.ptiles / 16;
    if (splits > by_work) splits = by_work;
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
  } else {
    // every split must own at least one pixel tile: its slice of the partial sums is read unconditionally
    const int per = ceil_div(p.ptiles, splits);
    splits = ceil_div(p.ptiles, per);
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
  return 0;
}

static int wgrad_tc_launch(const spyr_wgrad_desc* d, const void* x, const void* dy, float* partial, int slice0,
                           cudaStream_t stream) {
  WgradParams p;
  const int rc = wgrad_tc_plan(d, &p);
  if (rc) return rc;
  p.partial = partial;
  p.slice0 = slice0;
  const int bn = p.block_n;
  WgradMaps maps;
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t strides[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    if (spyr_tmap_encode(&maps.x, x, 4, dims, strides, box, 1)) return 3;
  }
  {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t strides[3] = {(uint64_t)d->Cout * 2, (uint64_t)d->W * d->Cout * 2, (uint64_t)d->H * d->W * d->Cout * 2};
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    if (spyr_tmap_encode(&maps.dy, dy, 4, dims, strides, box, 1)) return 3;
  }
  const int stage_bytes = (2 + bn / 64) * KP * 128;
  const size_t smem_bytes = (size_t)p.stages * stage_bytes + (2 * p.stages + 1) * 8 + 16 + 1024;
  SPYR_REQUIRE(smem_bytes <= 227 * 1024, "conv2d_wgrad: smem %zu too large", smem_bytes);
  static bool configured = false;
  if (!configured) {
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  dim3 grid(ceil_div(p.mchunks, 2), ceil_div(d->Cout, bn), p.splits);
  conv_wgrad_kernel<<<grid, 256, smem_bytes, stream>>>(maps, p);
  spyr_note_kernel(4);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}

// route: 0 = halo-tiled kernel, 1 = per-tap kernel; *slices = partial-sum slices per image (1: accumulate into dw directly)
static int wgrad_route(const spyr_wgrad_desc* d, int* route, int* slices) {
  SPYR_REQUIRE(d != nullptr, "conv2d_wgrad: null descriptor");
  SPYR_REQUIRE(d->ksize == 1 || d->ksize == 3, "conv2d_wgrad: ksize must be 1 or 3");
  SPYR_REQUIRE(d->Cin % 8 == 0 && d->Cout % 8 == 0, "conv2d_wgrad: Cin/Cout must be multiples of 8");
  SPYR_REQUIRE((is_pow2(d->H) && is_pow2(d->W)) || (d->H == 1 && d->W == 1), "conv2d_wgrad: H,W must be powers of two");
  static const bool legacy = getenv("SPYR_CONV_LEGACY") != nullptr;
  int splits = 0;
  *route = 1;
  if (!legacy && spyr_wgrad_halo_plan(d, &splits) == 0) {
    *route = 0;
  } else {
    WgradParams p;
    const int rc = wgrad_tc_plan(d, &p);
    if (rc) return rc;
    splits = p.per_image > 0 ? p.per_image : p.splits;
  }
  *slices = splits * (spyr_split() ? 3 : 1);
  return 0;
}

extern "C" long long spyr_conv2d_wgrad_scratch_floats(const spyr_wgrad_desc* d) {
  int route = 0, slices = 0;
  if (wgrad_route(d, &route, &slices)) return -1;
  if (slices <= 1) return 0;
  const long long cs = d->cin_stride > 0 ? d->cin_stride : d->Cin;
  return (long long)slices * (d->per_image ? d->B : 1) * d->ksize * d->ksize * cs * d->Cout;
}

static int wgrad_impl(const spyr_wgrad_desc* d, spyr_reduce_entry* deferred, cudaStream_t stream);

extern "C" int spyr_conv2d_wgrad(const spyr_wgrad_desc* d, void* stream_) {
  return wgrad_impl(d, nullptr, (cudaStream_t)stream_);
}
extern "C" int spyr_conv2d_wgrad_deferred(const spyr_wgrad_desc* d, spyr_reduce_entry* out, void* stream_) {
  SPYR_REQUIRE(out != nullptr, "conv2d_wgrad_deferred: out is NULL");
  memset(out, 0, sizeof(*out));
  return wgrad_impl(d, out, (cudaStream_t)stream_);
}

static int wgrad_impl(const spyr_wgrad_desc* d, spyr_reduce_entry* deferred, cudaStream_t stream) {
  int route = 0, slices = 0;
  const int rc0 = wgrad_route(d, &route, &slices);
  if (rc0) return rc0;
  const long long need = spyr_conv2d_wgrad_scratch_floats(d);
  SPYR_REQUIRE(need == 0 || (d->scratch != nullptr && d->scratch_floats >= need),
               "conv2d_wgrad: scratch of %lld floats needed (spyr_conv2d_wgrad_scratch_floats), got %lld", need,
               d->scratch != nullptr ? d->scratch_floats : 0LL);
  float* partial = slices > 1 ? d->scratch : nullptr;
  // split-BF16 mode: dw = x_hi*dy_hi + x_lo*dy_hi + x_hi*dy_lo, three launches into consecutive slice groups
  const int nsub = spyr_split() ? 3 : 1;
  const int per_sub = slices / nsub;
  const size_t xn = (size_t)d->B * d->H * d->W * d->Cin, yn = (size_t)d->B * d->H * d->W * d->Cout;
  for (int sub = 0; sub < nsub; ++sub) {
    const bf16* x = (const bf16*)d->x + (sub == 1 ? xn : 0);
    const bf16* dy = (const bf16*)d->dy + (sub == 2 ? yn : 0);
    const int rc = route == 0 ? spyr_wgrad_halo_launch(d, x, dy, partial, sub * per_sub, stream)
                              : wgrad_tc_launch(d, x, dy, partial, sub * per_sub, stream);
    if (rc) return rc;
  }
  if (partial != nullptr) {
    const int nimg = d->per_image ? d->B : 1;
    const int cs = d->cin_stride > 0 ? d->cin_stride : d->Cin;
    const int rows = cs < d->Cin ? cs : d->Cin;  // operand channels that have a row in dw
    const long long total = (long long)nimg * d->ksize * d->ksize * rows * (d->Cout / 4);
    int lanes = 1;  // threads per output vector: enough parallelism for small gradients with many slices
    while (lanes < 32 && lanes < slices && total * lanes < 131072) lanes *= 2;
    if (deferred != nullptr) {
      // the caller batches this sum with the others of its backward pass (spyr_reduce_batched)
      deferred->kind = 1;
      deferred->partial = partial;
      deferred->out0 = d->dw;
      deferred->nslices = slices; deferred->nimg = nimg; deferred->taps = d->ksize * d->ksize; deferred->rows = rows;
      deferred->cin_stride = cs; deferred->Cout = d->Cout; deferred->lanes = lanes;
      return 0;
    }
    const long long blocks = (total * lanes + 255) / 256;
    wgrad_reduce_kernel<<<(int)blocks, 256, 0, stream>>>(partial, slices, nimg, d->dw, d->ksize * d->ksize, rows, cs,
                                                        d->Cout, lanes, 0);
    spyr_count_launch();
    SPYR_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int spyr_reduce_batched(const spyr_reduce_entry* entries, int n, void* stream) {
  SPYR_REQUIRE(entries != nullptr && n >= 1 && n <= SPYR_REDUCE_BATCH, "reduce_batched: n=%d out of range", n);
  ReduceBatch b;
  memset(&b, 0, sizeof(b));
  int blocks = 0, m = 0;
  for (int i = 0; i < n; ++i) {
    const spyr_reduce_entry& e = entries[i];
    if (e.kind == 0) continue;
    long long nb;
    if (e.kind == 1) {
      const long long total = (long long)e.nimg * e.taps * e.rows * (e.Cout / 4);
      nb = (total * e.lanes + 255) / 256;
    } else {
      SPYR_REQUIRE(e.kind == 2, "reduce_batched: entry %d has kind %d", i, e.kind);
      nb = (e.n + 7) / 8;
    }
    b.e[m] = e;
    b.block0[m] = blocks;
    blocks += (int)nb;
    ++m;
  }
  if (m == 0) return 0;
  b.block0[m] = blocks;
  b.n = m;
  reduce_batched_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(b);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
