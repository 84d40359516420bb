// Persistent, halo-tiled implicit-GEMM convolution for feature maps of 16x16 pixels and larger (sm_100a).
//
// Why a second conv kernel: with one TMA load per (tap, 64-channel chunk) the kernel in conv_tc.cu moves
// 9 x 16 KB of activations from L2 to shared memory for every 128-pixel tile and is L2->SMEM bandwidth bound
// (~12 TB/s chip-wide, measured: 24 % of the tensor peak).  Here the activation operand of a 3x3 convolution is ONE
// halo tile per 64-channel chunk -- (8+2) x (16*MSUB+2) pixels, 128 B per pixel, SWIZZLE_128B, written by a single
// TMA box whose out-of-bounds rows are the convolution's zero padding -- and the nine taps are nine tcgen05.mma
// A-descriptors that start at different 128-byte rows of that tile: output pixel (th, tw) of tap (dy, dx) reads halo
// row (th+dy)*(8+2) + (tw+dx), so every 8-pixel output row is an 8-row core-matrix group and the groups are a
// constant (8+2)*128 bytes apart (the descriptor's stride-byte-offset).  The swizzle XOR is a function of the
// absolute shared-memory address, so unaligned start rows need no base offset (validated by tests/native, `halo`).
//
//   tile          256 (MSUB=2) or 128 output pixels of one image x BLOCK_N output channels
//   A traffic     1 halo tile per chunk instead of 9 tap tiles  (6.4x less)
//   B traffic     each weight stage (tap, chunk) feeds MSUB x 4 MMAs (2x less per FLOP at MSUB=2)
//   schedule      persistent CTAs (grid = #SMs), static round-robin over tiles
//   pipelines     A ring (2 halo buffers), B ring (weight stages), 2 TMEM accumulator sets: the epilogue of tile i
//                 overlaps the TMA + MMA of tile i+1
//   warps         0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..11 = epilogue (two column halves x four
//                 TMEM lane quarters)
//
// Same sources / epilogue contract as spyr_conv2d_fprop (see include/spyramid_b200.h); split-K and maps smaller than
// 16x8 stay on the conv_tc.cu kernel.
#include "common.cuh"
#include "../../include/spyramid_b200.h"

extern void spyr_count_launch();

namespace {

constexpr int KC = 64;
constexpr int THREADS = 384;
constexpr int EPI_WARPS = 8;
constexpr int A_BUFS = 2;

struct HaloParams {
  int B, H, W, Cout;
  int msub;
  int tiles_w, tiles_h, m_tiles, n_tiles, total_tiles;
  int nsrc;
  int border[3];   // 1: 3x3 conv, 0: 1x1
  int kchunks[3];
  int wmn[3];
  int wpi[3];
  int a_rows[3];   // rows (pixels) of the A box of this source
  int block_n, bn_cols;  // bn_cols: TMEM column stride of one accumulator (power of two >= 32)
  int a_buf_bytes, b_stage_bytes, b_stages;
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

struct HaloMaps {
  CUtensorMap x[3];
  CUtensorMap w[3];
};

// Per-tile epilogue constants staged in shared memory by the epilogue warps: the summed bias vectors of this N block and
// the ten FP32 stencil rows of the mask channel.  Every lane of a warp reads the same address (broadcast).
struct EpiConst {
  const float* bias;     // [block_n]
  const float* stencil;  // [10][block_n] or nullptr
};

__device__ __forceinline__ void epilogue_chunk(const HaloParams& p, const uint32_t* r, size_t pix, int col0, int c0,
                                               const EpiConst& ec, const float* mk, int mk_mode) {
  if (col0 >= p.Cout) return;
  if (p.y_f32 != nullptr) {
    float* dst = p.y_f32 + pix * p.Cout + col0;
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
    return;
  }
  const size_t off0 = pix * p.Cout + col0;
  // issue every global read of this 32-channel chunk before any arithmetic (read-only path, independent of the stores)
  uint4 dm[4], rs[4];
  if (p.dmask != nullptr) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (col0 + g * 8 + 8 <= p.Cout) dm[g] = __ldg(reinterpret_cast<const uint4*>(p.dmask + off0 + g * 8));
  }
  if (p.residual != nullptr) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (col0 + g * 8 + 8 <= p.Cout) rs[g] = __ldg(reinterpret_cast<const uint4*>(p.residual + off0 + g * 8));
  }
  const float sl = (p.act == 1) ? 0.f : ((p.act == 2) ? p.act_slope : 1.f);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    if (col + 8 > p.Cout) break;
    float v[8];
    const float4 b0 = *reinterpret_cast<const float4*>(ec.bias + c0 + g * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(ec.bias + c0 + g * 8 + 4);
    v[0] = __uint_as_float(r[g * 8 + 0]) + b0.x;
    v[1] = __uint_as_float(r[g * 8 + 1]) + b0.y;
    v[2] = __uint_as_float(r[g * 8 + 2]) + b0.z;
    v[3] = __uint_as_float(r[g * 8 + 3]) + b0.w;
    v[4] = __uint_as_float(r[g * 8 + 4]) + b1.x;
    v[5] = __uint_as_float(r[g * 8 + 5]) + b1.y;
    v[6] = __uint_as_float(r[g * 8 + 6]) + b1.z;
    v[7] = __uint_as_float(r[g * 8 + 7]) + b1.w;
    if (mk_mode == 1) {
      const float* st = ec.stencil + 9 * p.block_n + c0 + g * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += st[j];
    } else if (mk_mode == 2) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (mk[t] != 0.f) {
          const float* st = ec.stencil + t * p.block_n + c0 + g * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += mk[t] * st[j];
        }
      }
    }
    if (p.dmask != nullptr) {
      const uint32_t mw[4] = {dm[g].x, dm[g].y, dm[g].z, dm[g].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(mw[j]);
        if (!(f.x > 0.f)) v[2 * j] *= p.dmask_slope;
        if (!(f.y > 0.f)) v[2 * j + 1] *= p.dmask_slope;
      }
    }
    if (p.residual != nullptr) {
      const uint32_t rw[4] = {rs[g].x, rs[g].y, rs[g].z, rs[g].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(rw[j]);
        v[2 * j] += f.x;
        v[2 * j + 1] += f.y;
      }
    }
    const size_t off = off0 + g * 8;
    if (p.y_raw != nullptr) {
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]);
      o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]);
      o.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(p.y_raw + off) = o;
    }
    if (p.y_act != nullptr) {
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

__global__ void __launch_bounds__(THREADS, 1)
conv_halo_kernel(const __grid_constant__ HaloMaps maps, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + A_BUFS * p.a_buf_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + p.b_stages * p.b_stage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + A_BUFS;
  uint64_t* b_full = a_empty + A_BUFS;
  uint64_t* b_empty = b_full + p.b_stages;
  uint64_t* acc_full = b_empty + p.b_stages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* epi_const = reinterpret_cast<float*>(tmem_holder + 4);  // [2][11][block_n]: bias sum + 10 stencil rows

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) {
      tma_prefetch_desc(&maps.x[s]);
      tma_prefetch_desc(&maps.w[s]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < A_BUFS; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], EPI_WARPS);
    }
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
  const int b_chunks = (p.block_n + 63) >> 6;
  const int th_rows = 16 * p.msub;

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      int abuf = 0, bst = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int m_tile = tile % p.m_tiles, n_tile = tile / p.m_tiles;
        const int w0 = (m_tile % p.tiles_w) * 8;
        const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * th_rows;
        const int n0 = m_tile / (p.tiles_w * p.tiles_h);
        const int n_off = n_tile * p.block_n;
        for (int s = 0; s < p.nsrc; ++s) {
          const int bd = p.border[s];
          const int taps = bd ? 9 : 1;
          for (int c = 0; c < p.kchunks[s]; ++c) {
            mbar_wait(&a_empty[abuf], aph ^ 1);
            mbar_arrive_expect_tx(&a_full[abuf], (uint32_t)(p.a_rows[s] * 128));
            tma_load_4d(a_ring + abuf * p.a_buf_bytes, &maps.x[s], &a_full[abuf], c * KC, w0 - bd, h0 - bd, n0);
            if (++abuf == A_BUFS) {
              abuf = 0;
              aph ^= 1;
            }
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(&b_empty[bst], bph ^ 1);
              uint8_t* b_dst = b_ring + bst * p.b_stage_bytes;
              if (p.wmn[s]) {
                mbar_arrive_expect_tx(&b_full[bst], (uint32_t)(b_chunks * 8192));
                const int wtap = p.wpi[s] ? n0 : (bd ? 8 - tap : 0);
                for (int j = 0; j < b_chunks; ++j)
                  tma_load_3d(b_dst + j * 8192, &maps.w[s], &b_full[bst], n_off + j * 64, c * KC, wtap);
              } else {
                mbar_arrive_expect_tx(&b_full[bst], (uint32_t)(p.block_n * 128));
                tma_load_3d(b_dst, &maps.w[s], &b_full[bst], c * KC, n_off, p.wpi[s] ? n0 : tap);
              }
              if (++bst == p.b_stages) {
                bst = 0;
                bph ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: ONE thread runs the whole loop.  Descriptors are 64-bit integers advanced by plain adds on the
    // 14-bit start-address field (16-byte units); recomputing them per MMA made this thread, not the tensor pipe, the
    // bottleneck (ncu: 22 dependent instructions per tcgen05.mma).
    if (lane == 0) {
      const uint32_t idesc_k = umma_idesc_bf16(128, p.block_n, 0, 0);
      const uint32_t idesc_mn = umma_idesc_bf16(128, p.block_n, 0, 1);
      const uint64_t desc_base = ((uint64_t)1 << 46) | ((uint64_t)2 << 61);  // version 1, SWIZZLE_128B
      const uint32_t a_ring_lo = (smem_u32(a_ring) & 0x3FFFF) >> 4;
      const uint32_t b_ring_lo = (smem_u32(b_ring) & 0x3FFFF) >> 4;
      const uint32_t a_buf16 = (uint32_t)p.a_buf_bytes >> 4, b_stage16 = (uint32_t)p.b_stage_bytes >> 4;
      int abuf = 0, bst = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&acc_empty[buf], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * p.msub * p.bn_cols);
        uint32_t accum = 0;  // 0 only for the first MMA of each sub-tile accumulator
        for (int s = 0; s < p.nsrc; ++s) {
          const int bd = p.border[s];
          const int taps = bd ? 9 : 1;
          const uint32_t row16 = 8;                                  // one 128-byte pixel row in 16-byte units
          const uint32_t pitch16 = (uint32_t)(8 + 2 * bd) * row16;  // consecutive 8-pixel output rows
          const uint32_t sub16 = 16u * pitch16;                      // second 8x16 sub-tile
          const bool mn = p.wmn[s] != 0;
          const uint64_t a_hi = desc_base | ((uint64_t)pitch16 << 32);
          const uint64_t b_hi = mn ? (desc_base | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32))
                                   : (desc_base | ((uint64_t)(1024 >> 4) << 32));
          const uint32_t bk16 = mn ? (2048u >> 4) : 2u;  // K advance of 16 elements
          const uint32_t idesc = mn ? idesc_mn : idesc_k;
          for (int c = 0; c < p.kchunks[s]; ++c) {
            mbar_wait(&a_full[abuf], aph);
            const uint32_t a_lo = a_ring_lo + (uint32_t)abuf * a_buf16;
            uint32_t row0 = 0;  // halo row of output pixel (0,0) for the current tap: dy*10 + dx
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(&b_full[bst], bph);
              tc_fence_after();
              const uint64_t db0 = b_hi | (uint64_t)(b_ring_lo + (uint32_t)bst * b_stage16);
              const uint64_t da0 = a_hi | (uint64_t)(a_lo + row0 * row16);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                umma_bf16(acc0, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(bk16 * k), idesc, k == 0 ? accum : 1u);
              if (p.msub == 2) {
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  umma_bf16(acc0 + (uint32_t)p.bn_cols, da0 + (uint64_t)(sub16 + 2 * k), db0 + (uint64_t)(bk16 * k), idesc,
                            k == 0 ? accum : 1u);
              }
              accum = 1;
              umma_commit(&b_empty[bst]);
              if (++bst == p.b_stages) {
                bst = 0;
                bph ^= 1;
              }
              row0 += ((tap % 3) == 2) ? 8u : 1u;  // dx wraps: next halo row block (10 - 2)
            }
            umma_commit(&a_empty[abuf]);
            if (++abuf == A_BUFS) {
              abuf = 0;
              aph ^= 1;
            }
          }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int m = q * 32 + lane;
    const int et = threadIdx.x - 128;  // 0..255 among the epilogue threads
    int it = 0;
    int staged_n_off = -1, cbuf = 1;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      const int m_tile = tile % p.m_tiles, n_tile = tile / p.m_tiles;
      const int w0 = (m_tile % p.tiles_w) * 8;
      const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * th_rows;
      const int n = m_tile / (p.tiles_w * p.tiles_h);
      const int n_off = n_tile * p.block_n;
      if (n_off != staged_n_off) {
        // new N block: stage its constants into the other buffer, then one barrier among the 8 epilogue warps.  Warps
        // reach this barrier only after finishing the previous tile, so the buffer being overwritten is no longer read.
        cbuf ^= 1;
        float* dst = epi_const + cbuf * 11 * p.block_n;
        for (int c = et; c < p.block_n; c += EPI_WARPS * 32) {
          const int col = n_off + c;
          float b = 0.f;
          if (col < p.Cout) {
            if (p.bias != nullptr) b += __ldg(&p.bias[col]);
            if (p.bias2 != nullptr) b += __ldg(&p.bias2[col]);
            if (p.bias3 != nullptr) b += __ldg(&p.bias3[col]);
          }
          dst[c] = b;
          if (p.stencil_w != nullptr) {
#pragma unroll
            for (int t = 0; t < 10; ++t) dst[(1 + t) * p.block_n + c] = col < p.Cout ? __ldg(&p.stencil_w[t * p.Cout + col]) : 0.f;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        staged_n_off = n_off;
      }
      EpiConst ec;
      ec.bias = epi_const + cbuf * 11 * p.block_n;
      ec.stencil = ec.bias + p.block_n;
      const int w = w0 + (m & 7);
      float mk[2][9];
      int mk_mode[2] = {0, 0};
      if (p.stencil_mask != nullptr) {
        for (int sub = 0; sub < p.msub; ++sub) {
          const int h = h0 + sub * 16 + (m >> 3);
          bool all0 = true, all1 = true;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
            float v = 0.f;
            if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) v = __ldg(&p.stencil_mask[((size_t)n * p.H + hh) * p.W + ww]);
            mk[sub][t] = v;
            all0 = all0 && (v == 0.f);
            all1 = all1 && (v == 1.f);
          }
          mk_mode[sub] = all0 ? 0 : (all1 ? 1 : 2);
        }
      }
      mbar_wait(&acc_full[buf], acc_ph);
      tc_fence_after();
      for (int sub = 0; sub < p.msub; ++sub) {
        const int h = h0 + sub * 16 + (m >> 3);
        const size_t pix = ((size_t)n * p.H + h) * p.W + w;
        const uint32_t acc = tmem_base + (uint32_t)((buf * p.msub + sub) * p.bn_cols) + ((uint32_t)(q * 32) << 16);
        for (int c0 = half * 32; c0 < p.block_n; c0 += 64) {
          uint32_t r[32];
          tmem_ld32(acc + (uint32_t)c0, r);
          tmem_ld_wait();
          epilogue_chunk(p, r, pix, n_off + c0, c0, ec, mk[sub], mk_mode[sub]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

uint32_t pow2_at_least(int n, uint32_t lo) {
  uint32_t c = lo;
  while ((int)c < n) c <<= 1;
  return c;
}

}  // namespace

// Returns 0 on launch, -1 when the problem is not eligible (caller falls back to the per-tap kernel), >0 on error.
int spyr_conv_halo_launch(const spyr_conv_desc* d, cudaStream_t stream) {
  if (d->H < 16 || d->W < 8 || (d->H % 16) != 0 || (d->W % 8) != 0) return -1;
  if (d->splits > 1) return -1;
  if (d->y_f32 != nullptr && !d->f32_store) return -1;
  if (d->y_f32 == nullptr && (d->Cout % 8) != 0) return -1;
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cout = d->Cout;
  p.nsrc = d->nsrc;
  int bn = d->block_n;
  if (bn == 0) {
    // tcgen05.mma reads both operands from shared memory at ~64 B/clk/SM (measured with ncu: tc wavefronts), so the
    // 4 KB A slice of every M=128 instruction costs 64 clk: only N >= 256 per instruction keeps the tensor pipe above
    // 2/3 busy.  Wide layers therefore use one 128-pixel sub-tile x 256 channels, narrow ones two sub-tiles x Cout.
    bn = d->Cout >= 256 ? 256 : (d->Cout >= 128 ? 128 : ((d->Cout + 15) / 16) * 16);
    if (bn < 32) bn = 32;
  }
  if (bn > 256 || bn % 16 != 0) return -1;
  p.msub = (d->H % 32 == 0 && bn <= 128) ? 2 : 1;
  p.block_n = bn;
  p.bn_cols = (int)pow2_at_least(bn, 32);
  p.tmem_cols = pow2_at_least(2 * p.msub * p.bn_cols, 32);
  if (p.tmem_cols > 512) return -1;
  p.tiles_w = d->W / 8;
  p.tiles_h = d->H / (16 * p.msub);
  p.m_tiles = p.tiles_w * p.tiles_h * d->B;
  p.n_tiles = ceil_div(d->Cout, bn);
  p.total_tiles = p.m_tiles * p.n_tiles;
  HaloMaps maps;
  int max_rows = 0;
  p.b_stage_bytes = bn * 128;
  for (int s = 0; s < d->nsrc; ++s) {
    const spyr_conv_src& src = d->src[s];
    SPYR_REQUIRE(src.ksize == 1 || src.ksize == 3, "conv2d_fprop: ksize must be 1 or 3");
    SPYR_REQUIRE(src.cin % 8 == 0 && src.cin > 0, "conv2d_fprop: cin=%d must be a multiple of 8", src.cin);
    SPYR_REQUIRE(((uintptr_t)src.x & 15) == 0 && ((uintptr_t)src.w & 15) == 0, "conv2d_fprop: unaligned pointer");
    SPYR_REQUIRE(!src.w_per_image || src.ksize == 1, "conv2d_fprop: per-image weights need ksize 1");
    SPYR_REQUIRE(!src.w_mn_major || (d->Cout % 8 == 0), "conv2d_fprop: MN-major weights need Cout %% 8 == 0");
    const int bd = src.ksize == 3 ? 1 : 0;
    p.border[s] = bd;
    p.kchunks[s] = ceil_div(src.cin, KC);
    p.wmn[s] = src.w_mn_major ? 1 : 0;
    p.wpi[s] = src.w_per_image ? 1 : 0;
    const int bw = 8 + 2 * bd, bh = 16 * p.msub + 2 * bd;
    p.a_rows[s] = bw * bh;
    if (p.a_rows[s] > max_rows) max_rows = p.a_rows[s];
    if (src.w_mn_major && ceil_div(bn, 64) * 8192 > p.b_stage_bytes) p.b_stage_bytes = ceil_div(bn, 64) * 8192;
    {
      uint64_t dims[4] = {(uint64_t)src.cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
      uint64_t strides[3] = {(uint64_t)src.cin * 2, (uint64_t)d->W * src.cin * 2, (uint64_t)d->H * d->W * src.cin * 2};
      uint32_t box[4] = {KC, (uint32_t)bw, (uint32_t)bh, 1};
      if (spyr_tmap_encode(&maps.x[s], src.x, 4, dims, strides, box, 1)) return 3;
    }
    const uint64_t wslices = src.w_per_image ? (uint64_t)d->B : (uint64_t)(src.ksize * src.ksize);
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
  p.a_buf_bytes = ceil_div(max_rows * 128, 1024) * 1024;
  const int budget = 200 * 1024 - A_BUFS * p.a_buf_bytes;
  int stages = budget / p.b_stage_bytes;
  if (stages > 12) stages = 12;
  if (d->stages > 0 && d->stages < stages) stages = d->stages;
  if (stages < 2) return -1;
  p.b_stages = stages;
  p.bias = d->bias; p.bias2 = d->bias2; p.bias3 = d->bias3;
  p.stencil_mask = d->stencil_mask; p.stencil_w = d->stencil_w;
  p.dmask = (const bf16*)d->dmask; p.dmask_slope = d->dmask_slope;
  p.residual = (const bf16*)d->residual;
  p.y_raw = (bf16*)d->y_raw; p.y_act = (bf16*)d->y_act;
  p.act = d->act; p.act_slope = d->act_slope;
  p.y_f32 = d->y_f32;
  const size_t smem_bytes = (size_t)A_BUFS * p.a_buf_bytes + (size_t)stages * p.b_stage_bytes +
                            (2 * A_BUFS + 2 * stages + 4) * 8 + 16 + (size_t)2 * 11 * bn * 4 + 1024;
  static bool configured = false;
  if (!configured) {
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SPYR_CHECK_CUDA(cudaGetDevice(&dev));
    SPYR_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  conv_halo_kernel<<<grid, THREADS, smem_bytes, stream>>>(maps, p);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
