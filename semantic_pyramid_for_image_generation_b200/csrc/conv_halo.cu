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
#include <cstdlib>

#include "conv_halo_common.cuh"

extern void spyr_count_launch();
void spyr_note_kernel(int id);

namespace {
using namespace halo;

template <bool SPLIT, int EPI>
__global__ void __launch_bounds__(THREADS, 1)
conv_halo_kernel(const __grid_constant__ HaloMaps maps, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + p.a_bufs * p.a_buf_bytes;
  uint8_t* epi_stage = b_ring + p.b_stages * p.b_stage_bytes;  // EPI_STAGE_TOTAL bytes when p.tma_store
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + (p.tma_store ? EPI_STAGE_TOTAL : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.a_bufs;
  uint64_t* b_full = a_empty + p.a_bufs;
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
    if (p.tma_store) {
      tma_prefetch_desc(&maps.y[0]);
      tma_prefetch_desc(&maps.y[1]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.a_bufs; ++i) {
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
    {
      // ===== TMA producer: warp-uniform coordinates, one elected lane issues =====
      const bool issue = elect_one();
      int abuf = 0, bst = 0;
      uint32_t aph = 0, bph = 0;
      bool load_b = true;  // resident weights: only the first tile of this CTA loads them
      HaloTileIter ti;
      ti.init(blockIdx.x, p.m_tiles);
      for (; ti.nt < p.n_tiles; ti.next(gridDim.x, p.m_tiles)) {
        int w0, h0, n0;
        halo_tile_origin(p, ti.m, th_rows, w0, h0, n0);
        const int n_off = ti.nt * p.block_n;
        for (int s = 0; s < p.nsrc; ++s) {
          const int bd = p.border[s];
          const int taps = bd ? 9 : 1;
          for (int c = 0; c < p.kchunks[s]; ++c) {
            mbar_wait(&a_empty[abuf], aph ^ 1);
            if (issue) {
              mbar_arrive_expect_tx(&a_full[abuf], (uint32_t)(p.a_rows[s] * 128));
              tma_load_4d(a_ring + abuf * p.a_buf_bytes, &maps.x[s], &a_full[abuf], c * KC, w0 - bd, h0 - bd, n0);
            }
            if (++abuf == p.a_bufs) {
              abuf = 0;
              aph ^= 1;
            }
            for (int tap = 0; load_b && tap < taps; ++tap) {
              mbar_wait(&b_empty[bst], bph ^ 1);
              uint8_t* b_dst = b_ring + bst * p.b_stage_bytes;
              if (issue) {
                if (p.wmn[s]) {
                  mbar_arrive_expect_tx(&b_full[bst], (uint32_t)(b_chunks * 8192));
                  const int wtap = p.wpi[s] ? n0 : (bd ? 8 - tap : 0);
                  for (int j = 0; j < b_chunks; ++j)
                    tma_load_3d(b_dst + j * 8192, &maps.w[s], &b_full[bst], n_off + j * 64, c * KC, wtap);
                } else {
                  mbar_arrive_expect_tx(&b_full[bst], (uint32_t)(p.block_n * 128));
                  tma_load_3d(b_dst, &maps.w[s], &b_full[bst], c * KC, n_off, p.wpi[s] ? n0 : tap);
                }
              }
              if (++bst == p.b_stages) {
                bst = 0;
                bph ^= 1;
              }
            }
          }
        }
        if (p.b_resident) load_b = false;
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer.  The whole warp runs the loop with warp-uniform operands, only the tcgen05 instructions are
    // predicated on one elected lane: the descriptors then live in uniform registers and each UTCHMMA costs one or two
    // uniform adds.  (A loop run by `lane == 0` alone compiles every MMA into an R2UR + ELECT + branch waterfall,
    // ~12 instructions, which capped N=64 layers at 130 clk per MMA against a tensor-pipe floor of 75:
    // tests/native/mma_probe.cu.)  Descriptors are 64-bit integers advanced by plain adds on the 14-bit start-address
    // field (16-byte units).
    {
      const bool leader = elect_one();
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);
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
              if (!p.b_resident)
                mbar_wait(&b_full[bst], bph);
              else if (it == 0)
                mbar_wait(&b_full[bst], 0);
              tc_fence_after();
              const uint64_t db0 = b_hi | (uint64_t)(b_ring_lo + (uint32_t)bst * b_stage16);
              const uint64_t da0 = a_hi | (uint64_t)(a_lo + row0 * row16);
              if (leader) {
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  umma_bf16(acc0, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(bk16 * k), idesc, k == 0 ? accum : 1u);
                if (p.msub == 2) {
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k)
                    umma_bf16(acc0 + (uint32_t)p.bn_cols, da0 + (uint64_t)(sub16 + 2 * k), db0 + (uint64_t)(bk16 * k),
                              idesc, k == 0 ? accum : 1u);
                }
                if (!p.b_resident) umma_commit(&b_empty[bst]);
              }
              accum = 1;
              if (++bst == p.b_stages) {
                bst = 0;
                bph ^= 1;
              }
              row0 += ((tap % 3) == 2) ? 8u : 1u;  // dx wraps: next halo row block (10 - 2)
            }
            if (leader) umma_commit(&a_empty[abuf]);
            if (++abuf == p.a_bufs) {
              abuf = 0;
              aph ^= 1;
            }
          }
        }
        if (leader) umma_commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int m = q * 32 + lane;
    const int et = threadIdx.x - 128;  // 0..255 among the epilogue threads
    int it = 0;
    int staged_n_off = -1, cbuf = 1;
    int sbuf = 0;
    EpiStore es;
    es.maps = p.tma_store ? maps.y : nullptr;
    es.stage = smem_u32(epi_stage) + (uint32_t)((warp - 4) * 2 * EPI_STAGE_BYTES);
    es.lane = lane;
    es.sbuf = &sbuf;
    HaloTileIter ti;
    ti.init(blockIdx.x, p.m_tiles);
    for (; ti.nt < p.n_tiles; ti.next(gridDim.x, p.m_tiles), ++it) {
      const int buf = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      int w0, h0, n;
      halo_tile_origin(p, ti.m, th_rows, w0, h0, n);
      const int n_off = ti.nt * p.block_n;
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
      if (EPI == 0 && p.stencil_mask != nullptr) {
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
      if (p.dmask != nullptr || (p.residual != nullptr && !p.pool && !p.res_pooled)) {
        // the gate / residual operands of this tile are known before its accumulator is: pull them into L2 while the
        // MMAs still run, so the epilogue's loads see L2 latency instead of HBM latency
        for (int sub = 0; sub < p.msub; ++sub) {
          const size_t pix = ((size_t)n * p.H + (h0 + sub * 16 + (m >> 3))) * p.W + w;
          for (int c0 = half * 32; c0 < p.block_n; c0 += 64) {
            if (n_off + c0 >= p.Cout) break;
            const size_t off = pix * p.Cout + n_off + c0;
            if (p.dmask != nullptr) prefetch_l2(p.dmask + off);
            if (p.residual != nullptr && !p.pool && !p.res_pooled) prefetch_l2(p.residual + off);
          }
        }
      }
      mbar_wait(&acc_full[buf], acc_ph);
      tc_fence_after();
      for (int sub = 0; sub < p.msub; ++sub) {
        const int h = h0 + sub * 16 + (m >> 3);
        const size_t pix = ((size_t)n * p.H + h) * p.W + w;
        const uint32_t acc = tmem_base + (uint32_t)((buf * p.msub + sub) * p.bn_cols) + ((uint32_t)(q * 32) << 16);
        es.w = w0;
        es.h = h0 + sub * 16 + q * 4;
        es.n = n;
        for (int c0 = half * 32; c0 < p.block_n; c0 += 64) {
          uint32_t r[32];
          tmem_ld32(acc + (uint32_t)c0, r);
          tmem_ld_wait();
          epilogue_static<SPLIT, EPI>(p, r, pix, n_off + c0, c0, ec, mk[sub], mk_mode[sub], es);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (p.tma_store && lane == 0) bulk_wait_all();  // staged chunks must be read (and written out) before the CTA exits
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
  if ((p.tiles_w & (p.tiles_w - 1)) != 0 || (p.tiles_h & (p.tiles_h - 1)) != 0) return -1;  // shift / mask tile decode
  for (p.lw = 0; (1 << p.lw) < p.tiles_w; ++p.lw) {}
  for (p.lh = 0; (1 << p.lh) < p.tiles_h; ++p.lh) {}
  p.m_tiles = p.tiles_w * p.tiles_h * d->B;
  p.n_tiles = ceil_div(d->Cout, bn);
  p.total_tiles = p.m_tiles * p.n_tiles;
  HaloMaps maps;
  memset(&maps, 0, sizeof(maps));
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
  for (int s = d->nsrc; s < SPYR_CONV_MAX_SRC; ++s) {
    maps.x[s] = maps.x[0];
    maps.w[s] = maps.w[0];
  }
  p.a_buf_bytes = ceil_div(max_rows * 128, 1024) * 1024;
  // Shared-memory plan (218 KB usable next to barriers, epilogue constants and alignment slack).  If every weight slice
  // of the layer fits beside two halo tiles they are loaded once per CTA and stay resident (64-channel layers: 72 KB
  // instead of 73 KB of L2->SMEM traffic per tile); otherwise they stream through a ring of up to 12 stages.  What is
  // left deepens the halo ring (thin, HBM-bound 1x1 layers need > 2 tiles in flight to cover the memory latency).
  // Opt-in experiment: on B200 the staged TMA store measured no faster than 256-bit per-thread stores (the epilogue was
  // instruction-latency bound, not L1-wavefront bound) and costs 32 KB of shared memory.
  static const bool use_tma_store = getenv("SPYR_CONV_TMA_STORE") != nullptr;
  p.tma_store = (d->y_f32 == nullptr && (d->y_raw != nullptr || d->y_act != nullptr) && use_tma_store && !spyr_split()) ? 1 : 0;
  if (p.tma_store) {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t strides[3] = {(uint64_t)d->Cout * 2, (uint64_t)d->W * d->Cout * 2, (uint64_t)d->H * d->W * d->Cout * 2};
    uint32_t box[4] = {32, 8, 4, 1};
    const void* y0 = d->y_raw != nullptr ? d->y_raw : d->y_act;
    const void* y1 = d->y_act != nullptr ? d->y_act : d->y_raw;
    SPYR_REQUIRE(((uintptr_t)y0 & 15) == 0 && ((uintptr_t)y1 & 15) == 0, "conv2d_fprop: unaligned output pointer");
    if (spyr_tmap_encode(&maps.y[0], y0, 4, dims, strides, box, 2)) return 3;
    if (spyr_tmap_encode(&maps.y[1], y1, 4, dims, strides, box, 2)) return 3;
  } else {
    maps.y[0] = maps.x[0];
    maps.y[1] = maps.x[0];
  }
  const int usable = 218 * 1024 - 2 * 11 * bn * 4 - (p.tma_store ? EPI_STAGE_TOTAL : 0);
  int stages_per_tile = 0;
  for (int s = 0; s < d->nsrc; ++s) stages_per_tile += p.kchunks[s] * (p.border[s] ? 9 : 1);
  int stages;
  p.b_resident = (p.n_tiles == 1 && stages_per_tile <= 24 && d->stages == 0 &&
                  stages_per_tile * p.b_stage_bytes + 2 * p.a_buf_bytes <= usable) ? 1 : 0;
  for (int s = 0; s < d->nsrc; ++s)
    if (p.wpi[s]) p.b_resident = 0;  // per-image weights change with the tile
  if (p.b_resident) {
    stages = stages_per_tile;
  } else {
    stages = (usable - 18 * 1024 - 2 * p.a_buf_bytes) / p.b_stage_bytes;
    if (stages > 12) stages = 12;
    if (d->stages > 0 && d->stages < stages) stages = d->stages;
    if (stages < 2) return -1;
  }
  p.b_stages = stages;
  p.a_bufs = (usable - stages * p.b_stage_bytes) / p.a_buf_bytes;
  if (p.a_bufs > 4) p.a_bufs = 4;
  if (p.a_bufs < 2) p.a_bufs = 2;
  p.bias = d->bias; p.bias2 = d->bias2; p.bias3 = d->bias3;
  p.stencil_mask = d->stencil_mask; p.stencil_w = d->stencil_w;
  p.dmask = (const bf16*)d->dmask; p.dmask_slope = d->dmask_slope;
  p.residual = (const bf16*)d->residual;
  p.y_raw = (bf16*)d->y_raw; p.y_act = (bf16*)d->y_act;
  p.act = d->act; p.act_slope = d->act_slope;
  p.y_f32 = d->y_f32;
  p.split = (spyr_split() && d->y_f32 == nullptr) ? 1 : 0;
  p.y_plane = (long long)d->B * (d->pool ? (d->H / 2) * (d->W / 2) : d->H * d->W) * d->Cout;
  p.res_plane = p.y_plane;
  p.acc_scale = 1.f;
  if (p.split) {
    int hh_steps = 0;  // (tap, chunk) stages of the hi*hi sources (the last third, see spyr_conv2d_fprop), 4 MMAs each
    for (int s = 2 * (d->nsrc / 3); s < d->nsrc; ++s) hh_steps += p.kchunks[s] * (p.border[s] ? 9 : 1);
    p.acc_scale = 1.f + (float)(4 * hh_steps) * 2.9802322e-8f;
  }
  SPYR_REQUIRE(!p.split || (!d->pool && !d->residual_pooled),
               "conv2d_fprop: the pooled epilogue / pooled residual are not available in split-BF16 mode");
  p.epi_mode = epi_mode_for(p);
  p.res_pooled = d->residual_pooled ? 1 : 0;
  if (p.res_pooled)
    SPYR_REQUIRE(p.epi_mode != 0 && p.residual != nullptr && !d->pool && (d->H % 2) == 0 && (d->W % 2) == 0,
                 "conv2d_fprop: residual_pooled needs a residual, Cout %% 32 == 0 and BF16 outputs (Cout=%d)", d->Cout);
  p.pool = d->pool ? 1 : 0;
  if (p.pool) {
    SPYR_REQUIRE(p.epi_mode != 0 && p.dmask == nullptr && (d->H % 2) == 0 && (d->W % 2) == 0,
                 "conv2d_fprop: pool needs Cout %% 32 == 0, BF16 outputs and no gate / stencil (Cout=%d)", d->Cout);
  }
  const size_t smem_bytes = (size_t)p.a_bufs * p.a_buf_bytes + (size_t)stages * p.b_stage_bytes +
                            (size_t)(p.tma_store ? EPI_STAGE_TOTAL : 0) +
                            (2 * p.a_bufs + 2 * stages + 4) * 8 + 16 + (size_t)2 * 11 * bn * 4 + 1024;
  SPYR_REQUIRE(smem_bytes <= 227 * 1024, "conv2d_fprop: shared-memory plan of %zu bytes exceeds 227 KB", smem_bytes);
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SPYR_CHECK_CUDA(cudaGetDevice(&dev));
    SPYR_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (getenv("SPYR_CONV_SMS") != nullptr && atoi(getenv("SPYR_CONV_SMS")) > 0) num_sms = atoi(getenv("SPYR_CONV_SMS"));  // experiment
  }
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  void (*kernel)(HaloMaps, HaloParams) = conv_halo_kernel<true, 0>;
  const int epi_index = epi_static_index(p);
  if (!p.split) { SPYR_EPI_SWITCH(epi_index, kernel = conv_halo_kernel<false, kEpi>) }
  static bool configured[2][EPI_VARIANTS] = {};
  if (!configured[p.split ? 1 : 0][epi_index]) {
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[p.split ? 1 : 0][epi_index] = true;
  }
  kernel<<<grid, THREADS, smem_bytes, stream>>>(maps, p);
  spyr_note_kernel(1);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
