// CTA-pair (tcgen05 cta_group::2) variant of the halo-tiled convolution for layers with >= 128 output channels.
//
// Why: a single-CTA M=128 instruction reads 4 KB of activations plus the whole N x 16 weight slice from shared memory,
// and every CTA streams the full weight slice of each (tap, chunk) from L2.  With cta_group::2 two SMs execute one
// M=256 instruction: each SM stages and reads its own 128 activation rows but only HALF of the weight slice.
// Measured (profiles/r01_*pair*): TMA traffic of a 256->256 @64x64 launch drops from 796 MB to 437 MB and the shared-
// memory operand wavefronts per SM from 96 to 64 per instruction, while the time per instruction stays ~235 clk
// (ideal 128) -- i.e. on this shape the pair kernel buys L2 / shared-memory head-room, not yet tensor-pipe time.
//
// Structure = conv_halo.cu with the pair protocol:
//   * cluster of 2 CTAs; CTA r owns pixel tile 2*pair + r (its own halo buffers, its own TMEM rows) and weight rows
//     [r*N/2, (r+1)*N/2) of every (tap, chunk) slice;
//   * both CTAs issue TMA with .cta_group::2 so the bytes are counted on the LEADER's (rank 0) full barriers;
//   * only the leader issues tcgen05.mma.cta_group::2 (M = 256) and commits with .multicast::cluster to the empty /
//     accumulator-full barriers of BOTH CTAs;
//   * the epilogue warps of both CTAs drain their own TMEM half and arrive on the leader's accumulator-empty barrier
//     (CTA 1 through mapa + mbarrier.arrive.shared::cluster).
#include "common.cuh"
#include "../../include/spyramid_b200.h"
#include <cstdlib>

#include "conv_halo_common.cuh"

extern void spyr_count_launch();
void spyr_note_kernel(int id);

namespace {
using namespace halo;

constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // shared::cluster address -> same offset in the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}

template <bool SPLIT, int EPI>
__global__ void __launch_bounds__(THREADS, 1)
conv_halo2_kernel(const __grid_constant__ HaloMaps maps, const HaloParams p) {
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
  float* epi_const = reinterpret_cast<float*>(tmem_holder + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int half_n = p.block_n >> 1;  // weight rows (output channels) this CTA stages per (tap, chunk)

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
      mbar_init(&acc_empty[i], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  cluster_sync_all();  // both CTAs' barriers exist before anything can signal them remotely
  if (warp == 2) {
    tmem_alloc2(tmem_holder, p.tmem_cols);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int th_rows = 16 * p.msub;
  const int total_pairs = (p.m_tiles >> 1) * p.n_tiles;

  if (warp == 0) {
    {
      // ===== TMA producer (both CTAs): own halo tiles, own half of every weight slice; warp-uniform coordinates,
      // one elected lane issues =====
      const bool issue = elect_one();
      int abuf = 0, bst = 0;
      uint32_t aph = 0, bph = 0;
      HaloTileIter ti;
      ti.init(pair, p.m_tiles >> 1);
      for (; ti.nt < p.n_tiles; ti.next(num_pairs, p.m_tiles >> 1)) {
        int w0, h0, n0;
        halo_tile_origin(p, 2 * ti.m + (int)rank, th_rows, w0, h0, n0);
        const int n_off = ti.nt * p.block_n + (int)rank * half_n;
        for (int s = 0; s < p.nsrc; ++s) {
          const int bd = p.border[s];
          const int taps = bd ? 9 : 1;
          for (int c = 0; c < p.kchunks[s]; ++c) {
            mbar_wait(&a_empty[abuf], aph ^ 1);
            if (issue) {
              if (leader) mbar_arrive_expect_tx(&a_full[abuf], (uint32_t)(2 * p.a_rows[s] * 128));
              tma2_load_4d(a_ring + abuf * p.a_buf_bytes, &maps.x[s], &a_full[abuf], c * KC, w0 - bd, h0 - bd, n0);
            }
            if (++abuf == A_BUFS) {
              abuf = 0;
              aph ^= 1;
            }
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(&b_empty[bst], bph ^ 1);
              uint8_t* b_dst = b_ring + bst * p.b_stage_bytes;
              if (issue) {
                if (leader) mbar_arrive_expect_tx(&b_full[bst], (uint32_t)(p.block_n * 128));
                if (p.wmn[s]) {
                  const int wtap = p.wpi[s] ? n0 : (bd ? 8 - tap : 0);
                  for (int j = 0; j < (half_n >> 6); ++j)
                    tma2_load_3d(b_dst + j * 8192, &maps.w[s], &b_full[bst], n_off + j * 64, c * KC, wtap);
                } else {
                  tma2_load_3d(b_dst, &maps.w[s], &b_full[bst], c * KC, n_off, p.wpi[s] ? n0 : tap);
                }
              }
              if (++bst == p.b_stages) {
                bst = 0;
                bph ^= 1;
              }
            }
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (leader) {
      // ===== MMA issuer: the leader CTA's warp 1 drives both SMs.  Warp-uniform operands, one elected lane issues
      // (see conv_halo.cu: a single-lane loop costs an R2UR/ELECT waterfall per MMA) =====
      const bool issue = elect_one();
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);
      const uint32_t idesc_k = umma_idesc_bf16(256, p.block_n, 0, 0);
      const uint32_t idesc_mn = umma_idesc_bf16(256, p.block_n, 0, 1);
      const uint64_t desc_base = ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint32_t a_ring_lo = (smem_u32(a_ring) & 0x3FFFF) >> 4;
      const uint32_t b_ring_lo = (smem_u32(b_ring) & 0x3FFFF) >> 4;
      const uint32_t a_buf16 = (uint32_t)p.a_buf_bytes >> 4, b_stage16 = (uint32_t)p.b_stage_bytes >> 4;
      int abuf = 0, bst = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (int t = pair; t < total_pairs; t += num_pairs, ++it) {
        const int buf = it & 1;
        const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&acc_empty[buf], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * p.msub * p.bn_cols);
        uint32_t accum = 0;
        for (int s = 0; s < p.nsrc; ++s) {
          const int bd = p.border[s];
          const int taps = bd ? 9 : 1;
          const uint32_t pitch16 = (uint32_t)(8 + 2 * bd) * 8u;
          const uint32_t sub16 = 16u * pitch16;
          const bool mn = p.wmn[s] != 0;
          const uint64_t a_hi = desc_base | ((uint64_t)pitch16 << 32);
          const uint64_t b_hi = mn ? (desc_base | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32))
                                   : (desc_base | ((uint64_t)(1024 >> 4) << 32));
          const uint32_t bk16 = mn ? (2048u >> 4) : 2u;
          const uint32_t idesc = mn ? idesc_mn : idesc_k;
          for (int c = 0; c < p.kchunks[s]; ++c) {
            mbar_wait(&a_full[abuf], aph);
            const uint32_t a_lo = a_ring_lo + (uint32_t)abuf * a_buf16;
            uint32_t row0 = 0;
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(&b_full[bst], bph);
              tc_fence_after();
              const uint64_t db0 = b_hi | (uint64_t)(b_ring_lo + (uint32_t)bst * b_stage16);
              const uint64_t da0 = a_hi | (uint64_t)(a_lo + row0 * 8u);
              if (issue) {
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  umma2_bf16(acc0, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(bk16 * k), idesc, k == 0 ? accum : 1u);
                if (p.msub == 2) {
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k)
                    umma2_bf16(acc0 + (uint32_t)p.bn_cols, da0 + (uint64_t)(sub16 + 2 * k), db0 + (uint64_t)(bk16 * k),
                               idesc, k == 0 ? accum : 1u);
                }
                umma2_commit_mc(&b_empty[bst]);
              }
              accum = 1;
              if (++bst == p.b_stages) {
                bst = 0;
                bph ^= 1;
              }
              row0 += ((tap % 3) == 2) ? 8u : 1u;
            }
            if (issue) umma2_commit_mc(&a_empty[abuf]);
            if (++abuf == A_BUFS) {
              abuf = 0;
              aph ^= 1;
            }
          }
        }
        if (issue) umma2_commit_mc(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs, own TMEM rows = own pixel tile, all block_n columns) =====
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int m = q * 32 + lane;
    const int et = threadIdx.x - 128;
    int it = 0;
    int staged_n_off = -1, cbuf = 1;
    int sbuf = 0;
    EpiStore es;
    es.maps = nullptr;  // per-thread stores (TMA-store epilogue: conv_halo.cu only so far)
    es.stage = 0;
    es.lane = lane;
    es.sbuf = &sbuf;
    es.w = es.h = es.n = 0;
    HaloTileIter ti;
    ti.init(pair, p.m_tiles >> 1);
    for (; ti.nt < p.n_tiles; ti.next(num_pairs, p.m_tiles >> 1), ++it) {
      const int buf = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      int w0, h0, n;
      halo_tile_origin(p, 2 * ti.m + (int)rank, th_rows, w0, h0, n);
      const int n_off = ti.nt * p.block_n;
      if (n_off != staged_n_off) {
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
            for (int tt = 0; tt < 10; ++tt)
              dst[(1 + tt) * p.block_n + c] = col < p.Cout ? __ldg(&p.stencil_w[tt * p.Cout + col]) : 0.f;
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
          for (int tt = 0; tt < 9; ++tt) {
            const int hh = h + tt / 3 - 1, ww = w + tt % 3 - 1;
            float v = 0.f;
            if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) v = __ldg(&p.stencil_mask[((size_t)n * p.H + hh) * p.W + ww]);
            mk[sub][tt] = v;
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
        for (int c0 = half * 32; c0 < p.block_n; c0 += 64) {
          uint32_t r[32];
          tmem_ld32(acc + (uint32_t)c0, r);
          tmem_ld_wait();
          epilogue_static<SPLIT, EPI>(p, r, pix, n_off + c0, c0, ec, mk[sub], mk_mode[sub], es);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&acc_empty[buf], 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading its TMEM half / our barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, p.tmem_cols);
  }
}

uint32_t pow2_at_least2(int n, uint32_t lo) {
  uint32_t c = lo;
  while ((int)c < n) c <<= 1;
  return c;
}

}  // namespace

// Returns 0 on launch, -1 when not eligible (the caller falls back to conv_halo.cu), >0 on error.
int spyr_conv_halo2_launch(const spyr_conv_desc* d, cudaStream_t stream) {
  if (d->H < 16 || d->W < 8 || (d->H % 16) != 0 || (d->W % 8) != 0) return -1;
  if (d->splits > 1 || d->block_n != 0) return -1;
  if (d->y_f32 != nullptr && !d->f32_store) return -1;
  // Cout == 64 with K-major weights also pairs (each CTA stages 32 of the 64 weight rows: 59 instead of 75 clk of operand
  // fetch per MMA).  Its MN-major (input-gradient) form would need 64-byte-wide operand rows: single-CTA kernel.
  bool narrow = d->Cout == 64 && getenv("SPYR_PAIR_NO_N64") == nullptr;
  bool has3x3 = false;
  for (int s = 0; s < d->nsrc; ++s) {
    if (d->src[s].w_mn_major) narrow = false;
    if (d->src[s].ksize == 3) has3x3 = true;
  }
  if (!has3x3) narrow = false;  // thin 1x1 layers are epilogue-bound: the single-CTA kernel's deeper halo ring wins
  if (!narrow && (d->Cout < 128 || (d->Cout % 128) != 0)) return -1;
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cout = d->Cout;
  p.nsrc = d->nsrc;
  static const bool bn128 = getenv("SPYR_PAIR_BN128") != nullptr;  // experiment: less weight traffic per FLOP
  int bn = narrow ? 64 : ((d->Cout % 256 == 0 && !(bn128 && d->H % 32 == 0)) ? 256 : 128);  // whole N blocks only
  p.msub = (d->H % 32 == 0 && bn <= 128) ? 2 : 1;
  static const bool no_wave_model = getenv("SPYR_PAIR_NO_WAVE_MODEL") != nullptr;
  if (!narrow && !no_wave_model && (bn == 256 || p.msub == 2)) {
    // Wave quantisation on the 16x16 / 32x32 maps: 20 x 32x32 pixels x 256 channels are 80 pair tiles of 128 x 256 on 74
    // pair slots -- two waves, the second 8 % full.  Half-size work items (128 pixels x 128 channels per CTA) make it 160
    // items = three half-length waves.  Cost = waves x (item size + a fixed per-item overhead); the smaller item is taken
    // when it wins by >= 10 % (on large maps the bigger item's lower operand traffic per FLOP is worth more).
    static int slots = 0;
    if (slots == 0) {
      int dev = 0, sms = 0;
      SPYR_CHECK_CUDA(cudaGetDevice(&dev));
      SPYR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      slots = sms / 2;
    }
    const long long mt_big = (long long)(d->W / 8) * (d->H / (16 * p.msub)) * d->B;
    const long long mt_small = (long long)(d->W / 8) * (d->H / 16) * d->B;
    if ((mt_big & 1) == 0 && (mt_small & 1) == 0) {
      const long long items_big = (mt_big / 2) * (d->Cout / bn), items_small = (mt_small / 2) * (d->Cout / 128);
      const long long cost_big = ((items_big + slots - 1) / slots) * (p.msub * bn + 24);
      const long long cost_small = ((items_small + slots - 1) / slots) * (128 + 24);
      if (cost_small * 10 <= cost_big * 9) {
        bn = 128;
        p.msub = 1;
      }
    }
  }
  p.block_n = bn;
  p.bn_cols = bn;
  p.tmem_cols = pow2_at_least2(2 * p.msub * p.bn_cols, 32);
  if (p.tmem_cols > 512) return -1;
  p.tiles_w = d->W / 8;
  p.tiles_h = d->H / (16 * p.msub);
  if ((p.tiles_w & (p.tiles_w - 1)) != 0 || (p.tiles_h & (p.tiles_h - 1)) != 0) return -1;  // shift / mask tile decode
  for (p.lw = 0; (1 << p.lw) < p.tiles_w; ++p.lw) {}
  for (p.lh = 0; (1 << p.lh) < p.tiles_h; ++p.lh) {}
  p.m_tiles = p.tiles_w * p.tiles_h * d->B;
  if (p.m_tiles & 1) return -1;
  p.n_tiles = d->Cout / bn;
  p.total_tiles = p.m_tiles * p.n_tiles;
  HaloMaps maps;
  memset(&maps, 0, sizeof(maps));
  int max_rows = 0;
  const int half_n = bn / 2;
  p.b_stage_bytes = half_n * 128;
  for (int s = 0; s < d->nsrc; ++s) {
    const spyr_conv_src& src = d->src[s];
    SPYR_REQUIRE(src.ksize == 1 || src.ksize == 3, "conv2d_fprop: ksize must be 1 or 3");
    SPYR_REQUIRE(src.cin % 8 == 0 && src.cin > 0, "conv2d_fprop: cin=%d must be a multiple of 8", src.cin);
    SPYR_REQUIRE(((uintptr_t)src.x & 15) == 0 && ((uintptr_t)src.w & 15) == 0, "conv2d_fprop: unaligned pointer");
    SPYR_REQUIRE(!src.w_per_image || src.ksize == 1, "conv2d_fprop: per-image weights need ksize 1");
    const int bd = src.ksize == 3 ? 1 : 0;
    p.border[s] = bd;
    p.kchunks[s] = ceil_div(src.cin, KC);
    p.wmn[s] = src.w_mn_major ? 1 : 0;
    p.wpi[s] = src.w_per_image ? 1 : 0;
    const int bw = 8 + 2 * bd, bh = 16 * p.msub + 2 * bd;
    p.a_rows[s] = bw * bh;
    if (p.a_rows[s] > max_rows) max_rows = p.a_rows[s];
    if (src.w_mn_major && (half_n / 64) * 8192 > p.b_stage_bytes) p.b_stage_bytes = (half_n / 64) * 8192;
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
      uint32_t box[3] = {KC, (uint32_t)half_n, 1};
      if (spyr_tmap_encode(&maps.w[s], src.w, 3, dims, strides, box, 1)) return 3;
    }
  }
  for (int s = d->nsrc; s < SPYR_CONV_MAX_SRC; ++s) {
    maps.x[s] = maps.x[0];
    maps.w[s] = maps.w[0];
  }
  p.a_buf_bytes = ceil_div(max_rows * 128, 1024) * 1024;
  const int budget = 196 * 1024 - A_BUFS * p.a_buf_bytes;
  int stages = budget / p.b_stage_bytes;
  if (stages > 12) stages = 12;
  if (stages < 2) return -1;
  p.b_stages = stages;
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
  const size_t smem_bytes = (size_t)A_BUFS * p.a_buf_bytes + (size_t)stages * p.b_stage_bytes +
                            (2 * A_BUFS + 2 * stages + 4) * 8 + 16 + (size_t)2 * 11 * bn * 4 + 1024;
  void (*kernel)(HaloMaps, HaloParams) = conv_halo2_kernel<true, 0>;
  const int epi_index = epi_static_index(p);
  if (!p.split) { SPYR_EPI_SWITCH(epi_index, kernel = conv_halo2_kernel<false, kEpi>) }
  static bool configured[2][EPI_VARIANTS] = {};
  if (!configured[p.split ? 1 : 0][epi_index]) {
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[p.split ? 1 : 0][epi_index] = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SPYR_CHECK_CUDA(cudaGetDevice(&dev));
    SPYR_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (getenv("SPYR_CONV_SMS") != nullptr && atoi(getenv("SPYR_CONV_SMS")) > 0) num_sms = atoi(getenv("SPYR_CONV_SMS"));  // experiment
  }
  const int total_pairs = (p.m_tiles / 2) * p.n_tiles;
  int pairs = num_sms / 2;
  if (pairs > total_pairs) pairs = total_pairs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SPYR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, maps, p));
  spyr_note_kernel(0);
  spyr_count_launch();
  return 0;
}
