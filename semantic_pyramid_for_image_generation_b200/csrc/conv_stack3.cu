// 3x3 convolution with 64 output channels: the three taps of one kernel ROW share an activation window and are stacked
// along N, so one tcgen05.mma is M=128 x N=192 instead of M=128 x N=64.
//
// Why: tcgen05.mma reads both operands from shared memory and the 4 KB activation slice (A) of every M=128 instruction
// costs ~43 clk whatever N is (tests/native/mma_probe.cu: 75 / 107 / 171 clk for N = 64 / 128 / 256).  A 64-channel layer
// on the halo kernels issues nine N=64 instructions per 16 input channels, 9 x 75 = 675 clk, and sits on that floor
// (profiles/r02_conv_profile_d.txt: 739 TFLOP/s for 64->64 @ 256x256, the largest row of the step).  Cout = 64 leaves no
// wider N -- unless several taps share an A window.  For a fixed kernel row dy, the three taps (dy, -1), (dy, 0), (dy, +1)
// multiply DIFFERENT weights with the SAME input pixel when the products are indexed by the input pixel instead of the
// output pixel:
//
//     D[m][dx*64 + co] = sum_{dy, ci} X[row(m) + dy, col(m)][ci] * W[dy][dx][co][ci]          (3 x 4 MMAs per 64 ci)
//     out[r][c][co]    = D[(r, c-1)][0*64 + co] + D[(r, c)][1*64 + co] + D[(r, c+1)][2*64 + co]
//
// i.e. three N=192 instructions per 16 input channels, 3 x (43 + 96) = 417 clk, and the epilogue adds each thread's
// middle block to its left neighbour's first block and its right neighbour's last block.  With the tile laid out as
// 4 image rows x 32 window columns, TMEM lane = window column inside a warp's 32-lane quarter, so the neighbours are
// lanes l-1 and l+1 of the same warp: two warp shuffles per output value, no shared memory.  Window columns 0 and 31
// have no left / right neighbour: a tile produces 30 output columns from a 32-column window (tiles step 30 columns; the
// columns past the image are TMA zero fill and are not stored).
//
//   tile          4 rows x 30 columns of one image x 64 output channels; window 6 rows x 32 columns per 64-channel chunk
//   A operand     one TMA box (64 ch, 32, 6, 1) per chunk, SWIZZLE_128B; kernel row dy = descriptor start + dy * 32 rows
//                 (4 KB, so every start is 1 KB aligned and the 8-row groups are the standard 1 KB apart)
//   B operand     one TMA box (64 ch, 64 co, 3 taps) per (chunk, dy): the tap-major K-major weights [9][64][Cin] hold
//                 the three taps of a kernel row back to back, so the box lands as 192 contiguous 128-byte rows
//   accumulators  2 x 192 TMEM columns (epilogue of tile i overlaps the MMAs of tile i+1)
//   warps         0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..11 = epilogue (4 lane quarters = 4 image rows,
//                 x 2 halves of the 64 output channels)
//
// Same sources / epilogue contract as spyr_conv2d_fprop (include/spyramid_b200.h); eligibility in
// spyr_conv_stack3_launch.  Input gradients of 64-channel layers arrive here through the K-major, tap-flipped re-lay of
// the forward weights (spyr_weight_transpose_flip), exactly like on the pair kernel.
#include "common.cuh"
#include "../../include/spyramid_b200.h"
#include <cstdlib>

#include "conv_halo_common.cuh"

extern void spyr_count_launch();
void spyr_note_kernel(int id);

namespace {
using namespace halo;

// Developer switches (SPYR_S3_DBG, wrong results): 1 = no global part of the epilogue, 2 = no side-block loads / shuffles,
// 4 = no MMAs -- the A/B timings in profiles/r02_ncu_conv_stack3.txt.
constexpr int WIN_W = 32;             // window columns = lanes of one TMEM quarter
constexpr int OUT_W = WIN_W - 2;      // output columns per tile
constexpr int ROWS = 4;               // output rows per tile = TMEM lane quarters
constexpr int WIN_H = ROWS + 2;
constexpr int COUT = 64;
constexpr int N_MMA = 3 * COUT;       // three taps of a kernel row
constexpr int ACC_STRIDE = 256;       // TMEM columns between the two accumulator sets
constexpr int A_BYTES = WIN_H * WIN_W * 128;
constexpr int B_BYTES = N_MMA * 128;

// One epilogue per instantiation.  EPI = HaloParams::epi_mode (0: generic epilogue_chunk, else 1 + (OUT - 1) * 4 + DMASK * 2 +
// RES for epilogue_chunk_fast): the 13-way runtime switch of epilogue_dispatch put every variant's registers and code in
// one kernel (168 registers with spills, ~12 700 SASS instructions).
constexpr int IN_BYTES = OUT_W * 64;  // one staged operand row block: 30 pixels x 32 channels

// gate / residual / bias / activation of one 32-channel chunk -> packed BF16 outputs (oraw, oact)
template <int EPI>
__device__ __forceinline__ void stack3_epilogue(const HaloParams& p, const uint32_t* r, size_t pix, int c0, const EpiConst& ec,
                                                const EpiStore& es, uint32_t in_stage, int tma_in, uint32_t* oraw,
                                                uint32_t* oact) {
  constexpr int OUT = (EPI - 1) / 4 + 1;
  constexpr bool DMASK = ((EPI - 1) & 2) != 0;
  constexpr bool RES = ((EPI - 1) & 1) != 0;
  const size_t off0 = pix * COUT + c0;
  uint32_t dm[16], rs[16];
  float rsc = 1.f;
  // staged rows use the layout of the store staging: row = es.lane (64 bytes), 16-byte chunk index ^= (row >> 1) & 3
  const uint32_t srow = in_stage + (uint32_t)es.lane * 64u;
  const uint32_t sx = ((uint32_t)es.lane >> 1) & 3u;
  if (DMASK) {
    if (tma_in & 1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) lds128u(srow + (((uint32_t)c ^ sx) << 4), dm + 4 * c);
    } else {
      ldg256(p.dmask + off0, dm);
      ldg256(p.dmask + off0 + 16, dm + 8);
    }
  }
  if (RES) {
    if (tma_in & 2) {
#pragma unroll
      for (int c = 0; c < 4; ++c) lds128u(srow + 2048u + (((uint32_t)c ^ sx) << 4), rs + 4 * c);
    } else {
      size_t ro = off0;
      if (p.res_pooled) {
        const int w = (int)(pix % (size_t)p.W);
        const size_t row = pix / (size_t)p.W;
        ro = ((row >> 1) * (size_t)(p.W >> 1) + (size_t)(w >> 1)) * COUT + c0;
        rsc = 0.25f;
      }
      ldg256(p.residual + ro, rs);
      ldg256(p.residual + ro + 16, rs + 8);
    }
  }
  const float sl = (p.act == 1) ? 0.f : ((p.act == 2) ? p.act_slope : 1.f);
  const float gs = p.dmask_slope;
  const uint32_t bias_saddr = smem_u32(ec.bias + c0);
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 b = lds128f(bias_saddr + q4 * 16);
    float v[4] = {__uint_as_float(r[4 * q4]) + b.x, __uint_as_float(r[4 * q4 + 1]) + b.y,
                  __uint_as_float(r[4 * q4 + 2]) + b.z, __uint_as_float(r[4 * q4 + 3]) + b.w};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float a = v[2 * h], c = v[2 * h + 1];
      if (DMASK) {
        const float2 f = unpack_bf16x2(dm[2 * q4 + h]);
        a = f.x > 0.f ? a : a * gs;
        c = f.y > 0.f ? c : c * gs;
      }
      if (RES) {
        const float2 f = unpack_bf16x2(rs[2 * q4 + h]);
        a += rsc * f.x;
        c += rsc * f.y;
      }
      if (OUT & 1) oraw[2 * q4 + h] = pack_bf16x2(a, c);
      if (OUT & 2) {
        a = a > 0.f ? a : a * sl;
        c = c > 0.f ? c : c * sl;
        oact[2 * q4 + h] = pack_bf16x2(a, c);
      }
    }
  }
}

template <int EPI>
__device__ __forceinline__ void stack3_store(const HaloParams& p, size_t pix, int c0, const EpiStore& es, const uint32_t* oraw,
                                             const uint32_t* oact) {
  constexpr int OUT = (EPI - 1) / 4 + 1;
  const size_t off0 = pix * COUT + c0;
  if (es.maps != nullptr) {
    if (OUT & 1) epi_tma_store(es, &es.maps[0], c0, oraw);
    if (OUT & 2) epi_tma_store(es, &es.maps[1], c0, oact);
  } else {
    if (OUT & 1) {
      stg256(p.y_raw + off0, oraw);
      stg256(p.y_raw + off0 + 16, oraw + 8);
    }
    if (OUT & 2) {
      stg256(p.y_act + off0, oact);
      stg256(p.y_act + off0 + 16, oact + 8);
    }
  }
}

template <bool SPLIT, int EPI>
__global__ void __launch_bounds__(THREADS, 1)
conv_stack3_kernel(const __grid_constant__ HaloMaps maps, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + p.a_bufs * A_BYTES;
  uint8_t* epi_stage = b_ring + p.b_stages * B_BYTES;  // EPI_STAGE_TOTAL bytes when p.tma_store
  uint8_t* epi_in = epi_stage + (p.tma_store ? EPI_STAGE_TOTAL : 0);  // [8 warps][gate, residual][2 KB] when p.tma_in
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_in + (p.tma_in ? EPI_STAGE_TOTAL : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.a_bufs;
  uint64_t* b_full = a_empty + p.a_bufs;
  uint64_t* b_empty = b_full + p.b_stages;
  uint64_t* acc_full = b_empty + p.b_stages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* in_full = acc_empty + 2;  // one per epilogue warp
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(in_full + EPI_WARPS);
  float* epi_const = reinterpret_cast<float*>(tmem_holder + 4);  // [11][64]: bias sum + 10 stencil rows

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
    if (p.tma_in) {
      tma_prefetch_desc(&maps.g[0]);
      tma_prefetch_desc(&maps.g[1]);
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
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&in_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, p.tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 4) {
    // the layer's epilogue constants (one N block: staged once)
    const int et = threadIdx.x - 128;
    for (int c = et; c < COUT; c += EPI_WARPS * 32) {
      float b = 0.f;
      if (p.bias != nullptr) b += __ldg(&p.bias[c]);
      if (p.bias2 != nullptr) b += __ldg(&p.bias2[c]);
      if (p.bias3 != nullptr) b += __ldg(&p.bias3[c]);
      epi_const[c] = b;
      if (p.stencil_w != nullptr) {
#pragma unroll
        for (int t = 0; t < 10; ++t) epi_const[(1 + t) * COUT + c] = __ldg(&p.stencil_w[t * COUT + c]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===== TMA producer: warp-uniform coordinates, one elected lane issues =====
    const bool issue = elect_one();
    int abuf = 0, bst = 0;
    uint32_t aph = 0, bph = 0;
    bool load_b = true;  // resident weights: only the first tile of this CTA loads them
    TileIter ti;
    ti.init(p, blockIdx.x);
    for (; ti.n < p.B; ti.next(p)) {
      const int w0 = ti.tw * OUT_W;
      const int h0 = ti.th * ROWS;
      const int n0 = ti.n;
      for (int s = 0; s < p.nsrc; ++s) {
        for (int c = 0; c < p.kchunks[s]; ++c) {
          mbar_wait(&a_empty[abuf], aph ^ 1);
          if (issue) {
            mbar_arrive_expect_tx(&a_full[abuf], (uint32_t)A_BYTES);
            tma_load_4d(a_ring + abuf * A_BYTES, &maps.x[s], &a_full[abuf], c * KC, w0 - 1, h0 - 1, n0);
          }
          if (++abuf == p.a_bufs) {
            abuf = 0;
            aph ^= 1;
          }
          for (int dy = 0; load_b && dy < 3; ++dy) {
            mbar_wait(&b_empty[bst], bph ^ 1);
            if (issue) {
              mbar_arrive_expect_tx(&b_full[bst], (uint32_t)B_BYTES);
              tma_load_3d(b_ring + bst * B_BYTES, &maps.w[s], &b_full[bst], c * KC, 0, dy * 3);
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
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, only the tcgen05 instructions are predicated on one elected lane =====
    const bool leader = elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, *tmem_holder, 0);
    const uint32_t idesc = umma_idesc_bf16(128, N_MMA, 0, 0);
    // version 1, SWIZZLE_128B, 8-row groups 1 KB apart (both operands are dense 128-byte rows)
    const uint64_t desc_hi = ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)(1024 >> 4) << 32);
    const uint32_t a_ring_lo = (smem_u32(a_ring) & 0x3FFFF) >> 4;
    const uint32_t b_ring_lo = (smem_u32(b_ring) & 0x3FFFF) >> 4;
    int abuf = 0, bst = 0;
    uint32_t aph = 0, bph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&acc_empty[buf], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t acc0 = tb + (uint32_t)(buf * ACC_STRIDE);
      uint32_t accum = 0;
      for (int s = 0; s < p.nsrc; ++s) {
        for (int c = 0; c < p.kchunks[s]; ++c) {
          mbar_wait(&a_full[abuf], aph);
          const uint32_t a_lo = a_ring_lo + (uint32_t)abuf * (uint32_t)(A_BYTES >> 4);
          for (int dy = 0; dy < 3; ++dy) {
            if (!p.b_resident)
              mbar_wait(&b_full[bst], bph);
            else if (it == 0)
              mbar_wait(&b_full[bst], 0);
            tc_fence_after();
            const uint64_t db0 = desc_hi | (uint64_t)(b_ring_lo + (uint32_t)bst * (uint32_t)(B_BYTES >> 4));
            const uint64_t da0 = desc_hi | (uint64_t)(a_lo + (uint32_t)dy * (uint32_t)(WIN_W * 128 >> 4));
            if (leader) {
              if (!(p.dbg & 4)) {
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  umma_bf16(acc0, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, k == 0 ? accum : 1u);
              }
              if (!p.b_resident) umma_commit(&b_empty[bst]);
            }
            accum = 1;
            if (++bst == p.b_stages) {
              bst = 0;
              bph ^= 1;
            }
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
  } else if (warp >= 4) {
    // ===== epilogue: warp (q, half) owns image row q of the tile and output channels [32*half, 32*half + 32) =====
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int c0 = half * 32;
    int sbuf = 0;
    // Stores: a thread owns one pixel's 32 channels = 64 bytes, its neighbours' pixels are 128 bytes away, so a per-thread
    // store instruction touches 32 different lines (~53 L1 wavefronts each; measured 36 us of a 118 us launch).  The
    // fast epilogues instead stage the warp's 30 x 64-byte rows in shared memory (row = lane - 1; the two window-edge
    // lanes land in rows 30 / 31, outside the box) and one TMA store writes the (32 ch, 30 px) box; columns past the
    // image are clipped by the TMA unit.
    EpiStore es;
    es.maps = p.tma_store ? maps.y : nullptr;
    es.stage = smem_u32(epi_stage) + (uint32_t)((warp - 4) * 2 * EPI_STAGE_BYTES);
    es.lane = p.tma_store ? ((lane + 31) & 31) : lane;
    es.sbuf = &sbuf;
    es.w = es.h = es.n = 0;
    EpiConst ec;
    ec.bias = epi_const;
    ec.stencil = epi_const + COUT;
    // Gate / residual operands: per-thread loads would again touch 32 lines per instruction (+45 us on a gated 64 -> 64
    // input gradient).  Each warp instead fetches its (32 ch, 30 px) block of the NEXT tile with one TMA load per operand
    // right after it has consumed the current one, so the data is in shared memory long before that tile's accumulator.
    const uint32_t in_stage = smem_u32(epi_in) + (uint32_t)((warp - 4) * 2 * EPI_STAGE_BYTES);
    uint64_t* in_bar = &in_full[warp - 4];
    const uint32_t in_bytes = (uint32_t)(((p.tma_in & 1) + ((p.tma_in >> 1) & 1)) * IN_BYTES);
    uint32_t in_ph = 0;
    int it = 0;
    TileIter ti;
    ti.init(p, blockIdx.x);
    if (p.tma_in && es.lane == 0 && ti.n < p.B) {
      mbar_arrive_expect_tx(in_bar, in_bytes);
      if (p.tma_in & 1) tma_load_4d(epi_in + (warp - 4) * 2 * EPI_STAGE_BYTES, &maps.g[0], in_bar, c0, ti.tw * OUT_W, ti.th * ROWS + q, ti.n);
      if (p.tma_in & 2) tma_load_4d(epi_in + (warp - 4) * 2 * EPI_STAGE_BYTES + 2048, &maps.g[1], in_bar, c0, ti.tw * OUT_W, ti.th * ROWS + q, ti.n);
    }
    for (; ti.n < p.B; ti.next(p), ++it) {
      const int buf = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      const int w0 = ti.tw * OUT_W;
      const int h = ti.th * ROWS + q;
      const int n = ti.n;
      const int w = w0 - 1 + lane;  // image column of this lane's window column
      const bool valid = lane >= 1 && lane <= OUT_W && w < p.W;
      const int wc = w < 0 ? 0 : (w >= p.W ? p.W - 1 : w);  // edge lanes: any readable pixel of the row
      const size_t pix = ((size_t)n * p.H + h) * p.W + wc;
      es.w = w0;
      es.h = h;
      es.n = n;
      float mk[9];
      int mk_mode = 0;
      if (EPI == 0 && p.stencil_mask != nullptr && valid) {
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
      if (valid && !p.tma_in && (p.dmask != nullptr || (p.residual != nullptr && !p.res_pooled))) {
        // pull the gate / residual operands of this tile into L2 while its MMAs still run
        const size_t off = pix * COUT + c0;
        if (p.dmask != nullptr) prefetch_l2(p.dmask + off);
        if (p.residual != nullptr && !p.res_pooled) prefetch_l2(p.residual + off);
      }
      mbar_wait(&acc_full[buf], acc_ph);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * ACC_STRIDE) + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      uint32_t r[32], t[32];
      tmem_ld32(acc + COUT, r);  // middle tap: this lane's own output column
      tmem_ld32(acc, t);         // left tap (dx = -1): belongs to the output one column to the right
      tmem_ld_wait();
      if (!(p.dbg & 2)) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = __float_as_uint(__uint_as_float(r[j]) + __shfl_up_sync(0xffffffffu, __uint_as_float(t[j]), 1));
        tmem_ld32(acc + 2 * COUT, t);  // right tap (dx = +1): belongs to the output one column to the left
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = __float_as_uint(__uint_as_float(r[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(t[j]), 1));
      }
      // the accumulator set is in registers: hand it back before the global-memory part of the epilogue
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (EPI == 0) {
        if (valid && !(p.dbg & 1)) epilogue_chunk<SPLIT>(p, r, pix, c0, c0, ec, mk, mk_mode, es);
      } else if ((valid || p.tma_store) && !(p.dbg & 1)) {
        constexpr int E = EPI == 0 ? 1 : EPI;
        uint32_t oraw[16], oact[16];
        if (p.tma_in) {
          mbar_wait(in_bar, in_ph);
          in_ph ^= 1;
        }
        stack3_epilogue<E>(p, r, pix, c0, ec, es, in_stage, p.tma_in, oraw, oact);
        if (p.tma_in) {
          // the staged operands are consumed (their values went into oraw / oact): fetch the next tile's
          TileIter nx = ti;
          nx.next(p);
          __syncwarp();
          if (es.lane == 0 && nx.n < p.B) {
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(in_bar, in_bytes);
            uint8_t* dst = epi_in + (warp - 4) * 2 * EPI_STAGE_BYTES;
            if (p.tma_in & 1) tma_load_4d(dst, &maps.g[0], in_bar, c0, nx.tw * OUT_W, nx.th * ROWS + q, nx.n);
            if (p.tma_in & 2) tma_load_4d(dst + 2048, &maps.g[1], in_bar, c0, nx.tw * OUT_W, nx.th * ROWS + q, nx.n);
          }
        }
        stack3_store<E>(p, pix, c0, es, oraw, oact);
      }
      if ((p.dbg & 1) && r[0] == 0x12345678u) p.y_raw[0] = __float2bfloat16(1.f);  // keeps the loads alive when the epilogue is off
    }
    if (p.tma_store && es.lane == 0) bulk_wait_all();  // staged rows must be read (and written out) before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace

// Returns 0 on launch, -1 when the problem is not eligible (the caller goes on to the pair / halo kernels), >0 on error.
int spyr_conv_stack3_launch(const spyr_conv_desc* d, cudaStream_t stream) {
  static const bool off = getenv("SPYR_CONV_NO_STACK3") != nullptr;
  if (off) return -1;
  if (d->Cout != COUT || d->W < 128 || (d->H % ROWS) != 0 || d->H < ROWS) return -1;
  if (d->splits > 1 || d->block_n != 0 || d->stages != 0 || d->pool) return -1;
  if (d->y_f32 != nullptr && !d->f32_store) return -1;
  for (int s = 0; s < d->nsrc; ++s) {
    const spyr_conv_src& src = d->src[s];
    if (src.ksize != 3 || src.w_mn_major || src.w_per_image) return -1;
    if (src.cin % 8 != 0 || src.cin <= 0) return -1;
    if (((uintptr_t)src.x & 15) != 0 || ((uintptr_t)src.w & 15) != 0) return -1;
  }
  {
    // more than one 64-channel chunk per tile: the MMA time per tile doubles while the epilogue stays, the pair kernel
    // closes most of the gap, and on 128-wide maps (5 column tiles of 30 for 128 columns) it is ahead (55 vs 61 us)
    int chunks = 0;
    for (int s = 0; s < d->nsrc; ++s) chunks += ceil_div(d->src[s].cin, KC);
    if (chunks > 1 && d->W < 256 && !spyr_split()) return -1;
  }
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cout = COUT;
  p.nsrc = d->nsrc;
  p.msub = 1;
  p.block_n = COUT;
  p.bn_cols = ACC_STRIDE;
  p.tmem_cols = 512;
  p.tiles_w = ceil_div(d->W, OUT_W);
  p.tiles_h = d->H / ROWS;
  p.m_tiles = p.tiles_w * p.tiles_h * d->B;
  p.n_tiles = 1;
  p.total_tiles = p.m_tiles;
  HaloMaps maps;
  memset(&maps, 0, sizeof(maps));
  int stages_per_tile = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    const spyr_conv_src& src = d->src[s];
    p.border[s] = 1;
    p.kchunks[s] = ceil_div(src.cin, KC);
    p.a_rows[s] = WIN_H * WIN_W;
    stages_per_tile += 3 * p.kchunks[s];
    {
      uint64_t dims[4] = {(uint64_t)src.cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
      uint64_t strides[3] = {(uint64_t)src.cin * 2, (uint64_t)d->W * src.cin * 2, (uint64_t)d->H * d->W * src.cin * 2};
      uint32_t box[4] = {KC, WIN_W, WIN_H, 1};
      if (spyr_tmap_encode(&maps.x[s], src.x, 4, dims, strides, box, 1)) return 3;
    }
    {
      uint64_t dims[3] = {(uint64_t)src.cin, (uint64_t)COUT, 9};
      uint64_t strides[2] = {(uint64_t)src.cin * 2, (uint64_t)COUT * src.cin * 2};
      uint32_t box[3] = {KC, COUT, 3};
      if (spyr_tmap_encode(&maps.w[s], src.w, 3, dims, strides, box, 1)) return 3;
    }
  }
  for (int s = d->nsrc; s < SPYR_CONV_MAX_SRC; ++s) {
    maps.x[s] = maps.x[0];
    maps.w[s] = maps.w[0];
  }
  p.bias = d->bias; p.bias2 = d->bias2; p.bias3 = d->bias3;
  p.stencil_mask = d->stencil_mask; p.stencil_w = d->stencil_w;
  p.dmask = (const bf16*)d->dmask; p.dmask_slope = d->dmask_slope;
  p.residual = (const bf16*)d->residual;
  p.y_raw = (bf16*)d->y_raw; p.y_act = (bf16*)d->y_act;
  p.act = d->act; p.act_slope = d->act_slope;
  p.y_f32 = d->y_f32;
  p.split = (spyr_split() && d->y_f32 == nullptr) ? 1 : 0;
  if (p.split && d->residual_pooled) return -1;
  p.epi_mode = epi_mode_for(p);
  p.res_pooled = d->residual_pooled ? 1 : 0;
  if (p.res_pooled && (p.epi_mode == 0 || p.residual == nullptr)) return -1;
  // TMA-store epilogue: the compile-time specialised epilogues only (BF16 outputs, every lane runs the same code)
  // Shared-memory plan.  Resident weights come first (streaming them costs 24 KB per kernel row and chunk for every
  // 120-pixel tile), then the staging of the TMA-store epilogue, then the staging of the gate / residual loads.
  const int base_usable = 218 * 1024 - 11 * COUT * 4;
  const int resident_need = stages_per_tile * B_BYTES + 2 * A_BYTES;
  const bool can_resident = resident_need <= base_usable;
  static const bool no_tma_store = getenv("SPYR_S3_NO_TMA_STORE") != nullptr;
  p.tma_store = (p.epi_mode != 0 && !no_tma_store && (!can_resident || resident_need + EPI_STAGE_TOTAL <= base_usable)) ? 1 : 0;
  if (p.tma_store) {
    uint64_t dims[4] = {(uint64_t)COUT, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t strides[3] = {(uint64_t)COUT * 2, (uint64_t)d->W * COUT * 2, (uint64_t)d->H * d->W * COUT * 2};
    uint32_t box[4] = {32, OUT_W, 1, 1};
    const void* y0 = d->y_raw != nullptr ? d->y_raw : d->y_act;
    const void* y1 = d->y_act != nullptr ? d->y_act : d->y_raw;
    if (((uintptr_t)y0 & 15) != 0 || ((uintptr_t)y1 & 15) != 0) return -1;
    if (spyr_tmap_encode(&maps.y[0], y0, 4, dims, strides, box, 2)) return 3;
    if (spyr_tmap_encode(&maps.y[1], y1, 4, dims, strides, box, 2)) return 3;
  } else {
    maps.y[0] = maps.x[0];
    maps.y[1] = maps.x[0];
  }
  maps.g[0] = maps.x[0];
  maps.g[1] = maps.x[0];
  static const bool no_tma_in = getenv("SPYR_S3_NO_TMA_IN") != nullptr;
  if (p.tma_store && !no_tma_in && (!can_resident || resident_need + 2 * EPI_STAGE_TOTAL <= base_usable)) {
    uint64_t dims[4] = {(uint64_t)COUT, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t strides[3] = {(uint64_t)COUT * 2, (uint64_t)d->W * COUT * 2, (uint64_t)d->H * d->W * COUT * 2};
    uint32_t box[4] = {32, OUT_W, 1, 1};
    if (p.dmask != nullptr && ((uintptr_t)p.dmask & 15) == 0) {
      if (spyr_tmap_encode(&maps.g[0], p.dmask, 4, dims, strides, box, 2)) return 3;
      p.tma_in |= 1;
    }
    if (p.residual != nullptr && !p.res_pooled && ((uintptr_t)p.residual & 15) == 0) {
      if (spyr_tmap_encode(&maps.g[1], p.residual, 4, dims, strides, box, 2)) return 3;
      p.tma_in |= 2;
    }
  }
  p.a_buf_bytes = A_BYTES;
  p.b_stage_bytes = B_BYTES;
  const int usable = base_usable - (p.tma_store ? EPI_STAGE_TOTAL : 0) - (p.tma_in ? EPI_STAGE_TOTAL : 0);
  p.b_resident = (stages_per_tile * B_BYTES + 2 * A_BYTES <= usable) ? 1 : 0;
  int stages;
  if (p.b_resident) {
    stages = stages_per_tile;
  } else {
    stages = (usable - 3 * A_BYTES) / B_BYTES;
    if (stages > 6) stages = 6;
  }
  p.b_stages = stages;
  p.a_bufs = (usable - stages * B_BYTES) / A_BYTES;
  if (p.a_bufs > 4) p.a_bufs = 4;
  if (p.a_bufs < 2 || stages < 2) return -1;
  p.y_plane = (long long)d->B * d->H * d->W * COUT;
  p.res_plane = p.y_plane;
  p.acc_scale = 1.f;
  if (p.split) {
    // every output is the sum of three accumulators, each fed by a third of the hi*hi instructions (4 per chunk and
    // kernel row): the truncation shrink of the tensor pipe's accumulator is a third of the one-accumulator kernels'
    int hh_steps = 0;
    for (int s = 2 * (d->nsrc / 3); s < d->nsrc; ++s) hh_steps += 3 * p.kchunks[s];
    p.acc_scale = 1.f + (float)(4 * hh_steps) * 2.9802322e-8f;
  }
  p.dbg = getenv("SPYR_S3_DBG") ? atoi(getenv("SPYR_S3_DBG")) : 0;
  const size_t smem_bytes = (size_t)p.a_bufs * A_BYTES + (size_t)stages * B_BYTES + (size_t)(p.tma_store ? EPI_STAGE_TOTAL : 0) +
                            (size_t)(p.tma_in ? EPI_STAGE_TOTAL : 0) + (2 * p.a_bufs + 2 * stages + 4 + EPI_WARPS) * 8 + 16 +
                            (size_t)11 * COUT * 4 + 1024;
  if (smem_bytes > 227 * 1024) return -1;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SPYR_CHECK_CUDA(cudaGetDevice(&dev));
    SPYR_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (getenv("SPYR_CONV_SMS") != nullptr && atoi(getenv("SPYR_CONV_SMS")) > 0) num_sms = atoi(getenv("SPYR_CONV_SMS"));  // experiment
  }
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  p.step_w = grid % p.tiles_w;
  p.step_h = (grid / p.tiles_w) % p.tiles_h;
  p.step_n = grid / (p.tiles_w * p.tiles_h);
  void (*kernel)(HaloMaps, HaloParams) = nullptr;
  if (p.split) {
    kernel = conv_stack3_kernel<true, 0>;
  } else {
    switch (p.epi_mode) {
      case 0: kernel = conv_stack3_kernel<false, 0>; break;
      case 1: kernel = conv_stack3_kernel<false, 1>; break;
      case 2: kernel = conv_stack3_kernel<false, 2>; break;
      case 3: kernel = conv_stack3_kernel<false, 3>; break;
      case 4: kernel = conv_stack3_kernel<false, 4>; break;
      case 5: kernel = conv_stack3_kernel<false, 5>; break;
      case 6: kernel = conv_stack3_kernel<false, 6>; break;
      case 7: kernel = conv_stack3_kernel<false, 7>; break;
      case 8: kernel = conv_stack3_kernel<false, 8>; break;
      case 9: kernel = conv_stack3_kernel<false, 9>; break;
      case 10: kernel = conv_stack3_kernel<false, 10>; break;
      case 11: kernel = conv_stack3_kernel<false, 11>; break;
      default: kernel = conv_stack3_kernel<false, 12>; break;
    }
  }
  static bool configured[2][13] = {};
  bool& conf = configured[p.split ? 1 : 0][p.split ? 0 : p.epi_mode];
  if (!conf) {
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    conf = true;
  }
  kernel<<<grid, THREADS, smem_bytes, stream>>>(maps, p);
  spyr_note_kernel(5);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
