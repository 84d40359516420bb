// Shared by conv_halo.cu (one CTA per tile) and conv_halo2.cu (CTA pairs, cta_group::2): launch parameters and the fused
// epilogue of the halo-tiled convolution kernels.
#pragma once
#include "common.cuh"

namespace halo {

constexpr int KC = 64;
constexpr int THREADS = 384;
constexpr int EPI_WARPS = 8;
constexpr int A_BUFS = 2;

struct HaloParams {
  int B, H, W, Cout;
  int msub;
  int tiles_w, tiles_h, m_tiles, n_tiles, total_tiles;
  int nsrc;
  int border[SPYR_CONV_MAX_SRC];   // 1: 3x3 conv, 0: 1x1
  int kchunks[SPYR_CONV_MAX_SRC];
  int wmn[SPYR_CONV_MAX_SRC];
  int wpi[SPYR_CONV_MAX_SRC];
  int a_rows[SPYR_CONV_MAX_SRC];   // rows (pixels) of the A box of this source
  int block_n, bn_cols;  // bn_cols: TMEM column stride of one accumulator (power of two >= 32)
  int a_buf_bytes, b_stage_bytes, b_stages;
  int a_bufs;      // halo-tile ring depth (conv_halo.cu; the pair kernel uses A_BUFS)
  int epi_mode;    // 0: generic epilogue_chunk, else a specialised epilogue_chunk_fast (epi_mode_for)
  int pool;        // 2x2 average pool across lanes in the epilogue; outputs / residual live at (H/2, W/2)
  int res_pooled;  // residual lives at (H/2, W/2) and is added as 0.25 * residual[h/2][w/2]
  int tma_store;   // epilogue writes y_raw / y_act through shared-memory staging + TMA stores (maps.y)
  int b_resident;  // conv_halo.cu: every (source, chunk, tap) weight slice of the layer stays in shared memory
  int dbg;         // developer switches of conv_stack3.cu (SPYR_S3_DBG), 0 in production
  int tma_in;      // conv_stack3.cu: bit 0 = gate, bit 1 = residual arrive through TMA loads into shared memory (maps.g)
  int lw, lh;      // log2(tiles_w), log2(tiles_h) of the halo kernels (H, W are powers of two there)
  int step_w, step_h, step_n;  // gridDim.x decomposed over (tiles_w, tiles_h, images): TileIter advances without divisions
  int split;       // split-BF16 mode: y_raw / y_act / residual are hi + lo plane pairs (generic epilogue only)
  long long y_plane, res_plane;  // elements from the hi to the lo plane of the outputs / of the residual
  float acc_scale;  // split mode: 1 + (hi*hi accumulations per output) * 2^-25 -- the tensor pipe's FP32 accumulator truncates
                    // on every tcgen05.mma; over a long reduction that is a predictable shrink (tools/probe_split_accum.py)
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
  CUtensorMap x[SPYR_CONV_MAX_SRC];
  CUtensorMap w[SPYR_CONV_MAX_SRC];
  CUtensorMap y[2];  // y_raw, y_act as (C, W, H, B) with a (32 ch, 8, 4, 1) box, SWIZZLE_64B: epilogue TMA stores
  CUtensorMap g[2];  // conv_stack3.cu: gate (dmask) and residual maps, same box as its y maps: epilogue TMA loads
};

// Coordinates of the tiles a persistent CTA visits (tile = blockIdx.x, += gridDim.x), kept as (column tile, row tile,
// image) and advanced with two compare-and-subtract carries.  The three runtime divisions of tile -> (tw, th, n) cost
// ~130 dependent instructions (IABS / I2F / MUFU.RCP / IMAD.HI chains) per tile in EVERY role; the epilogue warps, two per
// scheduler and latency-bound, spent a quarter of their per-tile time on them (profiles/r02_ncu_conv_stack3.txt).
struct TileIter {
  int tw, th, n;
  __device__ __forceinline__ void init(const HaloParams& p, int tile) {
    tw = tile % p.tiles_w;
    const int r = tile / p.tiles_w;
    th = r % p.tiles_h;
    n = r / p.tiles_h;
  }
  __device__ __forceinline__ void next(const HaloParams& p) {
    tw += p.step_w;
    const int cw = tw >= p.tiles_w ? 1 : 0;
    tw -= cw ? p.tiles_w : 0;
    th += p.step_h + cw;
    const int ch = th >= p.tiles_h ? 1 : 0;
    th -= ch ? p.tiles_h : 0;
    n += p.step_n + ch;
  }
};

// conv_halo.cu / conv_halo2.cu: persistent loop over work items t = n_tile * M + m (t = first, += step) without the five
// runtime divisions of t -> (n_tile, m) -> (column tile, row tile, image): (m, n_tile) advance by compare-and-subtract,
// the pixel tile decodes with shifts because tiles_w and tiles_h are powers of two.
struct HaloTileIter {
  int m, nt;
  __device__ __forceinline__ void init(int first, int M) {
    nt = first / M;
    m = first - nt * M;
  }
  __device__ __forceinline__ void next(int step, int M) {
    m += step;
    while (m >= M) {
      m -= M;
      ++nt;
    }
  }
};
__device__ __forceinline__ void halo_tile_origin(const HaloParams& p, int m_tile, int th_rows, int& w0, int& h0, int& n0) {
  w0 = (m_tile & (p.tiles_w - 1)) * 8;
  h0 = ((m_tile >> p.lw) & (p.tiles_h - 1)) * th_rows;
  n0 = m_tile >> (p.lw + p.lh);
}

constexpr int EPI_STAGE_BYTES = 2048;  // one warp's 32 pixels x 32 channels BF16
constexpr int EPI_STAGE_TOTAL = EPI_WARPS * 2 * EPI_STAGE_BYTES;

// Where one warp's 32x32 chunk goes when the epilogue stores through TMA: the warp's two staging buffers (1 KB aligned)
// and the box coordinates.  A per-thread STG of this chunk touches 32 different 128-byte lines per request
// (~53 L1 data-pipe wavefronts, ncu) and competes with the tensor pipe's operand fetch for the same pipe; staging the
// chunk in shared memory (4 conflict-free STS.128 per thread) and letting the TMA engine write it costs ~4x fewer.
struct EpiStore {
  const CUtensorMap* maps;  // nullptr: per-thread global stores
  uint32_t stage;           // shared address of this warp's staging buffers (2 x EPI_STAGE_BYTES)
  int w, h, n;              // box origin (pixels) of this warp's 8 x 4 pixel block
  int lane;
  int* sbuf;                // which of the two buffers the next store uses
};

__device__ __forceinline__ void epi_tma_store(const EpiStore& es, const CUtensorMap* map, int col0, const uint32_t* o) {
  const int b = *es.sbuf;
  *es.sbuf = b ^ 1;
  if (es.lane == 0) bulk_wait_read<1>();  // the store issued two chunks ago has finished reading this buffer
  __syncwarp();
  // row = lane (64 bytes), SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3
  const uint32_t row = es.stage + (uint32_t)b * EPI_STAGE_BYTES + (uint32_t)es.lane * 64u;
  const uint32_t x = ((uint32_t)es.lane >> 1) & 3u;
#pragma unroll
  for (int c = 0; c < 4; ++c) sts128(row + (((uint32_t)c ^ x) << 4), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
  fence_proxy_async_smem();
  __syncwarp();
  if (es.lane == 0) {
    tma_store_4d(map, es.stage + (uint32_t)b * EPI_STAGE_BYTES, col0, es.w, es.h, es.n);
    bulk_commit();
  }
}

// Per-tile epilogue constants staged in shared memory by the epilogue warps: the summed bias vectors of this N block and
// the ten FP32 stencil rows of the mask channel.  Every lane of a warp reads the same address (broadcast).
struct EpiConst {
  const float* bias;     // [block_n]
  const float* stencil;  // [10][block_n] or nullptr
};

// SPLIT is a compile-time copy of p.split: the single-plane instantiation carries none of the lo-plane code (as a runtime
// branch it pushed both convolution kernels over their 168-register budget: 150 bytes of spills, +16 % kernel time)
template <bool SPLIT>
__device__ __forceinline__ void epilogue_chunk(const HaloParams& p, const uint32_t* r, size_t pix, int col0, int c0,
                                               const EpiConst& ec, const float* mk, int mk_mode, const EpiStore& es) {
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
  const bool full = col0 + 32 <= p.Cout;
  const bool wide = full && (p.Cout & 15) == 0;  // 32-byte aligned 16-channel groups -> 256-bit LDG/STG
  // issue every global read of this 32-channel chunk before any arithmetic (read-only path, independent of the stores)
  uint32_t dm[16], rs[16], rl[16];
  if (SPLIT && p.residual != nullptr) {
    // lo plane of the residual (the gate below only needs signs: hi plane)
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (col0 + g * 8 + 8 <= p.Cout) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.residual + p.res_plane + off0 + g * 8));
        rl[4 * g] = u.x; rl[4 * g + 1] = u.y; rl[4 * g + 2] = u.z; rl[4 * g + 3] = u.w;
      }
  }
  if (p.dmask != nullptr) {
    if (wide) {
      ldg256(p.dmask + off0, dm);
      ldg256(p.dmask + off0 + 16, dm + 8);
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (col0 + g * 8 + 8 <= p.Cout) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.dmask + off0 + g * 8));
          dm[4 * g] = u.x; dm[4 * g + 1] = u.y; dm[4 * g + 2] = u.z; dm[4 * g + 3] = u.w;
        }
    }
  }
  if (p.residual != nullptr) {
    if (wide) {
      ldg256(p.residual + off0, rs);
      ldg256(p.residual + off0 + 16, rs + 8);
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (col0 + g * 8 + 8 <= p.Cout) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.residual + off0 + g * 8));
          rs[4 * g] = u.x; rs[4 * g + 1] = u.y; rs[4 * g + 2] = u.z; rs[4 * g + 3] = u.w;
        }
    }
  }
  const float sl = (p.act == 1) ? 0.f : ((p.act == 2) ? p.act_slope : 1.f);
  uint32_t oraw[16], oact[16];
  uint32_t lraw[16], lact[16];  // lo planes (split mode)
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    if (col + 8 > p.Cout) break;
    float v[8];
    const float4 b0 = *reinterpret_cast<const float4*>(ec.bias + c0 + g * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(ec.bias + c0 + g * 8 + 4);
    const float as = SPLIT ? p.acc_scale : 1.f;
    v[0] = __uint_as_float(r[g * 8 + 0]) * as + b0.x;
    v[1] = __uint_as_float(r[g * 8 + 1]) * as + b0.y;
    v[2] = __uint_as_float(r[g * 8 + 2]) * as + b0.z;
    v[3] = __uint_as_float(r[g * 8 + 3]) * as + b0.w;
    v[4] = __uint_as_float(r[g * 8 + 4]) * as + b1.x;
    v[5] = __uint_as_float(r[g * 8 + 5]) * as + b1.y;
    v[6] = __uint_as_float(r[g * 8 + 6]) * as + b1.z;
    v[7] = __uint_as_float(r[g * 8 + 7]) * as + b1.w;
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
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(dm[4 * g + j]);
        if (!(f.x > 0.f)) v[2 * j] *= p.dmask_slope;
        if (!(f.y > 0.f)) v[2 * j + 1] *= p.dmask_slope;
      }
    }
    if (p.residual != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(rs[4 * g + j]);
        v[2 * j] += f.x;
        v[2 * j + 1] += f.y;
      }
      if (SPLIT) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(rl[4 * g + j]);
          v[2 * j] += f.x;
          v[2 * j + 1] += f.y;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) oraw[4 * g + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    if (SPLIT) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 h = unpack_bf16x2(oraw[4 * g + j]);
        lraw[4 * g + j] = pack_bf16x2(v[2 * j] - h.x, v[2 * j + 1] - h.y);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * sl;
#pragma unroll
    for (int j = 0; j < 4; ++j) oact[4 * g + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    if (SPLIT) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 h = unpack_bf16x2(oact[4 * g + j]);
        lact[4 * g + j] = pack_bf16x2(v[2 * j] - h.x, v[2 * j + 1] - h.y);
      }
    }
  }
  if (SPLIT) {
    // lo planes: per-thread 16-byte stores (the strict mode is bound by its 3x tensor work, not by this epilogue)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (col0 + g * 8 + 8 > p.Cout) break;
      if (p.y_raw != nullptr)
        *reinterpret_cast<uint4*>(p.y_raw + p.y_plane + off0 + g * 8) = make_uint4(lraw[4 * g], lraw[4 * g + 1], lraw[4 * g + 2], lraw[4 * g + 3]);
      if (p.y_act != nullptr)
        *reinterpret_cast<uint4*>(p.y_act + p.y_plane + off0 + g * 8) = make_uint4(lact[4 * g], lact[4 * g + 1], lact[4 * g + 2], lact[4 * g + 3]);
    }
  }
  if (es.maps != nullptr) {
    // channels past Cout (ragged last chunk) hold garbage in o*: the TMA store clips them
    if (p.y_raw != nullptr) epi_tma_store(es, &es.maps[0], col0, oraw);
    if (p.y_act != nullptr) epi_tma_store(es, &es.maps[1], col0, oact);
  } else if (wide) {
    if (p.y_raw != nullptr) {
      stg256(p.y_raw + off0, oraw);
      stg256(p.y_raw + off0 + 16, oraw + 8);
    }
    if (p.y_act != nullptr) {
      stg256(p.y_act + off0, oact);
      stg256(p.y_act + off0 + 16, oact + 8);
    }
  } else {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (col0 + g * 8 + 8 > p.Cout) break;
      if (p.y_raw != nullptr)
        *reinterpret_cast<uint4*>(p.y_raw + off0 + g * 8) = make_uint4(oraw[4 * g], oraw[4 * g + 1], oraw[4 * g + 2], oraw[4 * g + 3]);
      if (p.y_act != nullptr)
        *reinterpret_cast<uint4*>(p.y_act + off0 + g * 8) = make_uint4(oact[4 * g], oact[4 * g + 1], oact[4 * g + 2], oact[4 * g + 3]);
    }
  }
}


// ---- specialised epilogue: compile-time feature set, no branches inside the 32-channel chunk ----
// The generic epilogue_chunk above tests every optional feature per 8-channel group; ncu counted ~375 warp instructions
// per chunk (1.3 branches per FADD) executed serially by each of the 8 epilogue warps, which made the epilogue -- not the
// tensor pipe -- the critical path of the 64-channel layers (7000 clk per 256-pixel tile against 5400 clk of MMAs).
// OUT: 1 = y_raw, 2 = y_act, 3 = both.  Requires Cout % 32 == 0 (every chunk full, 32-byte aligned) and no stencil.
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ void lds128u(uint32_t addr, uint32_t* v) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}

template <int OUT, bool DMASK, bool RES>
__device__ __forceinline__ void epilogue_chunk_fast(const HaloParams& p, const uint32_t* r, size_t off0, int col0,
                                                    uint32_t bias_saddr, const EpiStore& es, size_t res_off, float res_scale) {
  uint32_t dm[16], rs[16];
  if (DMASK) {
    ldg256(p.dmask + off0, dm);
    ldg256(p.dmask + off0 + 16, dm + 8);
  }
  if (RES) {
    ldg256(p.residual + res_off, rs);
    ldg256(p.residual + res_off + 16, rs + 8);
  }
  const float sl = (p.act == 1) ? 0.f : ((p.act == 2) ? p.act_slope : 1.f);
  const float gs = p.dmask_slope;
  uint32_t oraw[16], oact[16];
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
        a += res_scale * f.x;
        c += res_scale * f.y;
      }
      if (OUT & 1) oraw[2 * q4 + h] = pack_bf16x2(a, c);
      if (OUT & 2) {
        a = a > 0.f ? a : a * sl;
        c = c > 0.f ? c : c * sl;
        oact[2 * q4 + h] = pack_bf16x2(a, c);
      }
    }
  }
  if (es.maps != nullptr) {
    if (OUT & 1) epi_tma_store(es, &es.maps[0], col0, oraw);
    if (OUT & 2) epi_tma_store(es, &es.maps[1], col0, oact);
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

// Pooled variant: thread <-> pixel (w = lane & 7, h = lane >> 3 within the warp's 8x4 block), so the 2x2 window of an
// even (w, h) is lanes {l, l^1, l^8, l^9}: two butterfly adds per value, then the eight even-even lanes hold the pooled
// pixels of the block and write 64 bytes each.  Saves the full-resolution write and the separate pooling pass.
template <int OUT, bool RES>
__device__ __forceinline__ void epilogue_chunk_pool(const HaloParams& p, const uint32_t* r, size_t pooled_off, int lane,
                                                    uint32_t bias_saddr) {
  const float sl = (p.act == 1) ? 0.f : ((p.act == 2) ? p.act_slope : 1.f);
  float v[32];
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 b = lds128f(bias_saddr + q4 * 16);
    v[4 * q4] = __uint_as_float(r[4 * q4]) + b.x;
    v[4 * q4 + 1] = __uint_as_float(r[4 * q4 + 1]) + b.y;
    v[4 * q4 + 2] = __uint_as_float(r[4 * q4 + 2]) + b.z;
    v[4 * q4 + 3] = __uint_as_float(r[4 * q4 + 3]) + b.w;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    v[j] += __shfl_xor_sync(0xffffffffu, v[j], 1);
    v[j] += __shfl_xor_sync(0xffffffffu, v[j], 8);
    v[j] *= 0.25f;
  }
  if ((lane & 9) != 0) return;
  uint32_t rs[16], oraw[16], oact[16];
  if (RES) {
    ldg256(p.residual + pooled_off, rs);
    ldg256(p.residual + pooled_off + 16, rs + 8);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float a = v[2 * j], c = v[2 * j + 1];
    if (RES) {
      const float2 f = unpack_bf16x2(rs[j]);
      a += f.x;
      c += f.y;
    }
    if (OUT & 1) oraw[j] = pack_bf16x2(a, c);
    if (OUT & 2) oact[j] = pack_bf16x2(a > 0.f ? a : a * sl, c > 0.f ? c : c * sl);
  }
  if (OUT & 1) {
    stg256(p.y_raw + pooled_off, oraw);
    stg256(p.y_raw + pooled_off + 16, oraw + 8);
  }
  if (OUT & 2) {
    stg256(p.y_act + pooled_off, oact);
    stg256(p.y_act + pooled_off + 16, oact + 8);
  }
}

// epi_mode: 0 = generic; otherwise 1 + (OUT - 1) * 4 + DMASK * 2 + RES
__host__ inline int epi_mode_for(const HaloParams& p) {
  if (p.split) return 0;  // split-BF16 outputs: generic epilogue only
  if (p.y_f32 != nullptr || p.stencil_mask != nullptr || p.stencil_w != nullptr || (p.Cout & 31) != 0) return 0;
  const int out = (p.y_raw != nullptr ? 1 : 0) | (p.y_act != nullptr ? 2 : 0);
  if (out == 0) return 0;
  return 1 + (out - 1) * 4 + (p.dmask != nullptr ? 2 : 0) + (p.residual != nullptr ? 1 : 0);
}

template <bool SPLIT>
__device__ __forceinline__ void epilogue_dispatch(const HaloParams& p, const uint32_t* r, size_t pix, int col0, int c0,
                                                  const EpiConst& ec, const float* mk, int mk_mode, const EpiStore& es) {
  if (SPLIT) {  // split-BF16 outputs: generic epilogue only (epi_mode_for returns 0)
    epilogue_chunk<true>(p, r, pix, col0, c0, ec, mk, mk_mode, es);
    return;
  }
  const uint32_t bs = smem_u32(ec.bias + c0);
  if (p.pool) {
    // pix = (n*H + h)*W + w of this lane's pixel -> pooled pixel (n, h/2, w/2); only even-even lanes use the offset
    const int w = (int)(pix % (size_t)p.W);
    const size_t row = pix / (size_t)p.W;  // n*H + h
    const size_t poff = ((row >> 1) * (size_t)(p.W >> 1) + (size_t)(w >> 1)) * p.Cout + col0;
    switch (p.epi_mode) {
      case 1: epilogue_chunk_pool<1, false>(p, r, poff, es.lane, bs); break;
      case 2: epilogue_chunk_pool<1, true>(p, r, poff, es.lane, bs); break;
      case 5: epilogue_chunk_pool<2, false>(p, r, poff, es.lane, bs); break;
      case 6: epilogue_chunk_pool<2, true>(p, r, poff, es.lane, bs); break;
      case 9: epilogue_chunk_pool<3, false>(p, r, poff, es.lane, bs); break;
      default: epilogue_chunk_pool<3, true>(p, r, poff, es.lane, bs); break;
    }
    return;
  }
  const size_t off0 = pix * p.Cout + col0;
  size_t ro = off0;
  float rsc = 1.f;
  if (p.res_pooled) {
    const int w = (int)(pix % (size_t)p.W);
    const size_t row = pix / (size_t)p.W;
    ro = ((row >> 1) * (size_t)(p.W >> 1) + (size_t)(w >> 1)) * p.Cout + col0;
    rsc = 0.25f;
  }
  switch (p.epi_mode) {
    case 1: epilogue_chunk_fast<1, false, false>(p, r, off0, col0, bs, es, ro, rsc); break;
    case 2: epilogue_chunk_fast<1, false, true>(p, r, off0, col0, bs, es, ro, rsc); break;
    case 3: epilogue_chunk_fast<1, true, false>(p, r, off0, col0, bs, es, ro, rsc); break;
    case 4: epilogue_chunk_fast<1, true, true>(p, r, off0, col0, bs, es, ro, rsc); break;
    case 5: epilogue_chunk_fast<2, false, false>(p, r, off0, col0, bs, es, ro, rsc); break;
    case 6: epilogue_chunk_fast<2, false, true>(p, r, off0, col0, bs, es, ro, rsc); break;
    case 7: epilogue_chunk_fast<2, true, false>(p, r, off0, col0, bs, es, ro, rsc); break;
    case 8: epilogue_chunk_fast<2, true, true>(p, r, off0, col0, bs, es, ro, rsc); break;
    case 9: epilogue_chunk_fast<3, false, false>(p, r, off0, col0, bs, es, ro, rsc); break;
    case 10: epilogue_chunk_fast<3, false, true>(p, r, off0, col0, bs, es, ro, rsc); break;
    case 11: epilogue_chunk_fast<3, true, false>(p, r, off0, col0, bs, es, ro, rsc); break;
    case 12: epilogue_chunk_fast<3, true, true>(p, r, off0, col0, bs, es, ro, rsc); break;
    default: epilogue_chunk<false>(p, r, pix, col0, c0, ec, mk, mk_mode, es); break;
  }
}

// ---- compile-time form of epilogue_dispatch ----
// EPI: 0 = generic epilogue_chunk; 1..12 = HaloParams::epi_mode (epilogue_chunk_fast<OUT, DMASK, RES>); 13..18 = the pooled
// epilogue, 13 + (OUT - 1) * 2 + RES.  One epilogue per kernel instantiation: the 19-way runtime switch kept every variant's
// registers and code in one kernel (168 registers with spills; conv_stack3.cu measured 486 -> 336 instructions per tile
// and chunk when its switch and its index divisions went away).
constexpr int EPI_VARIANTS = 19;
__host__ inline int epi_static_index(const HaloParams& p) {
  if (p.split || p.epi_mode == 0) return 0;
  if (p.pool) return 13 + ((p.epi_mode - 1) / 4) * 2 + ((p.epi_mode - 1) & 1);
  return p.epi_mode;
}

template <bool SPLIT, int EPI>
__device__ __forceinline__ void epilogue_static(const HaloParams& p, const uint32_t* r, size_t pix, int col0, int c0,
                                                const EpiConst& ec, const float* mk, int mk_mode, const EpiStore& es) {
  if (SPLIT || EPI == 0) {
    epilogue_chunk<SPLIT>(p, r, pix, col0, c0, ec, mk, mk_mode, es);
    return;
  }
  const uint32_t bs = smem_u32(ec.bias + c0);
  if (EPI >= 13) {
    constexpr int OUT = (EPI >= 13 ? (EPI - 13) / 2 : 0) + 1;
    constexpr bool RES = ((EPI - 13) & 1) != 0;
    const int w = (int)(pix % (size_t)p.W);
    const size_t row = pix / (size_t)p.W;  // n*H + h
    const size_t poff = ((row >> 1) * (size_t)(p.W >> 1) + (size_t)(w >> 1)) * p.Cout + col0;
    epilogue_chunk_pool<OUT, RES>(p, r, poff, es.lane, bs);
  } else {
    constexpr int E = (EPI >= 1 && EPI <= 12) ? EPI : 1;
    constexpr int OUT = (E - 1) / 4 + 1;
    constexpr bool DMASK = ((E - 1) & 2) != 0;
    constexpr bool RES = ((E - 1) & 1) != 0;
    const size_t off0 = pix * p.Cout + col0;
    size_t ro = off0;
    float rsc = 1.f;
    if (RES && p.res_pooled) {
      const int w = (int)(pix % (size_t)p.W);
      const size_t row = pix / (size_t)p.W;
      ro = ((row >> 1) * (size_t)(p.W >> 1) + (size_t)(w >> 1)) * p.Cout + col0;
      rsc = 0.25f;
    }
    epilogue_chunk_fast<OUT, DMASK, RES>(p, r, off0, col0, bs, es, ro, rsc);
  }
}

// switch (index) with `kEpi` bound to the compile-time value in `body`
#define SPYR_EPI_CASE(n, ...) case n: { constexpr int kEpi = n; __VA_ARGS__; } break;
#define SPYR_EPI_SWITCH(index, ...)                                                                                  \
  switch (index) {                                                                                                   \
    SPYR_EPI_CASE(0, __VA_ARGS__) SPYR_EPI_CASE(1, __VA_ARGS__) SPYR_EPI_CASE(2, __VA_ARGS__)                        \
    SPYR_EPI_CASE(3, __VA_ARGS__) SPYR_EPI_CASE(4, __VA_ARGS__) SPYR_EPI_CASE(5, __VA_ARGS__)                        \
    SPYR_EPI_CASE(6, __VA_ARGS__) SPYR_EPI_CASE(7, __VA_ARGS__) SPYR_EPI_CASE(8, __VA_ARGS__)                        \
    SPYR_EPI_CASE(9, __VA_ARGS__) SPYR_EPI_CASE(10, __VA_ARGS__) SPYR_EPI_CASE(11, __VA_ARGS__)                      \
    SPYR_EPI_CASE(12, __VA_ARGS__) SPYR_EPI_CASE(13, __VA_ARGS__) SPYR_EPI_CASE(14, __VA_ARGS__)                     \
    SPYR_EPI_CASE(15, __VA_ARGS__) SPYR_EPI_CASE(16, __VA_ARGS__) SPYR_EPI_CASE(17, __VA_ARGS__)                     \
    default: { constexpr int kEpi = 18; __VA_ARGS__; } break;                                                        \
  }

}  // namespace halo
