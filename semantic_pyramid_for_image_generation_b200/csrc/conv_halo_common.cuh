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


}  // namespace halo
