// Fused SAGAN self-attention forward (sm_100a): one kernel does S = Q K^T, the row softmax over the keys and
// O = softmax(S) V for a tile of 128 queries of one image; the 1024 x 256 attention map never goes through HBM on the
// way to O (it is optionally written once, in BF16, for the backward pass).
//
// Replaces the two torch.bmm + softmax of SelfAttention.forward (reference models.py:262-270; no 1/sqrt(d) scale):
//     attention_map = bmm(query^T, key).softmax(-1);   attention_features = bmm(value, attention_map^T)
//
//   TMA      Q tile [128 q][d<=64] and the whole K [nk<=256][d] (K-major, zero-filled to 64 channels), V [nk][dv] as
//            64-channel MN-major chunks; all SWIZZLE_128B
//   MMA 1    S (128 x nk, FP32) in TMEM columns [0, nk)
//   softmax  four warps, one query row per thread (TMEM lane == row): max, exp, sum straight from TMEM; the
//            un-normalised probabilities go to shared memory as the BF16 K-major A operand of MMA 2 (manual 128B swizzle
//            + fence.proxy.async), the normalised ones optionally to global memory
//   MMA 2    O (128 x dv) in TMEM columns [256, 256+dv); the epilogue scales rows by 1/sum and stores BF16
#include "common.cuh"
#include "../../include/spyramid_b200.h"

extern void spyr_count_launch();

namespace {

struct AttnParams {
  int B, HW, d, nk, dv;
  int ksteps;        // ceil(d / 16) MMAs for S
  int v_chunks;      // dv / 64
  bf16* o;
  bf16* p_out;
};
struct AttnMaps {
  CUtensorMap q, k, v;
};

__global__ void __launch_bounds__(256, 1)
sagan_attention_fwd_kernel(const __grid_constant__ AttnMaps maps, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;                               // 128 x 128 B
  uint8_t* k_s = q_s + 16 * 1024;                    // nk x 128 B (<= 32 KB)
  uint8_t* v_s = k_s + 32 * 1024;                    // v_chunks x nk x 128 B (<= 64 KB)
  uint8_t* p_s = v_s + 64 * 1024;                    // (nk/64) regions x 128 x 128 B (<= 64 KB)
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(p_s + 64 * 1024);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_p = bar_qk + 3;
  uint64_t* bar_o = bar_qk + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_qk + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_image = p.HW / 128;
  const int b = blockIdx.x / tiles_per_image;
  const int q0 = (blockIdx.x % tiles_per_image) * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_qk, (uint32_t)(128 * 128 + p.nk * 128));
      tma_load_3d(q_s, &maps.q, bar_qk, 0, q0, b);
      tma_load_3d(k_s, &maps.k, bar_qk, 0, 0, b);
      mbar_arrive_expect_tx(bar_v, (uint32_t)(p.v_chunks * p.nk * 128));
      for (int j = 0; j < p.v_chunks; ++j) tma_load_3d(v_s + j * p.nk * 128, &maps.v, bar_v, j * 64, 0, b);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint64_t base = ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t kmaj = base | ((uint64_t)(1024 >> 4) << 32);
      // ---- S = Q K^T ----
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t idesc_s = umma_idesc_bf16(128, p.nk, 0, 0);
      const uint64_t dq = kmaj | (uint64_t)((smem_u32(q_s) & 0x3FFFF) >> 4);
      const uint64_t dk = kmaj | (uint64_t)((smem_u32(k_s) & 0x3FFFF) >> 4);
      for (int k = 0; k < p.ksteps; ++k) umma_bf16(tmem_s, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k ? 1u : 0u);
      umma_commit(bar_s);
      // ---- O = P V ----
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_o = umma_idesc_bf16(128, p.dv, 0, 1);
      const uint64_t dp = kmaj | (uint64_t)((smem_u32(p_s) & 0x3FFFF) >> 4);
      // V: MN-major, 64-channel chunks nk*128 B apart (LBO), 8-key groups 1024 B apart (SBO)
      const uint64_t dvv = base | ((uint64_t)(((uint32_t)p.nk * 128u) >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
                           (uint64_t)((smem_u32(v_s) & 0x3FFFF) >> 4);
      const int regions = p.nk / 64;
      for (int g = 0; g < regions; ++g)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_o, dp + (uint64_t)(g * (16384 >> 4) + 2 * k), dvv + (uint64_t)(g * (8192 >> 4) + k * (2048 >> 4)),
                    idesc_o, (g | k) ? 1u : 0u);
      umma_commit(bar_o);
    }
  } else if (warp >= 4) {
    // ===== softmax + epilogue: thread <-> query row =====
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c0 = 0; c0 < p.nk; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_s + lane_addr + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    float sum = 0.f;
    for (int c0 = 0; c0 < p.nk; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_s + lane_addr + (uint32_t)c0, v);
      tmem_ld_wait();
      uint8_t* region = p_s + (c0 >> 6) * 16384 + r * 128;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          e[j] = __expf(__uint_as_float(v[u * 8 + j]) - mx);
          // accumulate what the tensor core will actually see (BF16-rounded), so rows of P sum to one after scaling
          e[j] = __bfloat162float(__float2bfloat16(e[j]));
          sum += e[j];
        }
        uint4 o;
        o.x = pack_bf16x2(e[0], e[1]);
        o.y = pack_bf16x2(e[2], e[3]);
        o.z = pack_bf16x2(e[4], e[5]);
        o.w = pack_bf16x2(e[6], e[7]);
        const int unit = ((c0 & 32) >> 3) + u;  // 16-byte unit inside the 128-byte row of this 64-key region
        *reinterpret_cast<uint4*>(region + ((unit ^ (r & 7)) << 4)) = o;
      }
    }
    // make the generic-proxy writes visible to the tensor core (async proxy), then hand P to the MMA warp
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    const float inv = 1.f / sum;
    if (p.p_out != nullptr) {
      bf16* dst = p.p_out + ((size_t)b * p.HW + q0 + r) * p.nk;
      for (int c0 = 0; c0 < p.nk; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float e[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) e[j] = __expf(__uint_as_float(v[u * 8 + j]) - mx) * inv;
          uint4 o;
          o.x = pack_bf16x2(e[0], e[1]);
          o.y = pack_bf16x2(e[2], e[3]);
          o.z = pack_bf16x2(e[4], e[5]);
          o.w = pack_bf16x2(e[6], e[7]);
          *reinterpret_cast<uint4*>(dst + c0 + u * 8) = o;
        }
      }
    }
    mbar_wait(bar_o, 0);
    tc_fence_after();
    bf16* orow = p.o + ((size_t)b * p.HW + q0 + r) * p.dv;
    for (int c0 = 0; c0 < p.dv; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_o + lane_addr + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[u * 8 + 0]) * inv, __uint_as_float(v[u * 8 + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v[u * 8 + 2]) * inv, __uint_as_float(v[u * 8 + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v[u * 8 + 4]) * inv, __uint_as_float(v[u * 8 + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v[u * 8 + 6]) * inv, __uint_as_float(v[u * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + c0 + u * 8) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

extern "C" int spyr_sagan_attention_fwd(const void* q, const void* k, const void* v, void* o, void* p_out, int B, int HW,
                                        int d, int nk, int dv, void* stream) {
  SPYR_REQUIRE(!spyr_split(), "sagan_attention_fwd: the fused kernel computes with single-plane BF16 operands; in split-BF16 "
               "mode run the attention as per-image GEMMs (spyr_conv2d_fprop with w_per_image) + spyr_softmax_rows_fwd");
  SPYR_REQUIRE(q && k && v && o && B > 0, "sagan_attention_fwd: bad arguments");
  SPYR_REQUIRE(HW % 128 == 0, "sagan_attention_fwd: queries per image (%d) must be a multiple of 128", HW);
  SPYR_REQUIRE(d % 8 == 0 && d >= 8 && d <= 64, "sagan_attention_fwd: query/key width %d must be in 8..64", d);
  SPYR_REQUIRE(nk % 64 == 0 && nk >= 64 && nk <= 256, "sagan_attention_fwd: %d keys (need 64..256, multiple of 64)", nk);
  SPYR_REQUIRE(dv % 64 == 0 && dv >= 64 && dv <= 128, "sagan_attention_fwd: value width %d must be 64 or 128", dv);
  AttnParams p;
  p.B = B; p.HW = HW; p.d = d; p.nk = nk; p.dv = dv;
  p.ksteps = ceil_div(d, 16);
  p.v_chunks = dv / 64;
  p.o = (bf16*)o;
  p.p_out = (bf16*)p_out;
  AttnMaps maps;
  {
    uint64_t dims[3] = {(uint64_t)d, (uint64_t)HW, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)HW * d * 2};
    uint32_t box[3] = {64, 128, 1};
    if (spyr_tmap_encode(&maps.q, q, 3, dims, strides, box, 1)) return 3;
  }
  {
    uint64_t dims[3] = {(uint64_t)d, (uint64_t)nk, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)nk * d * 2};
    uint32_t box[3] = {64, (uint32_t)nk, 1};
    if (spyr_tmap_encode(&maps.k, k, 3, dims, strides, box, 1)) return 3;
  }
  {
    uint64_t dims[3] = {(uint64_t)dv, (uint64_t)nk, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)dv * 2, (uint64_t)nk * dv * 2};
    uint32_t box[3] = {64, (uint32_t)nk, 1};
    if (spyr_tmap_encode(&maps.v, v, 3, dims, strides, box, 1)) return 3;
  }
  const size_t smem_bytes = (16 + 32 + 64 + 64) * 1024 + 64 + 1024;
  static bool configured = false;
  if (!configured) {
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(sagan_attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes));
    configured = true;
  }
  sagan_attention_fwd_kernel<<<B * (HW / 128), 256, smem_bytes, (cudaStream_t)stream>>>(maps, p);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
