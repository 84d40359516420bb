// Debug probe (not on the product path): checks that a 3x3 convolution can be driven from ONE halo tile in shared
// memory, i.e. that tcgen05.mma accepts A descriptors whose start address is an arbitrary 128-byte row of a
// SWIZZLE_128B tile written by TMA and whose 8-row groups are (TW+2)*128 bytes apart.
#include "../../semantic_pyramid_for_image_generation_b200/csrc/common.cuh"
#include "../../include/spyramid_b200.h"  // (relative to tests/native)

namespace {

struct ProbeMaps {
  CUtensorMap x;  // {C, W, H, B}, box {64, 10, 18, 1}
  CUtensorMap w;  // {C, Cout, 9}, box {64, BN, 1}
};

__global__ void __launch_bounds__(128, 1)
halo_probe_kernel(const __grid_constant__ ProbeMaps maps, int H, int W, int chunks, int bn, int base_off_mode,
                  float* __restrict__ y, int Cout) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int HALO_BYTES = 24 * 1024;  // 180 rows * 128 B = 23040, padded to a 1024 multiple
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + HALO_BYTES;              // 9 taps * bn * 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + 9 * bn * 128);
  uint64_t* mma_bar = bar + 1;
  uint32_t* holder = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = W / 8, tiles_h = H / 16;
  int t = blockIdx.x;
  const int w0 = (t % tiles_w) * 8;
  t /= tiles_w;
  const int h0 = (t % tiles_h) * 16;
  const int n0 = t / tiles_h;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(holder, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder;
  uint32_t phase = 0;
  for (int c = 0; c < chunks; ++c) {
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(bar, 180 * 128 + 9 * bn * 128);
      tma_load_4d(a_s, &maps.x, bar, c * 64, w0 - 1, h0 - 1, n0);
      for (int tap = 0; tap < 9; ++tap) tma_load_3d(b_s + tap * bn * 128, &maps.w, bar, c * 64, 0, tap);
    }
    mbar_wait(bar, phase);
    tc_fence_after();
    if (threadIdx.x == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, bn, 0, 0);
      const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t row0 = (uint32_t)((tap / 3) * 10 + (tap % 3));  // halo row of output pixel (0,0) for this tap
        for (int k = 0; k < 4; ++k) {
          const uint32_t start = a_addr + row0 * 128 + k * 32;
          uint64_t da = umma_smem_desc_sw128(start, 0, 1280);
          if (base_off_mode == 1) da |= (uint64_t)((start >> 7) & 7) << 49;
          const uint64_t db = umma_smem_desc_sw128(b_addr + tap * bn * 128 + k * 32, 0, 1024);
          umma_bf16(tmem, da, db, idesc, (c > 0 || tap > 0 || k > 0) ? 1u : 0u);
        }
      }
      umma_commit(mma_bar);
    }
    mbar_wait(mma_bar, phase);
    tc_fence_after();
    phase ^= 1;
  }
  // epilogue: thread m <-> pixel (th = m / 8, tw = m % 8)
  const int m = warp * 32 + lane;
  const int h = h0 + m / 8, w = w0 + m % 8;
  for (int c0 = 0; c0 < bn; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    tmem_ld_wait();
    float* dst = y + (((size_t)n0 * H + h) * W + w) * Cout + c0;
    for (int j = 0; j < 32; ++j)
      if (c0 + j < Cout) dst[j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

}  // namespace

// x NHWC bf16 [B,H,W,C] (C % 64 == 0, H % 16 == 0, W % 8 == 0), w bf16 [9][Cout][C] (Cout <= 64), y f32 [B,H,W,Cout]
extern "C" int spyr_dbg_halo_probe(const void* x, const void* w, float* y, int B, int H, int W, int C, int Cout,
                                   int base_off_mode, void* stream) {
  SPYR_REQUIRE(C % 64 == 0 && H % 16 == 0 && W % 8 == 0 && Cout <= 64 && Cout % 16 == 0, "halo_probe: bad shape");
  ProbeMaps maps;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, 10, 18, 1};
    if (spyr_tmap_encode(&maps.x, x, 4, dims, strides, box, 1)) return 3;
  }
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)Cout, 9};
    uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)Cout * C * 2};
    uint32_t box[3] = {64, (uint32_t)Cout, 1};
    if (spyr_tmap_encode(&maps.w, w, 3, dims, strides, box, 1)) return 3;
  }
  const size_t smem = 24 * 1024 + 9 * Cout * 128 + 64 + 1024;
  SPYR_CHECK_CUDA(cudaFuncSetAttribute(halo_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  halo_probe_kernel<<<B * (H / 16) * (W / 8), 128, smem, (cudaStream_t)stream>>>(maps, H, W, C / 64, Cout, base_off_mode, y,
                                                                                 Cout);
  SPYR_LAUNCH_CHECK();
  return 0;
}
