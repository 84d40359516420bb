// Micro-benchmark of tcgen05.mma issue/throughput on sm_100a: clocks per M128xNx16 BF16 instruction as a function of
// N, accumulator reuse, operand majors, descriptor strides and the A-operand source (shared memory or TMEM).
// Developer tool (not part of the product path): it explains the per-layer TFLOP/s in profiles/ and picks tile shapes.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o build/mma_probe tests/native/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../semantic_pyramid_for_image_generation_b200/csrc/common.cuh"

struct ProbeArgs {
  int M, N;          // instruction shape
  int n_acc;         // accumulators cycled round-robin (1 = every MMA accumulates into the same TMEM columns)
  int a_mn, b_mn;    // operand majors
  int sbo_a;         // stride between 8-row groups of A in bytes (1024 = dense, 1280 = halo layout)
  int a_row0;        // first row of A inside the tile (unaligned starts, halo taps)
  int a_tmem;        // 1: A operand read from tensor memory
  int iters;         // MMAs issued
  int uniform_issue; // 1: warp-uniform operands + elected issue (no per-MMA R2UR/waterfall code)
};

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs a, long long* clocks_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_holder;
  // A region: 64 KB, B region: 64 KB, pseudo-random small BF16 values
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < (128 * 1024) / 4; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    // two bf16 in [-1, 1): exponent 0x3F00 region
    w[i] = (0x3C00u + ((h >> 3) & 0x3FF)) | ((0xBC00u + ((h >> 13) & 0x3FF)) << 16);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_base_holder, 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_holder;
  long long t0 = 0, t1 = 0;
  if (a.uniform_issue) {
    // whole warp 0 runs the loop with warp-uniform operands (uniform datapath); only the MMA itself is elected
    if (threadIdx.x < 32) {
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc = umma_idesc_bf16(a.M, a.N, a.a_mn, a.b_mn);
      const uint32_t a_addr = __shfl_sync(0xffffffffu, smem_u32(smem), 0) + a.a_row0 * 128;
      const uint32_t b_addr = __shfl_sync(0xffffffffu, smem_u32(smem), 0) + 64 * 1024;
      const uint64_t da0 = a.a_mn ? umma_smem_desc_sw128(a_addr, 8192, 1024) : umma_smem_desc_sw128(a_addr, 16, a.sbo_a);
      const uint64_t db0 = a.b_mn ? umma_smem_desc_sw128(b_addr, 8192, 1024) : umma_smem_desc_sw128(b_addr, 16, 1024);
      const uint32_t a_step = a.a_mn ? (2048 >> 4) : (32 >> 4);
      const uint32_t b_step = a.b_mn ? (2048 >> 4) : (32 >> 4);
      const uint32_t a_t = tm + 448;
      const bool leader = elect_one();
      t0 = clock64();
      for (int it = 0; it < a.iters; it += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (leader) {
            if (a.a_tmem)
              umma_bf16_ts(tm, a_t + k * 8, db0 + (uint64_t)(k * b_step), idesc, 1);
            else
              umma_bf16(tm, da0 + (uint64_t)(k * a_step), db0 + (uint64_t)(k * b_step), idesc, 1);
          }
        }
      }
      if (leader) {
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        t1 = clock64();
        clocks_out[blockIdx.x] = t1 - t0;
      }
      __syncwarp();
    }
  } else if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(a.M, a.N, a.a_mn, a.b_mn);
    const uint32_t a_addr = smem_u32(smem) + a.a_row0 * 128;
    const uint32_t b_addr = smem_u32(smem) + 64 * 1024;
    // K-major: LBO unused (1), SBO = group stride; MN-major: LBO = 64-element column-block stride, SBO = 8-row(k) stride
    const uint64_t da0 = a.a_mn ? umma_smem_desc_sw128(a_addr, 8192, 1024) : umma_smem_desc_sw128(a_addr, 16, a.sbo_a);
    const uint64_t db0 = a.b_mn ? umma_smem_desc_sw128(b_addr, 8192, 1024) : umma_smem_desc_sw128(b_addr, 16, 1024);
    const uint32_t a_step = a.a_mn ? (2048 >> 4) : (32 >> 4);  // advance K by 16
    const uint32_t b_step = a.b_mn ? (2048 >> 4) : (32 >> 4);
    // accumulators occupy n_acc * N columns from column 0; a TMEM A operand sits at columns [448, 512)
    const uint32_t a_t = tmem + 448;
    t0 = clock64();
    int acc = 0;
    for (int it = 0; it < a.iters; ++it) {
      const uint32_t k = it & 3;
      const uint32_t d = tmem + acc * a.N;
      if (a.a_tmem)
        umma_bf16_ts(d, a_t + k * 8, db0 + (uint64_t)(k * b_step), idesc, 1);
      else
        umma_bf16(d, da0 + (uint64_t)(k * a_step), db0 + (uint64_t)(k * b_step), idesc, 1);
      if (++acc == a.n_acc) acc = 0;
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    t1 = clock64();
    clocks_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

static double run(const ProbeArgs& a, int ctas, const char* label) {
  long long* d_clk;
  cudaMalloc(&d_clk, sizeof(long long) * ctas);
  const int smem = 129 * 1024 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe_kernel<<<ctas, 128, smem>>>(a, d_clk);  // warm-up
  cudaEventRecord(e0);
  probe_kernel<<<ctas, 128, smem>>>(a, d_clk);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("%-44s CUDA error: %s\n", label, cudaGetErrorString(err));
    exit(1);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(ctas);
  cudaMemcpy(h.data(), d_clk, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (long long v : h) avg += (double)v;
  avg /= ctas;
  const double clk_per = avg / a.iters;
  const double macs = (double)a.M * a.N * 16;
  const double tf = 2.0 * macs * a.iters * ctas / (ms * 1e-3) / 1e12;
  printf("%-44s M=%3d N=%3d acc=%d  %7.1f clk/MMA  %6.0f MAC/clk/SM  %7.1f TFLOP/s (%d CTAs, %.3f ms)\n", label, a.M, a.N,
         a.n_acc, clk_per, macs / clk_per, tf, ctas, ms);
  cudaFree(d_clk);
  return clk_per;
}

int main(int argc, char** argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 148;
  const int iters = argc > 2 ? atoi(argv[2]) : 4096;
  ProbeArgs base{128, 256, 1, 0, 0, 1024, 0, 0, iters, 0};
  for (int N : {64, 128, 256}) {
    ProbeArgs a = base;
    a.N = N;
    run(a, ctas, "K-major A,B  1 accumulator");
    a.n_acc = 512 / N > 4 ? 4 : (448 / N);
    if (a.n_acc < 1) a.n_acc = 1;
    run(a, ctas, "K-major A,B  round-robin accumulators");
  }
  for (int N : {64, 128, 256}) {
    ProbeArgs a = base;
    a.N = N;
    a.uniform_issue = 1;
    run(a, ctas, "uniform issue, smem A");
    a.a_tmem = 1;
    run(a, ctas, "uniform issue, TMEM A");
    a.a_tmem = 0;
    a.M = 64;
    run(a, ctas, "uniform issue, M=64");
  }
  for (int N : {64, 128, 256}) {
    ProbeArgs a = base;
    a.N = N;
    a.uniform_issue = 1;
    a.sbo_a = 1280;
    a.a_row0 = 11;
    run(a, ctas, "uniform issue, halo A (SBO 1280, row0 11)");
    a.b_mn = 1;
    run(a, ctas, "uniform issue, halo A + B MN-major (dgrad)");
    a.sbo_a = 1024;
    a.a_row0 = 0;
    a.a_mn = 1;
    run(a, ctas, "uniform issue, A,B MN-major (wgrad)");
  }
  for (int N : {64, 128, 256}) {
    ProbeArgs a = base;
    a.N = N;
    a.sbo_a = 1280;
    a.a_row0 = 11;
    run(a, ctas, "halo A (SBO 1280, row0 11)");
  }
  for (int N : {64, 128, 256}) {
    ProbeArgs a = base;
    a.N = N;
    a.b_mn = 1;
    run(a, ctas, "B MN-major (dgrad weights)");
    a.a_mn = 1;
    run(a, ctas, "A,B MN-major (wgrad)");
  }
  for (int N : {64, 128, 256}) {
    ProbeArgs a = base;
    a.N = N;
    a.M = 64;
    run(a, ctas, "M=64");
  }
  for (int N : {64, 128, 256}) {
    ProbeArgs a = base;
    a.N = N;
    a.a_tmem = 1;
    run(a, ctas, "A from TMEM");
  }
  {
    ProbeArgs a = base;
    a.N = 256;
    run(a, 1, "single CTA (no chip-level power limit)");
    a.N = 64;
    run(a, 1, "single CTA (no chip-level power limit)");
  }
  return 0;
}
