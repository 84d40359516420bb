// Micro-benchmark of TMA box-load throughput per SM for the tile shapes of the convolution kernels (developer tool):
// one warp per CTA keeps R boxes in flight in a ring of R shared-memory buffers and walks over an NHWC BF16 tensor the
// way the persistent conv kernels do; reports clocks and bytes per box.  It separates "TMA / L2-bound" from "tensor-pipe
// bound" in the per-tile cycle counts of profiles/r01_ncu_conv_halo*_uniform_issue.txt.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o build/tma_probe tests/native/tma_probe.cu \
//          -Lsemantic_pyramid_for_image_generation_b200 -lspyramid_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../semantic_pyramid_for_image_generation_b200'
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../semantic_pyramid_for_image_generation_b200/csrc/common.cuh"

struct Args {
  int C, W, H, B;       // tensor (channels innermost)
  int bw, bh;           // box: 64 channels x bw x bh pixels
  int tiles_w, tiles_h; // tile grid per image (tile = (bw-2) x (bh-2) output pixels for halo boxes)
  int step_w, step_h, halo;
  int ring;             // boxes in flight
  int buf_bytes;
};

__global__ void __launch_bounds__(128, 1) tma_probe_kernel(const __grid_constant__ CUtensorMap map, Args a, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.ring; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&map);
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const bool issue = elect_one();
  const int total = a.tiles_w * a.tiles_h * a.B;
  const uint32_t bytes = (uint32_t)(a.bw * a.bh * 128);
  int issued = 0, done = 0;
  uint32_t phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t0 = clock64();
  int tile = blockIdx.x;
  // prologue: fill the ring
  int pending[8];
  int head = 0, count = 0;
  while (tile < total || count > 0) {
    while (count < a.ring && tile < total) {
      const int slot = (head + count) % a.ring;
      const int w0 = (tile % a.tiles_w) * a.step_w, h0 = ((tile / a.tiles_w) % a.tiles_h) * a.step_h;
      const int n0 = tile / (a.tiles_w * a.tiles_h);
      if (issue) {
        mbar_arrive_expect_tx(&full[slot], bytes);
        tma_load_4d(smem + slot * a.buf_bytes, &map, &full[slot], 0, w0 - a.halo, h0 - a.halo, n0);
      }
      pending[slot] = tile;
      ++count;
      ++issued;
      tile += gridDim.x;
    }
    mbar_wait(&full[head], phase[head]);
    phase[head] ^= 1;
    head = (head + 1) % a.ring;
    --count;
    ++done;
  }
  const long long t1 = clock64();
  if (issue) {
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = done;
  }
  (void)pending;
}

static void run(const char* label, int C, int H, int W, int B, int bw, int bh, int halo, int ring) {
  Args a;
  a.C = C; a.W = W; a.H = H; a.B = B; a.bw = bw; a.bh = bh; a.halo = halo;
  a.step_w = bw - 2 * halo; a.step_h = bh - 2 * halo;
  a.tiles_w = W / a.step_w; a.tiles_h = H / a.step_h;
  a.ring = ring;
  a.buf_bytes = ((bw * bh * 128 + 1023) / 1024) * 1024;
  void* d;
  const size_t bytes = (size_t)B * H * W * C * 2;
  cudaMalloc(&d, bytes);
  cudaMemset(d, 0, bytes);
  CUtensorMap map;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
  if (spyr_tmap_encode(&map, d, 4, dims, strides, box, 1)) { printf("tmap encode failed\n"); exit(1); }
  long long* out;
  cudaMalloc(&out, sizeof(long long) * 2 * 148);
  const int smem = ring * a.buf_bytes + 2048;
  cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  tma_probe_kernel<<<148, 128, smem>>>(map, a, out);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  tma_probe_kernel<<<148, 128, smem>>>(map, a, out);
  cudaEventRecord(e1);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: CUDA error %s\n", label, cudaGetErrorString(cudaGetLastError())); exit(1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(2 * 148);
  cudaMemcpy(h.data(), out, sizeof(long long) * 2 * 148, cudaMemcpyDeviceToHost);
  double clk = 0, boxes = 0;
  for (int i = 0; i < 148; ++i) { clk += h[2 * i]; boxes += h[2 * i + 1]; }
  const double total_boxes = (double)a.tiles_w * a.tiles_h * B;
  printf("%-34s C=%3d %dx%d box %2dx%2d ring %d: %7.0f clk/box  %5.1f clk/row  %6.2f TB/s  (%.0f boxes, %.3f ms)\n", label, C, H, W,
         bw, bh, ring, clk / boxes, clk / boxes / (bw * bh), total_boxes * bw * bh * (C < 64 ? C : 64) * 2 / (ms * 1e-3) / 1e12,
         total_boxes, ms);
  cudaFree(d); cudaFree(out);
}

int main() {
  for (int ring : {1, 2, 3, 4}) run("halo 3x3, 64 ch, msub 2", 64, 256, 256, 20, 10, 34, 1, ring);
  for (int ring : {2, 4}) run("halo 3x3, 64 ch, msub 1", 64, 256, 256, 20, 10, 18, 1, ring);
  for (int ring : {2, 4}) run("1x1, 32 ch (im2col rows)", 32, 256, 256, 20, 8, 32, 0, ring);
  for (int ring : {2, 4}) run("1x1, 64 ch", 64, 256, 256, 20, 8, 32, 0, ring);
  for (int ring : {2, 4}) run("halo 3x3, chunk of 256 ch", 256, 64, 64, 20, 10, 18, 1, ring);
  return 0;
}
