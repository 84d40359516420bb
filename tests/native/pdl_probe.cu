// Developer probe: what a kernel -> kernel edge costs inside a CUDA graph on this GPU, with plain stream-order edges and
// with programmatic dependent launch (the consumer calls griddepcontrol.wait first thing, so only its launch / block
// scheduling overlaps the producer).  Chains of N tiny kernels (one CTA wave) and of N "persistent" kernels (148 CTAs x
// 200 KB dynamic shared memory, like the convolution kernels, which cannot co-reside with their successor).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <bool PDL>
__global__ void link(float* buf, int spin) {
  extern __shared__ float sm[];
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  float v = buf[threadIdx.x & 31];
  for (int i = 0; i < spin; ++i) v = v * 1.0001f + 0.5f;
  if (threadIdx.x == 0 && blockIdx.x == 0) buf[0] = v;
  if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <bool PDL>
static float run(int n, int grid, int block, size_t smem, int spin, float* buf) {
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  CK(cudaFuncSetAttribute(link<PDL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaGraph_t g;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, link<PDL>, buf, spin));
  }
  CK(cudaStreamEndCapture(st, &g));
  cudaGraphExec_t ge;
  CK(cudaGraphInstantiate(&ge, g, 0));
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a, st));
  for (int i = 0; i < 10; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(b, st));
  CK(cudaStreamSynchronize(st));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms * 1e3f / (10.f * n);
}

int main() {
  float* buf;
  CK(cudaMalloc(&buf, 4096));
  CK(cudaMemset(buf, 0, 4096));
  const int n = 200;
  for (int spin : {0, 2000, 20000}) {
    printf("spin %5d  tiny (8 CTAs x 128):        plain %.2f us/kernel   PDL %.2f us/kernel\n", spin,
           run<false>(n, 8, 128, 0, spin, buf), run<true>(n, 8, 128, 0, spin, buf));
    printf("spin %5d  wave (296 CTAs x 256):       plain %.2f us/kernel   PDL %.2f us/kernel\n", spin,
           run<false>(n, 296, 256, 0, spin, buf), run<true>(n, 296, 256, 0, spin, buf));
    printf("spin %5d  persistent (148 x 384, 200K): plain %.2f us/kernel   PDL %.2f us/kernel\n", spin,
           run<false>(n, 148, 384, 200 * 1024, spin, buf), run<true>(n, 148, 384, 200 * 1024, spin, buf));
  }
  return 0;
}
