// Shared device/host helpers for the sm_100a kernels of the Semantic-Pyramid training step.
// PTX wrappers for mbarrier / TMA / tcgen05 (UMMA + TMEM), error plumbing for the C-ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/spyramid_b200.h"

// ------------------------------------------------------------------------------------------------
// C-ABI error plumbing: every entry point returns int (0 = ok); message via spyr_last_error().
// ------------------------------------------------------------------------------------------------
void spyr_set_error(const char* fmt, ...);
#define SPYR_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      spyr_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e), \
                     cudaGetErrorString(_e));                                              \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)
#define SPYR_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      spyr_set_error(__VA_ARGS__);     \
      return 2;                        \
    }                                  \
  } while (0)
#define SPYR_LAUNCH_CHECK() SPYR_CHECK_CUDA(cudaGetLastError())
// kernels that split a flat element index with 32-bit arithmetic (64-bit integer division costs ~100 instructions)
#define SPYR_N32(n) \
  SPYR_REQUIRE((long long)(n) < 2147483647LL, "%s: %lld elements exceed the 32-bit index range", __func__, (long long)(n))

typedef __nv_bfloat16 bf16;

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- TMA (cp.async.bulk.tensor, tiled mode, OOB -> zero fill) ----
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// TMA store of a 4-D box (shared -> global, bulk async-group completion); out-of-bounds elements are clipped.
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], BF16 x BF16 -> FP32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (thread i of the warp <-> lane base+i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors (cute/arch/mma_sm100_desc.hpp bit layout) ----
// smem matrix descriptor: start[0,14) (addr>>4), LBO[16,30) (>>4), SBO[32,46) (>>4), version=1 [46,48),
// layout_type[61,64): 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor kind::f16, BF16 inputs, FP32 accumulate, M=128.
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;   // c_format = F32
  d |= 1u << 7;   // a_format = BF16
  d |= 1u << 10;  // b_format = BF16
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// ---- 256-bit global accesses (sm_100: LDG/STG.256): half the LSU requests of 128-bit vectors ----
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- misc math ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float lrelu_f(float v, float slope) { return v > 0.f ? v : v * slope; }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}

// ------------------------------------------------------------------------------------------------
// Split-BF16 ("strict") precision mode.  spyr_set_precision(1) switches every BF16 feature map and packed weight to a
// pair of planes: hi = bf16(v), lo = bf16(v - hi), the lo plane stored directly behind the hi plane (at element offset
// n = number of elements of the map).  Kernels see such a map as an `Act`: a pointer plus the distance of its lo plane
// (0 in the default single-plane mode).  v is recovered as float(hi) + float(lo) (~16 mantissa bits).
// ------------------------------------------------------------------------------------------------
bool spyr_split();  // capi.cu: current precision mode (process-wide)

// S is the compile-time precision mode of a kernel instantiation: with S = false every lo-plane access is compiled out
// (as a run-time test on `lo` the extra code cost the bandwidth-bound kernels 10-50 % through registers and code size).
// Host code builds the generic ActT<true> (`Act`); kernels take ActT<S> and launches go through SPYR_WITH_SPLIT.
template <bool S>
struct ActT {
  bf16* p;
  long long lo;  // elements from the hi plane to the lo plane; 0 = single plane (always ignored when S is false)
  __host__ __device__ __forceinline__ ActT() : p(nullptr), lo(0) {}
  __host__ __device__ __forceinline__ ActT(bf16* p_, long long lo_) : p(p_), lo(lo_) {}
  template <bool T>
  __host__ __device__ __forceinline__ ActT(const ActT<T>& o) : p(o.p), lo(o.lo) {}
  __host__ __device__ __forceinline__ ActT operator+(long long off) const { return ActT(p + off, lo); }
  __host__ __device__ __forceinline__ ActT operator+(size_t off) const { return ActT(p + off, lo); }
  __host__ __device__ __forceinline__ ActT operator+(int off) const { return ActT(p + off, lo); }
  __host__ __device__ __forceinline__ bool null() const { return p == nullptr; }
};
typedef ActT<true> Act;
// host: the map at `ptr` with `n` elements, in the current precision mode
static inline Act make_act(const void* ptr, long long n) {
  return Act(reinterpret_cast<bf16*>(const_cast<void*>(ptr)), (ptr != nullptr && spyr_split()) ? n : 0);
}
// runs a launch statement with `kS` bound to the current precision mode as a compile-time constant
#define SPYR_WITH_SPLIT(...)      \
  do {                            \
    if (spyr_split()) {           \
      constexpr bool kS = true;    \
      __VA_ARGS__;                \
    } else {                      \
      constexpr bool kS = false;   \
      __VA_ARGS__;                \
    }                             \
  } while (0)

__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2(w[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
// 8 consecutive channels of a map -> FP32
__device__ __forceinline__ void ld8(const bf16* p, float* v) { unpack8(*reinterpret_cast<const uint4*>(p), v); }
template <bool S>
__device__ __forceinline__ void ld8(const ActT<S>& a, float* v) {
  unpack8(*reinterpret_cast<const uint4*>(a.p), v);
  if (S && a.lo != 0) {
    float l[8];
    unpack8(*reinterpret_cast<const uint4*>(a.p + a.lo), l);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += l[j];
  }
}
template <bool S>
__device__ __forceinline__ void ld8_nc(const ActT<S>& a, float* v) {  // read-only path
  unpack8(__ldg(reinterpret_cast<const uint4*>(a.p)), v);
  if (S && a.lo != 0) {
    float l[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(a.p + a.lo)), l);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += l[j];
  }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]);
  o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]);
  o.w = pack_bf16x2(v[6], v[7]);
  return o;
}
__device__ __forceinline__ void st8(bf16* p, const float* v) { *reinterpret_cast<uint4*>(p) = pack8(v); }
// residual of the BF16 rounding: lo[j] = v[j] - float(bf16(v[j]))  (exact in FP32)
__device__ __forceinline__ void split_residual8(const float* v, float* lo) {
#pragma unroll
  for (int j = 0; j < 8; ++j) lo[j] = v[j] - __bfloat162float(__float2bfloat16(v[j]));
}
template <bool S>
__device__ __forceinline__ void st8(const ActT<S>& a, const float* v) {
  *reinterpret_cast<uint4*>(a.p) = pack8(v);
  if (S && a.lo != 0) {
    float l[8];
    split_residual8(v, l);
    *reinterpret_cast<uint4*>(a.p + a.lo) = pack8(l);
  }
}
// what a later pass reads back from a map written with st8: the stored value (BF16, or hi + lo in split mode)
template <bool S>
__device__ __forceinline__ float stored_value(float v, const ActT<S>& a) {
  const float h = __bfloat162float(__float2bfloat16(v));
  return (S && a.lo != 0) ? h + __bfloat162float(__float2bfloat16(v - h)) : h;
}

// ------------------------------------------------------------------------------------------------
// Deterministic grid-wide reductions: every block writes its partial vector to a caller-provided scratch buffer, the
// last block to arrive (ticket counter) sums the partials in block order.  No floating-point atomics anywhere, so
// results are bit-identical from run to run.  Tickets come from a small per-translation-unit pool and reset themselves.
// ------------------------------------------------------------------------------------------------
constexpr int SPYR_TICKETS = 4096;
static __device__ unsigned int spyr_ticket_pool[SPYR_TICKETS];
static __device__ unsigned int spyr_ticket_pool2[SPYR_TICKETS];  // blocks of consecutive tickets (spyr_next_tickets)
// host: next ticket of this translation unit (round robin; a ticket is busy only while its kernel runs)
static inline unsigned int* spyr_next_ticket() {
  static unsigned int* base = nullptr;
  static unsigned int next = 0;
  if (base == nullptr) {
    void* sym = nullptr;
    if (cudaGetSymbolAddress(&sym, spyr_ticket_pool) != cudaSuccess) return nullptr;
    base = reinterpret_cast<unsigned int*>(sym);
  }
  return base + (next++ % SPYR_TICKETS);
}
// `n` consecutive tickets (per-sample tails); nullptr when n exceeds the pool
static inline unsigned int* spyr_next_tickets(int n) {
  static unsigned int next = 0;
  if (n > SPYR_TICKETS / 4) return nullptr;
  static unsigned int* pool = nullptr;
  if (pool == nullptr) {
    void* sym = nullptr;
    if (cudaGetSymbolAddress(&sym, spyr_ticket_pool2) != cudaSuccess) return nullptr;
    pool = reinterpret_cast<unsigned int*>(sym);
  }
  if (next + (unsigned)n > (unsigned)SPYR_TICKETS) next = 0;
  unsigned int* out = pool + next;
  next += (unsigned)n;
  return out;
}
// true in exactly one block of the grid: the one that arrives last.  Call with all threads of the block after the block's
// partial results have been written to global memory.
__device__ __forceinline__ bool spyr_last_block(unsigned int* ticket, unsigned int nblocks) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == nblocks - 1) ? 1u : 0u;
    if (s_last) *ticket = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0u;
}
// scalar element access of a map (slow path: layout conversions, 3-channel tails)
template <bool S>
__device__ __forceinline__ float ldf(const ActT<S>& a, size_t i) {
  float v = __bfloat162float(a.p[i]);
  if (S && a.lo != 0) v += __bfloat162float(a.p[i + a.lo]);
  return v;
}
template <bool S>
__device__ __forceinline__ void stf(const ActT<S>& a, size_t i, float v) {
  const bf16 h = __float2bfloat16(v);
  a.p[i] = h;
  if (S && a.lo != 0) a.p[i + a.lo] = __float2bfloat16(v - __bfloat162float(h));
}
// Tail of a deterministic grid reduction, run by the LAST block only (spyr_last_block): sums the partial vectors
// scratch[b][n], b in [0, nb), in a fixed order and hands each total to sink(i, total).  blockDim.x / n threads share
// one output (block-strided partial sums combined in lane order), so the summation order depends on the launch shape
// only.  All threads of the block must call it.
template <typename T, typename Sink>
__device__ __forceinline__ void spyr_sum_partials(const T* __restrict__ scratch, int nb, int n, Sink sink) {
  __shared__ T tmp[1024];
  const int nt = (int)blockDim.x;
  int lanes = nt / n;
  if (lanes < 1) lanes = 1;
  if (lanes > 16) lanes = 16;
  const int per_pass = nt / lanes;
  const int local = (int)threadIdx.x / lanes, lane = (int)threadIdx.x % lanes;
  for (int i0 = 0; i0 < n; i0 += per_pass) {
    const int i = i0 + local;
    T s = (T)0;
    if (local < per_pass && i < n) {
#pragma unroll 8
      for (int b = lane; b < nb; b += lanes) s += __ldcg(scratch + (size_t)b * n + i);
    }
    tmp[threadIdx.x] = s;
    __syncthreads();
    if (lane == 0 && local < per_pass && i < n) {
      T t = (T)0;
      for (int l = 0; l < lanes; ++l) t += tmp[threadIdx.x + l];
      sink(i, t);
    }
    __syncthreads();
  }
}
// Second stage of a deterministic grid reduction as its own (parallel) kernel: one warp per output sums the nb partial
// vectors scratch[b][n] (lane-strided, then a fixed xor tree) and hands the total to a sink.  Used where n * nb is large
// (bias column sums, stencil weight gradients): a single last block would serialise ~1M loads.
//   mode 0: out0[i] += t (+ out1, out2 when given)
//   mode 1: out0[((i / C) * cin_stride + ci_row) * C + i % C] += t      (mask-channel rows of a strided weight gradient)
//   mode 2: i < split ? out0[i] += t : out1[i - split] += t
struct SumSink {
  float* out0;
  float* out1;
  float* out2;
  int mode, C, cin_stride, ci_row, split;
};
__device__ __forceinline__ float spyr_warp_tree(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void spyr_sum_partials_body(const float* __restrict__ scratch, int nb, int n, const SumSink& sk,
                                                       long long block) {
  const int i = (int)((block * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  float s = 0.f;
  for (int b = lane; b < nb; b += 32) s += __ldcg(scratch + (size_t)b * n + i);
  s = spyr_warp_tree(s);
  if (lane != 0) return;
  if (sk.mode == 0) {
    sk.out0[i] += s;
    if (sk.out1 != nullptr) sk.out1[i] += s;
    if (sk.out2 != nullptr) sk.out2[i] += s;
  } else if (sk.mode == 1) {
    sk.out0[((size_t)(i / sk.C) * sk.cin_stride + sk.ci_row) * sk.C + i % sk.C] += s;
  } else {
    if (i < sk.split) sk.out0[i] += s;
    else sk.out1[i - sk.split] += s;
  }
}
static __global__ void spyr_sum_partials_kernel(const float* __restrict__ scratch, int nb, int n, SumSink sk) {
  spyr_sum_partials_body(scratch, nb, n, sk, (long long)blockIdx.x);
}
static inline cudaError_t spyr_launch_sum_partials(const float* scratch, int nb, int n, const SumSink& sk, cudaStream_t st) {
  spyr_sum_partials_kernel<<<(n + 7) / 8, 256, 0, st>>>(scratch, nb, n, sk);
  return cudaGetLastError();
}
// SPYR_REDUCE_BLOCKS (include/spyramid_b200.h) bounds the grid of every reduction kernel, and so its scratch

// 32 accumulator columns of one row -> global memory; `ncols` columns are valid.  accumulate: dst += (the row has ONE
// owner, so a plain read-modify-write is race-free and deterministic); else plain store into a partial-sum slice.
__device__ __forceinline__ void wgrad_store32(float* dst, const uint32_t* r, int ncols, bool accumulate) {
  if (ncols >= 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                             __uint_as_float(r[j + 3]));
      if (accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(dst + j);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *reinterpret_cast<float4*>(dst + j) = v;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) dst[j] = accumulate ? dst[j] + __uint_as_float(r[j]) : __uint_as_float(r[j]);
  }
}

// ---- host: TMA descriptor encode through the runtime's driver entry point (no -lcuda link) ----
int spyr_tmap_encode(CUtensorMap* map, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle128);
