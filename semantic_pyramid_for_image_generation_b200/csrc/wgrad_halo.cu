// Halo-tiled weight gradient of the 3x3 convolutions (sm_100a):
//     dw[tap][ci][co] += sum_{pixels p} x[p + d_tap][ci] * dy[p][co]
//
// The per-tap kernel in conv_tc.cu re-loads a shifted activation tile for every tap and is L2->SMEM bound (18 % of
// the tensor peak, profiles/r01_conv_profile_halo_v2.txt).  Here one CTA owns an input-channel block, a group of
// taps and an output-channel block, and walks over 8x16-pixel tiles of its share of the batch.  Per tile it loads
//   * ONE halo tile of x per 64-channel chunk: (8+2) x (16+2) pixels x 128 B (SWIZZLE_128B, TMA OOB fill = padding)
//   * ONE 8x16 tile of dy per 64 output channels
// and issues, for every tap of its group, tcgen05.mma with BOTH operands MN-major (the reduction runs over pixels):
// the A descriptor of tap (dy,dx) starts at halo row dy*10+dx, its 8-pixel K groups are 10 rows = 1280 B apart
// (stride-byte-offset), and the two 64-row halves of the M=128 accumulator are
//   * the two 64-channel chunks of a 128-channel block (leading-byte-offset = halo buffer size), or
//   * for 64-channel inputs, two different taps (leading-byte-offset = distance of their windows).
// Each (half, tap) accumulator lives in TMEM for the whole pixel loop.  The pixel range is split over CTAs so that the
// grid fills the 148 SMs: split z stores its partial sum to slice z of a scratch buffer and a second kernel
// (wgrad_reduce_kernel, conv_tc.cu) adds the slices in order; a single split adds to the gradient directly.  No
// floating-point atomics: the result is bit-reproducible.
#include "common.cuh"
#include "../../include/spyramid_b200.h"

extern void spyr_count_launch();
void spyr_note_kernel(int id);

namespace {

constexpr int MAX_GROUPS = 3;
constexpr int HALO_ROWS = 180;                 // (8+2) x (16+2)
constexpr int HALO_BYTES = 24 * 1024;          // 180 rows x 128 B = 23040, padded to a multiple of 1024
constexpr int TILE_PIX = 128;                  // 8 x 16 output pixels = reduction length of one stage

struct Group {           // one M=128 accumulator
  int row0;              // halo row where the window of the lower half starts
  int lbo;               // bytes from the lower to the upper 64-row half
  int tap_lo, ci_lo;     // destination of rows 0..63  (tap < 0: discard)
  int tap_hi, ci_hi;     // destination of rows 64..127
};

struct WHParams {
  int B, H, W, Cin, Cout, cin_stride;
  int tiles_w, tiles_h, ptiles;
  int nchunk;            // 64-channel chunks of x per CTA (1 or 2)
  int ngroups;
  Group groups[MAX_GROUPS];
  int ci0_stride;        // input channels per ci block (64 * nchunk)
  int tapsets;           // tap groups per (ci block, co block)
  int units;             // ci blocks * tapsets * co blocks
  int ci_blocks, co_blocks;
  int block_n, bn_cols, b_chunks;
  int stage_bytes, stages, splits;
  uint32_t tmem_cols;
  float* dw;
  float* partial;        // scratch [slices][taps*cin_stride*Cout] or nullptr (direct accumulation into dw)
  int slice0;
};

struct WHMaps {
  CUtensorMap x, dy;
};

// tap sets.  nchunk == 2: three sets, one kernel row each, every tap is one accumulator (halves = channel chunks).
// nchunk == 1: two sets pairing taps: {(0,3) (1,4) (2,5)} and {(6,7) (7,8 -> keep only 8)}.
__device__ __forceinline__ int set_groups(const WHParams& p, int tapset, Group* g) {
  if (p.nchunk == 2) {
    for (int dx = 0; dx < 3; ++dx) {
      const int tap = tapset * 3 + dx;
      g[dx].row0 = tapset * 10 + dx;
      g[dx].lbo = HALO_BYTES;
      g[dx].tap_lo = tap; g[dx].ci_lo = 0;
      g[dx].tap_hi = tap; g[dx].ci_hi = 64;
    }
    return 3;
  }
  if (tapset == 0) {
    for (int dx = 0; dx < 3; ++dx) {
      g[dx].row0 = dx;
      g[dx].lbo = 10 * 128;
      g[dx].tap_lo = dx; g[dx].ci_lo = 0;
      g[dx].tap_hi = 3 + dx; g[dx].ci_hi = 0;
    }
    return 3;
  }
  g[0].row0 = 20; g[0].lbo = 128; g[0].tap_lo = 6; g[0].ci_lo = 0; g[0].tap_hi = 7; g[0].ci_hi = 0;
  g[1].row0 = 21; g[1].lbo = 128; g[1].tap_lo = -1; g[1].ci_lo = 0; g[1].tap_hi = 8; g[1].ci_hi = 0;
  return 2;
}

__global__ void __launch_bounds__(256, 1)
wgrad_halo_kernel(const __grid_constant__ WHMaps maps, const WHParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* done_bar = empty_bar + p.stages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // unit decode: blockIdx.x = unit, blockIdx.y = pixel split
  int u = blockIdx.x;
  const int tapset = u % p.tapsets;
  u /= p.tapsets;
  const int cib = u % p.ci_blocks;
  const int cob = u / p.ci_blocks;
  const int ci_base = cib * p.ci0_stride, n_off = cob * p.block_n;
  const int per = (p.ptiles + p.splits - 1) / p.splits;
  const int t_begin = blockIdx.y * per;
  const int t_end = min(p.ptiles, t_begin + per);
  const int a_bytes = p.nchunk * HALO_BYTES;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.dy);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  Group groups[MAX_GROUPS];
  const int ngroups = set_groups(p, tapset, groups);

  if (t_begin < t_end) {
    if (warp == 0) {
      if (lane == 0) {
        // ===== TMA producer =====
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
          int r = t;
          const int w0 = (r % p.tiles_w) * 8;
          r /= p.tiles_w;
          const int h0 = (r % p.tiles_h) * 16;
          const int n0 = r / p.tiles_h;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * p.stage_bytes;
          uint8_t* b_dst = a_dst + a_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(p.nchunk * HALO_ROWS * 128 + p.b_chunks * TILE_PIX * 128));
          for (int c = 0; c < p.nchunk; ++c)
            tma_load_4d(a_dst + c * HALO_BYTES, &maps.x, &full_bar[stage], ci_base + c * 64, w0 - 1, h0 - 1, n0);
          for (int j = 0; j < p.b_chunks; ++j)
            tma_load_4d(b_dst + j * TILE_PIX * 128, &maps.dy, &full_bar[stage], n_off + j * 64, w0, h0, n0);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      {
        // ===== MMA issuer: warp-uniform operands, one elected lane issues (see conv_halo.cu) =====
        const bool issue = elect_one();
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t idesc = umma_idesc_bf16(128, p.block_n, 1, 1);
        const uint64_t base = ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        // B (dy tile): 8-pixel K groups 1024 B apart, 64-column chunks TILE_PIX*128 B apart
        const uint64_t b_hi = base | ((uint64_t)((TILE_PIX * 128) >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32);
        uint64_t a_hi[MAX_GROUPS];
        uint32_t a_row16[MAX_GROUPS];
#pragma unroll
        for (int g = 0; g < MAX_GROUPS; ++g) {
          a_hi[g] = base | ((uint64_t)((uint32_t)groups[g].lbo >> 4) << 16) | ((uint64_t)(1280 >> 4) << 32);
          a_row16[g] = (uint32_t)groups[g].row0 * 8u;
        }
        const uint32_t smem16 = (smem_u32(smem) & 0x3FFFF) >> 4;
        const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, a16 = (uint32_t)a_bytes >> 4;
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = smem16 + (uint32_t)stage * stage16;
          const uint64_t db0 = b_hi | (uint64_t)(a_lo + a16);
          if (issue) {
#pragma unroll
            for (int g = 0; g < MAX_GROUPS; ++g) {
              if (g < ngroups) {
                const uint64_t da0 = a_hi[g] | (uint64_t)(a_lo + a_row16[g]);
                const uint32_t acc = tmem_u + (uint32_t)(g * p.bn_cols);
#pragma unroll
                for (int k = 0; k < TILE_PIX / 16; ++k)
                  // 16 pixels per MMA: two 8-pixel tile rows -> A advances 2 halo rows of 10 pixels (2560 B), B 2048 B
                  umma_bf16(acc, da0 + (uint64_t)(k * (2560 >> 4)), db0 + (uint64_t)(k * (2048 >> 4)), idesc,
                            (t > t_begin || k > 0) ? 1u : 0u);
              }
            }
            umma_commit(&empty_bar[stage]);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (issue) umma_commit(done_bar);
        __syncwarp();
      }
    } else if (warp >= 4) {
      // ===== epilogue: TMEM -> partial-sum slice (plain stores) or dw (single split: owner read-modify-write) =====
      float* out = p.partial != nullptr ? p.partial + (size_t)(p.slice0 + (int)blockIdx.y) * 9 * p.cin_stride * p.Cout : p.dw;
      const int q = warp & 3;
      const int m = q * 32 + lane;
      mbar_wait(done_bar, 0);
      tc_fence_after();
      for (int g = 0; g < ngroups; ++g) {
        const int tap = (m < 64) ? groups[g].tap_lo : groups[g].tap_hi;
        const int ci = ci_base + ((m < 64) ? groups[g].ci_lo : groups[g].ci_hi) + (m & 63);
        const bool valid = tap >= 0 && ci < p.Cin && ci < p.cin_stride;
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + (uint32_t)(g * p.bn_cols) + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
          tmem_ld_wait();
          if (!valid) continue;
          const int col0 = n_off + c0;
          float* dst = out + ((size_t)tap * p.cin_stride + ci) * p.Cout + col0;
          wgrad_store32(dst, r, p.Cout - col0, p.partial == nullptr);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

uint32_t pow2_at_least(int n, uint32_t lo) {
  uint32_t c = lo;
  while ((int)c < n) c <<= 1;
  return c;
}

}  // namespace

// Fills the launch plan; returns 0 when the halo-tiled kernel takes the problem, -1 if it is not eligible (the caller uses
// the per-tap kernel), >0 on error.
static int halo_plan(const spyr_wgrad_desc* d, WHParams* pp) {
  WHParams& p = *pp;
  if (d->ksize != 3 || d->per_image) return -1;
  if (d->H < 16 || d->W < 8 || (d->H % 16) != 0 || (d->W % 8) != 0) return -1;
  if ((d->Cin % 64) != 0 || (d->Cout % 8) != 0) return -1;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
  p.cin_stride = d->cin_stride > 0 ? d->cin_stride : d->Cin;
  p.tiles_w = d->W / 8;
  p.tiles_h = d->H / 16;
  p.ptiles = p.tiles_w * p.tiles_h * d->B;
  p.nchunk = (d->Cin % 128 == 0) ? 2 : 1;
  p.ci0_stride = 64 * p.nchunk;
  p.ci_blocks = d->Cin / p.ci0_stride;
  p.tapsets = p.nchunk == 2 ? 3 : 2;
  int bn = d->Cout >= 128 ? 128 : ((d->Cout + 15) / 16) * 16;
  if (bn < 32) bn = 32;
  p.block_n = bn;
  p.bn_cols = (int)pow2_at_least(bn, 32);
  p.b_chunks = ceil_div(bn, 64);
  p.co_blocks = ceil_div(d->Cout, bn);
  p.units = p.ci_blocks * p.tapsets * p.co_blocks;
  p.tmem_cols = pow2_at_least(3 * p.bn_cols, 32);
  if (p.tmem_cols > 512) return -1;
  p.stage_bytes = ceil_div(p.nchunk * HALO_BYTES + p.b_chunks * TILE_PIX * 128, 1024) * 1024;
  int stages = (200 * 1024) / p.stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) return -1;
  p.stages = stages;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SPYR_CHECK_CUDA(cudaGetDevice(&dev));
    SPYR_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // one wave of CTAs: every extra split adds a full accumulator of FP32 reductions to the epilogue
  int splits = d->splits > 0 ? d->splits : num_sms / p.units;
  if (splits < 1) splits = 1;
  if (splits > p.ptiles) splits = p.ptiles;
  // short pixel loops cannot fill the 2..5-stage pipeline: small maps stay on the per-tap kernel (64-pixel steps)
  if (d->splits <= 0 && p.ptiles / splits < 6) return -1;
  {
    // every split owns at least one pixel tile (its slice of the partial sums is read unconditionally)
    const int per = ceil_div(p.ptiles, splits);
    splits = ceil_div(p.ptiles, per);
  }
  p.splits = splits;
  p.dw = d->dw;
  return 0;
}

int spyr_wgrad_halo_plan(const spyr_wgrad_desc* d, int* splits) {
  WHParams p;
  const int rc = halo_plan(d, &p);
  if (rc == 0) *splits = p.splits;
  return rc;
}

int spyr_wgrad_halo_launch(const spyr_wgrad_desc* d, const void* x, const void* dy, float* partial, int slice0,
                           cudaStream_t stream) {
  WHParams p;
  const int rc = halo_plan(d, &p);
  if (rc) return rc > 0 ? rc : 2;
  p.partial = partial;
  p.slice0 = slice0;
  const int stages = p.stages, splits = p.splits;
  WHMaps maps;
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t strides[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
    uint32_t box[4] = {64, 10, 18, 1};
    if (spyr_tmap_encode(&maps.x, x, 4, dims, strides, box, 1)) return 3;
  }
  {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t strides[3] = {(uint64_t)d->Cout * 2, (uint64_t)d->W * d->Cout * 2, (uint64_t)d->H * d->W * d->Cout * 2};
    uint32_t box[4] = {64, 8, 16, 1};
    if (spyr_tmap_encode(&maps.dy, dy, 4, dims, strides, box, 1)) return 3;
  }
  const size_t smem_bytes = (size_t)stages * p.stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    SPYR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  dim3 grid(p.units, splits);
  wgrad_halo_kernel<<<grid, 256, smem_bytes, stream>>>(maps, p);
  spyr_note_kernel(3);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
