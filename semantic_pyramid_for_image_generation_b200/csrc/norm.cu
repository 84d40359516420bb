// Batch-norm / conditional batch-norm (train-mode batch statistics) fused with LeakyReLU and the
// bilinear x2 (align_corners=True) upsampling that follows or precedes it in the generator.
//
// Replaces: ConditionalBatchNorm.forward (models.py:491-506), nn.BatchNorm2d(64) + nn.UpsamplingBilinear2d +
// nn.LeakyReLU of Generator.final_block (models.py:51-54), and the CBN -> LeakyReLU -> UpsamplingBilinear2d chain
// of GeneratorResidualBlock.main_block / residual_mapping (models.py:295-310), forward and backward.
//
// Affine convention shared by every entry point: scale = scale_ptr[row*row_stride + c], shift likewise,
// row = cls[b] (class index of sample b) or 0 when cls == NULL.  CBN: scale_ptr = emb, shift_ptr = emb + C,
// row_stride = 2C (models.py:501).  Plain BN: scale_ptr = weight, shift_ptr = bias, row_stride = 0, cls = NULL.
#include "common.cuh"
#include "../../include/spyramid_b200.h"

extern void spyr_count_launch();

namespace {


// p -> (p / d, p % d); every map width here is a power of two, which turns the division into a shift
__device__ __forceinline__ int pow2_shift(int d) { return (d & (d - 1)) == 0 ? __ffs(d) - 1 : -1; }
__device__ __forceinline__ void divmod(int p, int d, int shift, int& q, int& r) {
  if (shift >= 0) {
    q = p >> shift;
    r = p & (d - 1);
  } else {
    q = p / d;
    r = p - q * d;
  }
}

// ATen upsample_bilinear2d, align_corners=True: src = dst * (in-1)/(out-1) in float
struct Lerp {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lerp lerp_src(int o, int in, float scale) {
  const float r = scale * (float)o;
  Lerp L;
  L.i0 = (int)r;
  L.i1 = L.i0 + (L.i0 < in - 1 ? 1 : 0);
  L.w1 = r - (float)L.i0;
  L.w0 = 1.f - L.w1;
  return L;
}
// hi-res positions that read low-res index j, with their weights (transpose of lerp_src)
struct Gather {
  int n;
  int o[8];
  float w[8];
};
__device__ __forceinline__ Gather gather_src(int j, int in, float scale) {
  Gather G;
  G.n = 0;
  const int out = 2 * in;
  int lo = (int)floorf((float)(j - 1) / scale) - 1;
  int hi = (int)ceilf((float)(j + 1) / scale) + 1;
  if (lo < 0) lo = 0;
  if (hi > out - 1) hi = out - 1;
  for (int o = lo; o <= hi; ++o) {
    const Lerp L = lerp_src(o, in, scale);
    const float w = (L.i0 == j ? L.w0 : 0.f) + (L.i1 == j ? L.w1 : 0.f);
    if (w != 0.f && G.n < 8) {
      G.o[G.n] = o;
      G.w[G.n] = w;
      ++G.n;
    }
  }
  // x2 with align_corners gives 3 or 4 taps per low-resolution index: pad to GATHER_TAPS with zero-weight repeats of a
  // valid position so that the gather loops have a fixed trip count (16 independent loads instead of a rolled loop)
  for (int k = G.n; k < 8; ++k) {
    G.o[k] = G.n > 0 ? G.o[G.n - 1] : 0;
    G.w[k] = 0.f;
  }
  return G;
}
constexpr int GATHER_TAPS = 4;

// Gather tables of one kernel launch in shared memory: rows [0,H) then columns [0,W).  gather_src is ~100 instructions
// with local arrays; evaluated once per table entry instead of twice per output vector it stops being the cost of the
// transposed-upsample kernels (up2_bwd: 146 -> 40 us on the 64-channel 256x256 gradient).
constexpr int GATHER_MAX = 256;  // H + W of the low-resolution map
__device__ __forceinline__ void build_gather_tables(Gather* tab, int H, int W, float shs, float sws) {
  for (int i = threadIdx.x; i < H + W; i += blockDim.x) tab[i] = i < H ? gather_src(i, H, shs) : gather_src(i - H, W, sws);
  __syncthreads();
}

template <bool S>
__device__ __forceinline__ void up2_load(const ActT<S>& x, int b, int oh, int ow, int H, int W, int cg, int c, float sh,
                                         float sw, float* v) {
  const Lerp Lh = lerp_src(oh, H, sh), Lw = lerp_src(ow, W, sw);
  float a[8], bq[8], cq[8], d[8];
  const size_t base = (size_t)b * H * W;
  ld8(x + ((base + (size_t)Lh.i0 * W + Lw.i0) * cg + c) * 8, a);
  ld8(x + ((base + (size_t)Lh.i0 * W + Lw.i1) * cg + c) * 8, bq);
  ld8(x + ((base + (size_t)Lh.i1 * W + Lw.i0) * cg + c) * 8, cq);
  ld8(x + ((base + (size_t)Lh.i1 * W + Lw.i1) * cg + c) * 8, d);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = Lh.w0 * (Lw.w0 * a[j] + Lw.w1 * bq[j]) + Lh.w1 * (Lw.w0 * cq[j] + Lw.w1 * d[j]);
}

// ---------------------------------------------------------------------------------------------
// statistics
// ---------------------------------------------------------------------------------------------
// The kernels below are templated on their mode: one body per mode keeps the plain same-resolution cases at ~40
// registers (full occupancy) instead of inheriting the register count of the interpolating variants.
template <int up2, bool S>
__global__ void bn_stats_kernel(const ActT<S> x, int B, int H, int W, int cg, const ActT<S> xu_out,
                                double* __restrict__ partials) {
  extern __shared__ float sh[];  // [prows][2][C]
  const int C = cg * 8;
  const int c = threadIdx.x % cg;
  const int pr = threadIdx.x / cg;
  const int prows = blockDim.x / cg;
  const int OH = up2 ? 2 * H : H, OW = up2 ? 2 * W : W;
  const float shs = up2 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
  const float sws = up2 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  const long long npix = (long long)B * OH * OW;
  const int sh_w = pow2_shift(OW), sh_h = pow2_shift(OH);
  float s1[8], s2[8], v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (pr < prows)
#pragma unroll(up2 ? 2 : 4)  // independent 16-byte loads in flight per thread: 4 (plain) or 2 x 4 corners (interpolating)
    for (long long p = (long long)blockIdx.x * prows + pr; p < npix; p += (long long)gridDim.x * prows) {
      if (up2) {
        int rowi, ow, b, oh;
        divmod((int)p, OW, sh_w, rowi, ow);
        divmod(rowi, OH, sh_h, b, oh);
        up2_load(x, b, oh, ow, H, W, cg, c, shs, sws, v);
        if (!xu_out.null()) {
          // materialise up2(x) once; the statistics are those of the stored values the later passes read
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = stored_value(v[j], xu_out);
          st8(xu_out + (p * cg + c) * 8, v);
        }
      } else {
        ld8(x + (p * cg + c) * 8, v);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += v[j];
        s2[j] += v[j] * v[j];
      }
    }
  if (pr < prows) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sh[(pr * 2 + 0) * C + c * 8 + j] = s1[j];
      sh[(pr * 2 + 1) * C + c * 8 + j] = s2[j];
    }
  }
  __syncthreads();
  if (partials == nullptr) return;  // plain materialisation of up2(x)
  // one partial vector per block; bn_finalize adds them in block order (no atomics: bit-reproducible)
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < prows; ++r) s += (double)sh[r * 2 * C + i];
    partials[(size_t)blockIdx.x * 2 * C + i] = s;
  }
}

__device__ __forceinline__ double warp_tree_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// mean/rstd from the block partials of bn_stats (train: one warp per channel adds the nb partial sums lane-strided, then
// a fixed xor tree) or from the running buffers (eval); running-stat update as nn.BatchNorm2d
__global__ void bn_finalize_kernel(const double* __restrict__ partials, int nb, double count, int C, float eps,
                                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ nbt, float* __restrict__ mean_rstd, int training) {
  const int c = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  double s1 = 0.0, s2 = 0.0;
  if (c < C && training) {
    // up to 296 partial vectors, ~10 per lane, each load a different line: keep several in flight (the adds stay in order)
#pragma unroll 5
    for (int b = lane; b < nb; b += 32) {
      s1 += __ldcg(partials + (size_t)b * 2 * C + c);
      s2 += __ldcg(partials + (size_t)b * 2 * C + C + c);
    }
  }
  s1 = warp_tree_d(s1);
  s2 = warp_tree_d(s2);
  if (lane != 0) return;
  if (c < C) {
    if (training) {
      const double mean = s1 / count;
      double var = s2 / count - mean * mean;
      if (var < 0.0) var = 0.0;
      mean_rstd[c] = (float)mean;
      mean_rstd[C + c] = (float)(1.0 / sqrt(var + (double)eps));
      if (running_mean != nullptr) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    } else {
      mean_rstd[c] = running_mean[c];
      mean_rstd[C + c] = rsqrtf(running_var[c] + eps);
    }
  }
  if (training && nbt != nullptr && c == 0) *nbt += 1;
}

// ---------------------------------------------------------------------------------------------
// forward apply.  mode 0: a = lrelu(aff(x))            (same resolution)
//                 mode 1: a = up2(lrelu(aff(x))), xu = up2(x)   (generator block, models.py:295-298,308)
//                 mode 2: a = lrelu(aff(up2(x)))       (final block: upsample -> BN -> LeakyReLU, models.py:52-54)
// ---------------------------------------------------------------------------------------------
// Channel-group-stationary: a thread keeps the affine parameters of its 8 channels (of ONE sample: blockIdx.y) in
// registers and walks over output pixels, so the per-element work is one 16-byte load (four for the bilinear modes)
// and one or two 16-byte stores.
template <int mode, bool S>
__global__ void bn_act_kernel(const ActT<S> x, const float* __restrict__ mean_rstd,
                              const float* __restrict__ scale_ptr, const float* __restrict__ shift_ptr, int row_stride,
                              const int* __restrict__ cls, float slope, const ActT<S> out_a,
                              const ActT<S> out_xu, int H, int W, int cg) {
  const int C = cg * 8;
  const int OH = mode ? 2 * H : H, OW = mode ? 2 * W : W;
  const int b = blockIdx.y;
  const int c = threadIdx.x % cg;
  const int pr = threadIdx.x / cg;
  const int prows = blockDim.x / cg;
  if (pr >= prows) return;
  const int row = cls != nullptr ? cls[b] : 0;
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = c * 8 + j;
    const float r = mean_rstd[C + ch];
    const float s = scale_ptr[(size_t)row * row_stride + ch] * r;
    sc[j] = s;
    sf[j] = shift_ptr[(size_t)row * row_stride + ch] - mean_rstd[ch] * s;
  }
  const int npix = OH * OW;
  const size_t in_base = (size_t)b * H * W, out_base = (size_t)b * npix;
  const float shs = mode ? (float)(H - 1) / (float)(OH - 1) : 0.f, sws = mode ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  const int sh_w = pow2_shift(OW);
#pragma unroll(mode == 0 ? 4 : 2)
  for (int p = blockIdx.x * prows + pr; p < npix; p += gridDim.x * prows) {
    float v[8];
    const size_t o = ((out_base + p) * cg + c) * 8;
    if (mode == 0) {
      ld8(x + o, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = lrelu_f(v[j] * sc[j] + sf[j], slope);
      st8(out_a + o, v);
      continue;
    }
    int oh, ow;
    divmod(p, OW, sh_w, oh, ow);
    const Lerp Lh = lerp_src(oh, H, shs), Lw = lerp_src(ow, W, sws);
    float q[4][8];
    ld8(x + ((in_base + (size_t)Lh.i0 * W + Lw.i0) * cg + c) * 8, q[0]);
    ld8(x + ((in_base + (size_t)Lh.i0 * W + Lw.i1) * cg + c) * 8, q[1]);
    ld8(x + ((in_base + (size_t)Lh.i1 * W + Lw.i0) * cg + c) * 8, q[2]);
    ld8(x + ((in_base + (size_t)Lh.i1 * W + Lw.i1) * cg + c) * 8, q[3]);
    const float w00 = Lh.w0 * Lw.w0, w01 = Lh.w0 * Lw.w1, w10 = Lh.w1 * Lw.w0, w11 = Lh.w1 * Lw.w1;
    if (!out_xu.null()) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = Lh.w0 * (Lw.w0 * q[0][j] + Lw.w1 * q[1][j]) + Lh.w1 * (Lw.w0 * q[2][j] + Lw.w1 * q[3][j]);
      st8(out_xu + o, v);
    }
    (void)w00; (void)w01; (void)w10; (void)w11;
    if (mode == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) q[k][j] = lrelu_f(q[k][j] * sc[j] + sf[j], slope);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = Lh.w0 * (Lw.w0 * q[0][j] + Lw.w1 * q[1][j]) + Lh.w1 * (Lw.w0 * q[2][j] + Lw.w1 * q[3][j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float u = Lh.w0 * (Lw.w0 * q[0][j] + Lw.w1 * q[1][j]) + Lh.w1 * (Lw.w0 * q[2][j] + Lw.w1 * q[3][j]);
        v[j] = lrelu_f(u * sc[j] + sf[j], slope);
      }
    }
    st8(out_a + o, v);
  }
}

// ---------------------------------------------------------------------------------------------
// backward, pass 1: per-(sample, channel) sums S[b][0][c] = sum_p gy, S[b][1][c] = sum_p gy * xhat.
//   mode 0: g is already d/dy (same resolution, LeakyReLU gate applied by the conv epilogue)
//   mode 1: g is d/da at 2H x 2W with a = up2(lrelu(y)):  gy = up2^T(g) * lrelu'(y), written to gy_out
//   mode 2: g is d/d(pre-LeakyReLU) at 2H x 2W (gate applied upstream): gy = up2^T(g), written to gy_out
//   mode 3: g is d/dy at 2H x 2W and the normalised tensor is up2(x) (final block): reduce at 2H x 2W
// ---------------------------------------------------------------------------------------------
template <int mode, bool S>
__global__ void bn_bwd_reduce_kernel(const ActT<S> g, const ActT<S> x,
                                     const float* __restrict__ mean_rstd, const float* __restrict__ scale_ptr,
                                     const float* __restrict__ shift_ptr, int row_stride, const int* __restrict__ cls,
                                     float slope, const ActT<S> gy_out, float* __restrict__ partials,
                                     float* __restrict__ Sout, unsigned int* tickets, int H, int W, int cg) {
  extern __shared__ float sh[];  // [prows][2][C]
  __shared__ Gather gtab[GATHER_MAX];
  const int C = cg * 8;
  const int b = blockIdx.y;
  const int c = threadIdx.x % cg;
  const int pr = threadIdx.x / cg;
  const int prows = blockDim.x / cg;
  const int row = cls != nullptr ? cls[b] : 0;
  if (mode == 1 || mode == 2)
    build_gather_tables(gtab, H, W, (float)(H - 1) / (float)(2 * H - 1), (float)(W - 1) / (float)(2 * W - 1));
  float mu[8], rs[8], sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = c * 8 + j;
    mu[j] = mean_rstd[ch];
    rs[j] = mean_rstd[C + ch];
    sc[j] = scale_ptr[(size_t)row * row_stride + ch];
    sf[j] = shift_ptr[(size_t)row * row_stride + ch];
  }
  const float shs = (float)(H - 1) / (float)(2 * H - 1), sws = (float)(W - 1) / (float)(2 * W - 1);
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  const int npix = (mode == 3) ? 4 * H * W : H * W;
  const int sh_w = pow2_shift(W), sh_2w = pow2_shift(2 * W);
  if (pr < prows)
#pragma unroll(mode == 0 ? 4 : 1)
  for (int p = blockIdx.x * prows + pr; p < npix; p += gridDim.x * prows) {
      const size_t off = (((size_t)b * npix + p) * cg + c) * 8;
      float xv[8], gy[8];
      if (mode == 3) {
        // statistics were taken on up2(x): reduce at the high resolution, x interpolated on the fly
        int oh, ow;
        divmod(p, 2 * W, sh_2w, oh, ow);
        up2_load(x, b, oh, ow, H, W, cg, c, shs, sws, xv);
        ld8(g + off, gy);
      } else if (mode == 0) {
        ld8(x + off, xv);
        ld8(g + off, gy);
      } else {
        ld8(x + off, xv);
        int h, w;
        divmod(p, W, sh_w, h, w);
        const Gather& Gh = gtab[h];
        const Gather& Gw = gtab[H + w];
#pragma unroll
        for (int j = 0; j < 8; ++j) gy[j] = 0.f;
        if (Gh.n <= GATHER_TAPS && Gw.n <= GATHER_TAPS) {
#pragma unroll
          for (int a = 0; a < GATHER_TAPS; ++a) {
            float raw[GATHER_TAPS][8];
#pragma unroll
            for (int e = 0; e < GATHER_TAPS; ++e)
              ld8_nc(g + ((((size_t)b * 2 * H + Gh.o[a]) * 2 * W + Gw.o[e]) * cg + c) * 8, raw[e]);
#pragma unroll
            for (int e = 0; e < GATHER_TAPS; ++e) {
              const float wt = Gh.w[a] * Gw.w[e];
#pragma unroll
              for (int j = 0; j < 8; ++j) gy[j] += wt * raw[e][j];
            }
          }
        } else {
          for (int a = 0; a < Gh.n; ++a)
            for (int e = 0; e < Gw.n; ++e) {
              float gv[8];
              ld8(g + ((((size_t)b * 2 * H + Gh.o[a]) * 2 * W + Gw.o[e]) * cg + c) * 8, gv);
              const float wt = Gh.w[a] * Gw.w[e];
#pragma unroll
              for (int j = 0; j < 8; ++j) gy[j] += wt * gv[j];
            }
        }
        if (mode == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float y = sc[j] * ((xv[j] - mu[j]) * rs[j]) + sf[j];
            if (!(y > 0.f)) gy[j] *= slope;
          }
        }
        // the reduction uses the stored value that pass 2 will read back
#pragma unroll
        for (int j = 0; j < 8; ++j) gy[j] = stored_value(gy[j], gy_out);
        st8(gy_out + off, gy);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += gy[j];
        s2[j] += gy[j] * ((xv[j] - mu[j]) * rs[j]);
      }
    }
  if (pr < prows) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sh[(pr * 2 + 0) * C + c * 8 + j] = s1[j];
      sh[(pr * 2 + 1) * C + c * 8 + j] = s2[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < prows; ++r) s += sh[r * 2 * C + i];
    partials[((size_t)b * gridDim.x + blockIdx.x) * 2 * C + i] = s;
  }
  // the last block of THIS sample adds the sample's gridDim.x partial vectors in block order (one ticket per sample, so
  // the B tails run in parallel): Sout[b][0][c] = sum gy, Sout[b][1][c] = sum gy * xhat -- no atomics on the data
  if (spyr_last_block(tickets + b, gridDim.x))
    spyr_sum_partials<float>(partials + (size_t)b * gridDim.x * 2 * C, (int)gridDim.x, 2 * C,
                             [&](int i, float total) { Sout[(size_t)b * 2 * C + i] = total; });
}

// pass 1b: channel means M[0][c] = (1/N) sum_b scale[b,c] S1[b,c], M[1][c] likewise with S2.  Read-only loop: the loads of
// the B samples are independent and pipeline.  (Together with the parameter gradients below, whose read-modify-writes
// chain through memory, this kernel took 20 us on 4 CTAs and sat between bn_bwd_reduce and bn_bwd_apply eleven times per
// step; the parameter gradients are leaves of the backward pass and now run in their own kernel, off that chain.)
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ Sx, int B, int C, float count,
                                       const float* __restrict__ scale_ptr, int row_stride, const int* __restrict__ cls,
                                       float* __restrict__ M) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float m1 = 0.f, m2 = 0.f;
#pragma unroll 4
  for (int b = 0; b < B; ++b) {
    const int row = cls != nullptr ? __ldg(cls + b) : 0;
    const float sc = __ldg(scale_ptr + (size_t)row * row_stride + c);
    const float a1 = __ldg(Sx + ((size_t)b * 2 + 0) * C + c), a2 = __ldg(Sx + ((size_t)b * 2 + 1) * C + c);
    m1 += sc * a1;
    m2 += sc * a2;
  }
  M[c] = m1 / count;
  M[C + c] = m2 / count;
}
// affine parameter gradients: d_scale[row(b)][c] += S2[b,c], d_shift[row(b)][c] += S1[b,c].  One thread owns channel c of
// every row and visits the samples in order: plain read-modify-write, fixed order (two samples of one class hit the same
// row).  Without class rows everything lands in row 0: summed in registers, one update.
__global__ void bn_bwd_params_kernel(const float* __restrict__ Sx, int B, int C, int row_stride,
                                     const int* __restrict__ cls, float* __restrict__ d_scale,
                                     float* __restrict__ d_shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (cls == nullptr) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
    for (int b = 0; b < B; ++b) {
      s1 += __ldg(Sx + ((size_t)b * 2 + 0) * C + c);
      s2 += __ldg(Sx + ((size_t)b * 2 + 1) * C + c);
    }
    d_scale[c] += s2;
    d_shift[c] += s1;
    return;
  }
  for (int b = 0; b < B; ++b) {
    const int row = cls[b];
    d_scale[(size_t)row * row_stride + c] += Sx[((size_t)b * 2 + 1) * C + c];
    d_shift[(size_t)row * row_stride + c] += Sx[((size_t)b * 2 + 0) * C + c];
  }
}

// pass 2: gx = rstd * (scale * gy - M1 - xhat * M2) (+ residual).  x_up2: gy/gx live at 2H x 2W and xhat is taken
// from up2(x) (final block, where the statistics are those of the upsampled tensor)
template <int x_up2, bool S>
__global__ void bn_bwd_apply_kernel(const ActT<S> gy, const ActT<S> x,
                                    const float* __restrict__ mean_rstd, const float* __restrict__ scale_ptr,
                                    int row_stride, const int* __restrict__ cls, const float* __restrict__ M,
                                    const ActT<S> residual, const ActT<S> gx, int H, int W, int cg) {
  const int C = cg * 8;
  const int OH = x_up2 ? 2 * H : H, OW = x_up2 ? 2 * W : W;
  const int b = blockIdx.y;
  const int c = threadIdx.x % cg;
  const int pr = threadIdx.x / cg;
  const int prows = blockDim.x / cg;
  if (pr >= prows) return;
  const int row = cls != nullptr ? cls[b] : 0;
  float mu[8], rs[8], sc[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = c * 8 + j;
    mu[j] = mean_rstd[ch];
    rs[j] = mean_rstd[C + ch];
    sc[j] = scale_ptr[(size_t)row * row_stride + ch];
    m1[j] = M[ch];
    m2[j] = M[C + ch];
  }
  const int npix = OH * OW;
  const size_t base = (size_t)b * npix;
  const float shs = x_up2 ? (float)(H - 1) / (float)(OH - 1) : 0.f, sws = x_up2 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  const int sh_w = pow2_shift(OW);
#pragma unroll(x_up2 ? 1 : 4)
  for (int p = blockIdx.x * prows + pr; p < npix; p += gridDim.x * prows) {
    const size_t off = ((base + p) * cg + c) * 8;
    float g[8], xv[8], o[8];
    ld8(gy + off, g);
    if (x_up2) {
      int oh, ow;
      divmod(p, OW, sh_w, oh, ow);
      up2_load(x, b, oh, ow, H, W, cg, c, shs, sws, xv);
    } else {
      ld8(x + off, xv);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (xv[j] - mu[j]) * rs[j];
      o[j] = rs[j] * (sc[j] * g[j] - m1[j] - xh * m2[j]);
    }
    if (!residual.null()) {
      float rv[8];
      ld8(residual + off, rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += rv[j];
    }
    st8(gx + off, o);
  }
}

// plain transposed bilinear x2 (align_corners=True): g_lo = up2^T(g_hi)  (skip path of the generator block)
template <bool S>
__global__ void up2_bwd_kernel(const ActT<S> g, const ActT<S> out, int B, int H, int W, int cg) {
  __shared__ Gather gtab[GATHER_MAX];
  build_gather_tables(gtab, H, W, (float)(H - 1) / (float)(2 * H - 1), (float)(W - 1) / (float)(2 * W - 1));
  const long long n = (long long)B * H * W * cg;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned)idx % (unsigned)cg);  // 32-bit index split: element counts are < 2^31 (checked at launch)
    int t = (int)((unsigned)idx / (unsigned)cg);
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int b = (int)(t / H);
    const Gather& Gh = gtab[h];
    const Gather& Gw = gtab[H + w];
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (Gh.n <= GATHER_TAPS && Gw.n <= GATHER_TAPS) {
#pragma unroll
      for (int a = 0; a < GATHER_TAPS; ++a) {
        float raw[GATHER_TAPS][8];
#pragma unroll
        for (int e = 0; e < GATHER_TAPS; ++e)
          ld8_nc(g + ((((size_t)b * 2 * H + Gh.o[a]) * 2 * W + Gw.o[e]) * cg + c) * 8, raw[e]);
#pragma unroll
        for (int e = 0; e < GATHER_TAPS; ++e) {
          const float wt = Gh.w[a] * Gw.w[e];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += wt * raw[e][j];
        }
      }
    } else {
      for (int a = 0; a < Gh.n; ++a)
        for (int e = 0; e < Gw.n; ++e) {
          float gv[8];
          ld8(g + ((((size_t)b * 2 * H + Gh.o[a]) * 2 * W + Gw.o[e]) * cg + c) * 8, gv);
          const float wt = Gh.w[a] * Gw.w[e];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += wt * gv[j];
        }
    }
    st8(out + idx * 8, acc);
  }
}

// class index of each one-hot row (class_id.argmax(dim=-1), models.py:151,501); first maximum wins like torch
// One warp per row: every lane keeps the first maximum of its strided elements, a butterfly picks the larger value and, on
// ties, the smaller index.  (A single thread scanning the 365 entries took 12-24 us at the head of every G / D forward.)
template <typename T>
__global__ void argmax_rows_kernel(const T* __restrict__ onehot, int n, int* __restrict__ out) {
  const int b = blockIdx.x;
  const int lane = threadIdx.x;
  const T* row = onehot + (size_t)b * n;
  T m = row[0];
  int best = 0;
  for (int i = lane; i < n; i += 32) {
    const T v = row[i];
    if (v > m || (v == m && i < best)) {
      m = v;
      best = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T om = __shfl_xor_sync(0xffffffffu, m, o);
    const int ob = __shfl_xor_sync(0xffffffffu, best, o);
    if (om > m || (om == m && ob < best)) {
      m = om;
      best = ob;
    }
  }
  if (lane == 0) out[b] = best;
}

struct Threads {
  int prows, threads;
};
inline Threads pick_threads(int cg, int max_threads);
// grid of bn_stats (= number of partial vectors bn_finalize adds): a function of the problem shape only
inline int bn_stats_blocks(long long npix, int cg) {
  const int prows = 256 / cg < 1 ? 1 : 256 / cg;
  long long want = (npix + prows * 16 - 1) / (prows * 16);
  return (int)(want < 1 ? 1 : (want > SPYR_REDUCE_BLOCKS ? SPYR_REDUCE_BLOCKS : want));
}
// x-grid of bn_bwd_reduce (partial vectors per sample): B * gx <= 4 * SPYR_REDUCE_BLOCKS
inline int bn_bwd_blocks(int npix_per_image, int cg, int B) {
  const int prows = 256 / cg < 1 ? 1 : 256 / cg;
  int gx = (npix_per_image + prows * 8 - 1) / (prows * 8);
  const int cap = (4 * SPYR_REDUCE_BLOCKS) / B;
  if (gx > cap) gx = cap;
  return gx < 1 ? 1 : gx;
}
inline Threads pick_threads(int cg, int max_threads) {
  Threads t;
  t.prows = max_threads / cg;
  if (t.prows < 1) t.prows = 1;
  t.threads = t.prows * cg;
  return t;
}

}  // namespace

#define SPYR_C8(C) SPYR_REQUIRE((C) > 0 && (C) % 8 == 0, "%s: channel count %d must be a multiple of 8", __func__, (int)(C))

static int bn_stats_launch(const void* x, int B, int H, int W, int C, int up2, void* xu_out, double* partials,
                           void* stream_);

extern "C" int spyr_bn_stats(const void* x, int B, int H, int W, int C, int up2, double* partials, void* stream) {
  return bn_stats_launch(x, B, H, W, C, up2, nullptr, partials, stream);
}
extern "C" int spyr_up2_stats(const void* x, int B, int H, int W, int C, void* xu_out, double* partials, void* stream) {
  SPYR_REQUIRE(xu_out != nullptr, "up2_stats: xu_out is NULL");
  return bn_stats_launch(x, B, H, W, C, 1, xu_out, partials, stream);
}
static int bn_stats_launch(const void* x, int B, int H, int W, int C, int up2, void* xu_out, double* partials,
                           void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPYR_C8(C);
  const int cg = C / 8;
  SPYR_REQUIRE(cg <= 256, "bn_stats: C=%d too large", C);
  const Threads t = pick_threads(cg, 256);
  const long long npix = (long long)B * H * W * (up2 ? 4 : 1);
  // without statistics (plain materialisation of up2(x)) nothing is reduced: the grid is not bounded by the partials
  long long want = (npix + t.prows * 16 - 1) / (t.prows * 16);
  const int grid = partials != nullptr ? bn_stats_blocks(npix, cg) : (int)(want < 1 ? 1 : (want > 1184 ? 1184 : want));
  const Act xa = make_act(x, (long long)B * H * W * C), xu = make_act(xu_out, npix * C);
  if (up2)
    SPYR_WITH_SPLIT(bn_stats_kernel<1, kS><<<grid, t.threads, (size_t)t.prows * 2 * C * 4, stream>>>(xa, B, H, W, cg, xu, partials));
  else
    SPYR_WITH_SPLIT(bn_stats_kernel<0, kS><<<grid, t.threads, (size_t)t.prows * 2 * C * 4, stream>>>(xa, B, H, W, cg, xu, partials));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_bn_finalize(const double* partials, double count, int C, float eps, float momentum, float* running_mean,
                                float* running_var, long long* num_batches_tracked, float* mean_rstd, int training,
                                void* stream) {
  SPYR_REQUIRE(training || (running_mean && running_var), "bn_finalize: eval mode needs running statistics");
  SPYR_REQUIRE(!training || partials != nullptr, "bn_finalize: training mode needs the partial sums of spyr_bn_stats");
  SPYR_C8(C);
  const int nb = training ? bn_stats_blocks((long long)(count + 0.5), C / 8) : 0;  // the grid spyr_bn_stats used
  bn_finalize_kernel<<<ceil_div(C, 8), 256, 0, (cudaStream_t)stream>>>(partials, nb, count, C, eps, momentum, running_mean,
                                                                       running_var, num_batches_tracked, mean_rstd,
                                                                       training);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_bn_act(const void* x, const float* mean_rstd, const float* scale_ptr, const float* shift_ptr,
                           int row_stride, const int* cls, float slope, int mode, void* out_a, void* out_xu, int B, int H,
                           int W, int C, void* stream) {
  SPYR_C8(C);
  SPYR_REQUIRE(mode >= 0 && mode <= 2, "bn_act: bad mode %d", mode);
  SPYR_REQUIRE(mode == 0 || (H > 1 && W > 1), "bn_act: upsampling needs H,W > 1");
  const int cg = C / 8;
  SPYR_REQUIRE(cg <= 256, "bn_act: C=%d too large", C);
  const Threads t = pick_threads(cg, 256);
  const int npix = H * W * (mode ? 4 : 1);
  int gx = (npix + t.prows * 4 - 1) / (t.prows * 4);
  const int cap = (2368 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  const long long nin = (long long)B * H * W * C, nout = (long long)B * npix * C;
#define SPYR_BN_ACT(M)                                                                                                  \
  SPYR_WITH_SPLIT(bn_act_kernel<M, kS><<<dim3(gx, B), t.threads, 0, (cudaStream_t)stream>>>(make_act(x, nin), mean_rstd, scale_ptr, shift_ptr, \
                                                                        row_stride, cls, slope, make_act(out_a, nout),   \
                                                                        make_act(out_xu, nout), H, W, cg))
  if (mode == 0) SPYR_BN_ACT(0);
  else if (mode == 1) SPYR_BN_ACT(1);
  else SPYR_BN_ACT(2);
#undef SPYR_BN_ACT
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_bn_bwd_reduce(const void* g, const void* x, const float* mean_rstd, const float* scale_ptr,
                                  const float* shift_ptr, int row_stride, const int* cls, float slope, int mode,
                                  void* gy_out, float* S, int B, int H, int W, int C, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPYR_C8(C);
  SPYR_REQUIRE(mode == 0 || mode == 3 || gy_out != nullptr, "bn_bwd_reduce: modes 1/2 need gy_out");
  SPYR_REQUIRE(mode >= 0 && mode <= 3, "bn_bwd_reduce: bad mode %d", mode);
  SPYR_REQUIRE((mode != 1 && mode != 2) || H + W <= GATHER_MAX, "bn_bwd_reduce: H + W = %d exceeds the gather table", H + W);
  const int cg = C / 8;
  SPYR_REQUIRE(cg <= 256, "bn_bwd_reduce: C=%d too large", C);
  SPYR_REQUIRE(S != nullptr, "bn_bwd_reduce: the partial-sum buffer is NULL");
  const Threads t = pick_threads(cg, 256);
  const int npix = H * W * (mode == 3 ? 4 : 1);
  const int gx = bn_bwd_blocks(npix, cg, B);  // B * gx partial vectors of 2C floats: SPYR_BN_BWD_PARTIAL_BYTES(C)
  SPYR_REQUIRE(gx * B <= 4 * SPYR_REDUCE_BLOCKS, "bn_bwd_reduce: batch %d exceeds %d", B, 4 * SPYR_REDUCE_BLOCKS);
  dim3 grid(gx, B);
  unsigned int* tickets = spyr_next_tickets(B);  // one per sample
  SPYR_REQUIRE(tickets != nullptr, "bn_bwd_reduce: batch %d exceeds the ticket pool", B);
  // H, W are x's dims; g lives at 2H x 2W in modes 1, 2 and 3, gy_out (modes 1, 2) at H x W
  const long long nx = (long long)B * H * W * C, ng = (mode == 1 || mode == 2) ? 4 * nx : (mode == 3 ? 4 * nx : nx);
#define SPYR_BN_RED(M)                                                                  \
  SPYR_WITH_SPLIT(bn_bwd_reduce_kernel<M, kS><<<grid, t.threads, (size_t)t.prows * 2 * C * 4, stream>>>(     \
      make_act(g, ng), make_act(x, nx), mean_rstd, scale_ptr, shift_ptr, row_stride, cls, slope, make_act(gy_out, nx),  \
      S + (size_t)B * 2 * C, S, tickets, H, W, cg))
  if (mode == 0) SPYR_BN_RED(0);
  else if (mode == 1) SPYR_BN_RED(1);
  else if (mode == 2) SPYR_BN_RED(2);
  else SPYR_BN_RED(3);
#undef SPYR_BN_RED
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_bn_bwd_finalize(const float* S, int B, int C, float count, const float* scale_ptr, int row_stride,
                                    const int* cls, float* M, float* d_scale, float* d_shift, void* stream) {
  bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(S, B, C, count, scale_ptr, row_stride, cls, M);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  if (d_scale != nullptr) return spyr_bn_bwd_params(S, B, C, row_stride, cls, d_scale, d_shift, stream);
  return 0;
}
extern "C" int spyr_bn_bwd_params(const float* S, int B, int C, int row_stride, const int* cls, float* d_scale,
                                  float* d_shift, void* stream) {
  SPYR_REQUIRE(S != nullptr && d_scale != nullptr && d_shift != nullptr, "bn_bwd_params: bad arguments");
  bn_bwd_params_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(S, B, C, row_stride, cls, d_scale, d_shift);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_bn_bwd_apply(const void* gy, const void* x, const float* mean_rstd, const float* scale_ptr,
                                 int row_stride, const int* cls, const float* M, const void* residual, void* gx, int B,
                                 int H, int W, int C, int x_up2, void* stream) {
  SPYR_C8(C);
  const int cg = C / 8;
  SPYR_REQUIRE(cg <= 256, "bn_bwd_apply: C=%d too large", C);
  const Threads t = pick_threads(cg, 256);
  const int npix = H * W * (x_up2 ? 4 : 1);
  int gxx = (npix + t.prows * 4 - 1) / (t.prows * 4);
  const int cap = (2368 + B - 1) / B;
  if (gxx > cap) gxx = cap;
  if (gxx < 1) gxx = 1;
  const long long nx = (long long)B * H * W * C, ng = (long long)B * npix * C;
  if (x_up2)
    SPYR_WITH_SPLIT(bn_bwd_apply_kernel<1, kS><<<dim3(gxx, B), t.threads, 0, (cudaStream_t)stream>>>(
        make_act(gy, ng), make_act(x, nx), mean_rstd, scale_ptr, row_stride, cls, M, make_act(residual, ng), make_act(gx, ng),
        H, W, cg));
  else
    SPYR_WITH_SPLIT(bn_bwd_apply_kernel<0, kS><<<dim3(gxx, B), t.threads, 0, (cudaStream_t)stream>>>(
        make_act(gy, ng), make_act(x, nx), mean_rstd, scale_ptr, row_stride, cls, M, make_act(residual, ng), make_act(gx, ng),
        H, W, cg));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_up2_bwd(const void* g_hi, void* g_lo, int B, int H, int W, int C, void* stream) {
  SPYR_C8(C);
  SPYR_REQUIRE(H > 1 && W > 1, "up2_bwd: H,W must be > 1");
  const long long n = (long long)B * H * W * (C / 8);
  SPYR_N32(n);
  SPYR_REQUIRE(H + W <= GATHER_MAX, "up2_bwd: H + W = %d exceeds the gather table (%d)", H + W, GATHER_MAX);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;  // the tables are built once per block: a few blocks per SM, grid-stride loop
  SPYR_WITH_SPLIT(up2_bwd_kernel<kS><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(make_act(g_hi, n * 32), make_act(g_lo, n * 8), B, H, W, C / 8));
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
extern "C" int spyr_argmax_rows(const void* onehot, int is_int64, int B, int n, int* out, void* stream) {
  if (is_int64)
    argmax_rows_kernel<long long><<<B, 32, 0, (cudaStream_t)stream>>>((const long long*)onehot, n, out);
  else
    argmax_rows_kernel<float><<<B, 32, 0, (cudaStream_t)stream>>>((const float*)onehot, n, out);
  spyr_count_launch();
  SPYR_LAUNCH_CHECK();
  return 0;
}
