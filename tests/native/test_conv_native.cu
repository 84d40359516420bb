// Standalone GPU check of the tcgen05 conv kernels through the C-ABI against a plain CPU loop.
// Build: see Makefile target `native_tests`.  Run on a B200: ./build/test_conv_native [wgrad LBO SBO]
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/spyramid_b200.h"

static float bf16_round(float f) { return __bfloat162float(__float2bfloat16(f)); }
static uint32_t rng_state = 12345u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

struct Dev {
  void* p = nullptr;
  size_t n = 0;
  void alloc(size_t bytes) {
    CK(cudaMalloc(&p, bytes));
    n = bytes;
    CK(cudaMemset(p, 0, bytes));
  }
  ~Dev() {
    if (p) cudaFree(p);
  }
};

static std::vector<__nv_bfloat16> to_bf16(const std::vector<float>& v) {
  std::vector<__nv_bfloat16> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16(v[i]);
  return o;
}

static double rel_l2(const std::vector<float>& a, const std::vector<float>& b) {
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    num += (double)(a[i] - b[i]) * (a[i] - b[i]);
    den += (double)b[i] * b[i];
  }
  return sqrt(num / (den + 1e-30));
}

// fprop reference: x NHWC, w [tap][cout][cin]
static void ref_conv(const std::vector<float>& x, const std::vector<float>& w, int B, int H, int W, int Cin, int Cout,
                     int ks, std::vector<float>& y) {
  const int pad = ks / 2;
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int ww = 0; ww < W; ++ww)
        for (int co = 0; co < Cout; ++co) {
          double acc = 0;
          for (int t = 0; t < ks * ks; ++t) {
            const int hh = h + t / ks - pad, wx = ww + t % ks - pad;
            if (hh < 0 || hh >= H || wx < 0 || wx >= W) continue;
            const float* xp = &x[(((size_t)b * H + hh) * W + wx) * Cin];
            const float* wp = &w[((size_t)t * Cout + co) * Cin];
            for (int ci = 0; ci < Cin; ++ci) acc += (double)xp[ci] * wp[ci];
          }
          y[(((size_t)b * H + h) * W + ww) * Cout + co] += (float)acc;
        }
}

static int test_fprop(int B, int H, int W, int Cin, int Cout, int ks, int nsrc, bool epi, int splits, int block_n) {
  std::vector<float> xs[3], ws[3];
  const int cins[3] = {Cin, Cin > 64 ? Cin / 2 : Cin, 64};
  const int kss[3] = {ks, 3, 1};
  std::vector<float> yref((size_t)B * H * W * Cout, 0.f);
  Dev dx[3], dw[3];
  spyr_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.B = B; d.H = H; d.W = W; d.Cout = Cout; d.nsrc = nsrc;
  for (int s = 0; s < nsrc; ++s) {
    xs[s].resize((size_t)B * H * W * cins[s]);
    ws[s].resize((size_t)kss[s] * kss[s] * Cout * cins[s]);
    for (auto& v : xs[s]) v = bf16_round(frand());
    for (auto& v : ws[s]) v = bf16_round(frand() * 0.05f);
    ref_conv(xs[s], ws[s], B, H, W, cins[s], Cout, kss[s], yref);
    auto xb = to_bf16(xs[s]);
    auto wb = to_bf16(ws[s]);
    dx[s].alloc(xb.size() * 2);
    dw[s].alloc(wb.size() * 2);
    CK(cudaMemcpy(dx[s].p, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw[s].p, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
    d.src[s].x = dx[s].p; d.src[s].w = dw[s].p; d.src[s].cin = cins[s]; d.src[s].ksize = kss[s];
  }
  const size_t ny = yref.size();
  std::vector<float> bias(Cout), res(ny), dmask(ny), smask((size_t)B * H * W), sw((size_t)10 * Cout);
  std::vector<float> yact(ny);
  Dev dbias, dres, ddm, dsm, dsw, dyraw, dyact, dyf32;
  if (epi) {
    for (auto& v : bias) v = frand();
    for (auto& v : res) v = bf16_round(frand());
    for (auto& v : dmask) v = bf16_round(frand());
    for (size_t i = 0; i < smask.size(); ++i) {
      const int b = (int)(i / ((size_t)H * W));
      smask[i] = (b % 3 == 0) ? 1.f : ((b % 3 == 1) ? 0.f : (frand() > 0 ? 1.f : 0.f));
    }
    for (int c = 0; c < Cout; ++c) {
      float s = 0;
      for (int t = 0; t < 9; ++t) {
        sw[(size_t)t * Cout + c] = frand() * 0.1f;
        s += sw[(size_t)t * Cout + c];
      }
      sw[(size_t)9 * Cout + c] = s;
    }
    for (int b = 0; b < B; ++b)
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
          for (int c = 0; c < Cout; ++c) {
            const size_t i = (((size_t)b * H + h) * W + w) * Cout + c;
            float v = yref[i] + bias[c];
            for (int t = 0; t < 9; ++t) {
              const int hh = h + t / 3 - 1, wx = w + t % 3 - 1;
              if (hh < 0 || hh >= H || wx < 0 || wx >= W) continue;
              v += smask[((size_t)b * H + hh) * W + wx] * sw[(size_t)t * Cout + c];
            }
            if (!(dmask[i] > 0.f)) v *= 0.2f;
            v += res[i];
            yref[i] = v;
            yact[i] = v > 0 ? v : 0.2f * v;
          }
    dbias.alloc(Cout * 4); CK(cudaMemcpy(dbias.p, bias.data(), Cout * 4, cudaMemcpyHostToDevice));
    auto rb = to_bf16(res); dres.alloc(ny * 2); CK(cudaMemcpy(dres.p, rb.data(), ny * 2, cudaMemcpyHostToDevice));
    auto mb = to_bf16(dmask); ddm.alloc(ny * 2); CK(cudaMemcpy(ddm.p, mb.data(), ny * 2, cudaMemcpyHostToDevice));
    dsm.alloc(smask.size() * 4); CK(cudaMemcpy(dsm.p, smask.data(), smask.size() * 4, cudaMemcpyHostToDevice));
    dsw.alloc(sw.size() * 4); CK(cudaMemcpy(dsw.p, sw.data(), sw.size() * 4, cudaMemcpyHostToDevice));
    d.bias = (const float*)dbias.p; d.residual = dres.p; d.dmask = ddm.p; d.dmask_slope = 0.2f;
    d.stencil_mask = (const float*)dsm.p; d.stencil_w = (const float*)dsw.p;
    d.act = 2; d.act_slope = 0.2f;
  }
  std::vector<float> got(ny), got_act(ny);
  d.block_n = block_n;
  if (splits > 0) {
    dyf32.alloc(ny * 4 * (size_t)splits);  // one FP32 slice per split (summed in split order by the consumer)
    d.y_f32 = (float*)dyf32.p; d.splits = splits;
  } else {
    dyraw.alloc(ny * 2); d.y_raw = dyraw.p;
    if (epi) { dyact.alloc(ny * 2); d.y_act = dyact.p; }
  }
  int rc = spyr_conv2d_fprop(&d, 0);
  if (rc) { printf("  fprop rc=%d: %s\n", rc, spyr_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  fprop kernel error: %s\n", cudaGetErrorString(e)); exit(2); }
  if (splits > 0) {
    std::vector<float> slices(ny * (size_t)splits);
    CK(cudaMemcpy(slices.data(), dyf32.p, slices.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < ny; ++i) {
      float acc = 0.f;
      for (int sp = 0; sp < splits; ++sp) acc += slices[(size_t)sp * ny + i];
      got[i] = acc;
    }
  } else {
    std::vector<__nv_bfloat16> gb(ny);
    CK(cudaMemcpy(gb.data(), dyraw.p, ny * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < ny; ++i) got[i] = __bfloat162float(gb[i]);
    if (epi) {
      CK(cudaMemcpy(gb.data(), dyact.p, ny * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < ny; ++i) got_act[i] = __bfloat162float(gb[i]);
    }
  }
  const double err = rel_l2(got, yref);
  double err2 = 0;
  if (epi && splits == 0) err2 = rel_l2(got_act, yact);
  const bool ok = err < 5e-3 && err2 < 5e-3;
  printf("%s fprop B=%d %dx%d Cin=%d Cout=%d k=%d nsrc=%d epi=%d splits=%d bn=%d  relL2=%.3e act=%.3e\n",
         ok ? "PASS" : "FAIL", B, H, W, Cin, Cout, ks, nsrc, (int)epi, splits, block_n, err, err2);
  return ok ? 0 : 1;
}

static int test_wgrad(int B, int H, int W, int Cin, int Cout, int ks, int lbo, int sbo, int splits) {
  const size_t nx = (size_t)B * H * W * Cin, ny = (size_t)B * H * W * Cout;
  std::vector<float> x(nx), dy(ny);
  for (auto& v : x) v = bf16_round(frand());
  for (auto& v : dy) v = bf16_round(frand() * 0.1f);
  const int taps = ks * ks, pad = ks / 2;
  std::vector<float> ref((size_t)taps * Cin * Cout, 0.f);
  std::vector<double> acc(ref.size(), 0.0);
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        const float* dyp = &dy[(((size_t)b * H + h) * W + w) * Cout];
        for (int t = 0; t < taps; ++t) {
          const int hh = h + t / ks - pad, wx = w + t % ks - pad;
          if (hh < 0 || hh >= H || wx < 0 || wx >= W) continue;
          const float* xp = &x[(((size_t)b * H + hh) * W + wx) * Cin];
          for (int ci = 0; ci < Cin; ++ci) {
            double* a = &acc[((size_t)t * Cin + ci) * Cout];
            const double xv = xp[ci];
            for (int co = 0; co < Cout; ++co) a[co] += xv * dyp[co];
          }
        }
      }
  for (size_t i = 0; i < ref.size(); ++i) ref[i] = (float)acc[i];
  Dev dx, ddy, ddw;
  auto xb = to_bf16(x); auto yb = to_bf16(dy);
  dx.alloc(nx * 2); ddy.alloc(ny * 2); ddw.alloc(ref.size() * 4);
  CK(cudaMemcpy(dx.p, xb.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ddy.p, yb.data(), ny * 2, cudaMemcpyHostToDevice));
  spyr_wgrad_desc d;
  memset(&d, 0, sizeof(d));
  d.B = B; d.H = H; d.W = W; d.Cin = Cin; d.Cout = Cout; d.ksize = ks;
  d.x = dx.p; d.dy = ddy.p; d.dw = (float*)ddw.p; d.splits = splits; d.dbg_lbo = lbo; d.dbg_sbo = sbo;
  CK(cudaMemset(ddw.p, 0, ref.size() * 4));
  Dev dscratch;
  const long long need = spyr_conv2d_wgrad_scratch_floats(&d);
  if (need < 0) { printf("  wgrad scratch query: %s\n", spyr_last_error()); return 1; }
  if (need > 0) { dscratch.alloc((size_t)need * 4); d.scratch = (float*)dscratch.p; d.scratch_floats = need; }
  int rc = spyr_conv2d_wgrad(&d, 0);
  if (rc) { printf("  wgrad rc=%d: %s\n", rc, spyr_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  wgrad kernel error: %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<float> got(ref.size());
  CK(cudaMemcpy(got.data(), ddw.p, ref.size() * 4, cudaMemcpyDeviceToHost));
  const double err = rel_l2(got, ref);
  const bool ok = err < 2e-3;
  printf("%s wgrad B=%d %dx%d Cin=%d Cout=%d k=%d lbo=%d sbo=%d splits=%d relL2=%.3e\n", ok ? "PASS" : "FAIL", B, H, W,
         Cin, Cout, ks, lbo, sbo, splits, err);
  return ok ? 0 : 1;
}

// input-gradient mode: same fprop pack w[t][co][ci], read MN-major with flipped taps
static int test_dgrad(int B, int H, int W, int Cin_f, int Cout_f, int ks, bool f32) {
  const int taps = ks * ks, pad = ks / 2;
  std::vector<float> g((size_t)B * H * W * Cout_f), w((size_t)taps * Cout_f * Cin_f);
  for (auto& v : g) v = bf16_round(frand());
  for (auto& v : w) v = bf16_round(frand() * 0.05f);
  std::vector<float> ref((size_t)B * H * W * Cin_f, 0.f);
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int x = 0; x < W; ++x)
        for (int t = 0; t < taps; ++t) {
          // forward: y[h][x] += in[h+dy][x+dx] * w[t]  =>  gin[h+dy][x+dx] += g[h][x] * w[t]
          const int hh = h + t / ks - pad, xx = x + t % ks - pad;
          if (hh < 0 || hh >= H || xx < 0 || xx >= W) continue;
          const float* gp = &g[(((size_t)b * H + h) * W + x) * Cout_f];
          float* rp = &ref[(((size_t)b * H + hh) * W + xx) * Cin_f];
          for (int co = 0; co < Cout_f; ++co) {
            const float gv = gp[co];
            const float* wp = &w[((size_t)t * Cout_f + co) * Cin_f];
            for (int ci = 0; ci < Cin_f; ++ci) rp[ci] += gv * wp[ci];
          }
        }
  Dev dg, dw, dy;
  auto gb = to_bf16(g); auto wb = to_bf16(w);
  dg.alloc(gb.size() * 2); dw.alloc(wb.size() * 2);
  CK(cudaMemcpy(dg.p, gb.data(), gb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw.p, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
  spyr_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.B = B; d.H = H; d.W = W; d.Cout = Cin_f; d.nsrc = 1;
  d.src[0].x = dg.p; d.src[0].w = dw.p; d.src[0].cin = Cout_f; d.src[0].ksize = ks; d.src[0].w_mn_major = 1;
  const size_t ny = ref.size();
  dy.alloc(ny * (f32 ? 4 : 2));
  if (f32) { d.y_f32 = (float*)dy.p; d.f32_store = 1; } else d.y_raw = dy.p;
  int rc = spyr_conv2d_fprop(&d, 0);
  if (rc) { printf("  dgrad rc=%d: %s\n", rc, spyr_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  dgrad kernel error: %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<float> got(ny);
  if (f32) CK(cudaMemcpy(got.data(), dy.p, ny * 4, cudaMemcpyDeviceToHost));
  else {
    std::vector<__nv_bfloat16> ob(ny);
    CK(cudaMemcpy(ob.data(), dy.p, ny * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < ny; ++i) got[i] = __bfloat162float(ob[i]);
  }
  const double err = rel_l2(got, ref);
  const bool ok = err < (f32 ? 1e-5 : 5e-3);
  printf("%s dgrad B=%d %dx%d Cin_f=%d Cout_f=%d k=%d f32=%d relL2=%.3e\n", ok ? "PASS" : "FAIL", B, H, W, Cin_f, Cout_f,
         ks, (int)f32, err);
  return ok ? 0 : 1;
}

// per-image weights (batched GEMM): mn=0: y[b,p,n] = sum_k x[b,p,k] w[b][n][k];  mn=1: w[b][k][n]
static int test_per_image(int B, int H, int W, int K, int N, int mn) {
  std::vector<float> x((size_t)B * H * W * K), w((size_t)B * N * K);
  for (auto& v : x) v = bf16_round(frand());
  for (auto& v : w) v = bf16_round(frand() * 0.1f);
  std::vector<float> ref((size_t)B * H * W * N, 0.f);
  for (int b = 0; b < B; ++b)
    for (int p = 0; p < H * W; ++p)
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        for (int k = 0; k < K; ++k)
          acc += (double)x[((size_t)b * H * W + p) * K + k] * (mn ? w[((size_t)b * K + k) * N + n] : w[((size_t)b * N + n) * K + k]);
        ref[((size_t)b * H * W + p) * N + n] = (float)acc;
      }
  Dev dx, dw, dy;
  auto xb = to_bf16(x); auto wb = to_bf16(w);
  dx.alloc(xb.size() * 2); dw.alloc(wb.size() * 2); dy.alloc(ref.size() * 4);
  CK(cudaMemcpy(dx.p, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw.p, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
  spyr_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.B = B; d.H = H; d.W = W; d.Cout = N; d.nsrc = 1;
  d.src[0].x = dx.p; d.src[0].w = dw.p; d.src[0].cin = K; d.src[0].ksize = 1; d.src[0].w_mn_major = mn;
  d.src[0].w_per_image = 1;
  d.y_f32 = (float*)dy.p; d.f32_store = 1;
  int rc = spyr_conv2d_fprop(&d, 0);
  if (rc) { printf("  per_image rc=%d: %s\n", rc, spyr_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  per_image kernel error: %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<float> got(ref.size());
  CK(cudaMemcpy(got.data(), dy.p, ref.size() * 4, cudaMemcpyDeviceToHost));
  const double err = rel_l2(got, ref);
  const bool ok = err < 1e-5;
  printf("%s per_image B=%d %dx%d K=%d N=%d mn=%d relL2=%.3e\n", ok ? "PASS" : "FAIL", B, H, W, K, N, mn, err);
  return ok ? 0 : 1;
}

static int test_wgrad_per_image(int B, int H, int W, int Cin, int Cout) {
  const size_t nx = (size_t)B * H * W * Cin, ny = (size_t)B * H * W * Cout;
  std::vector<float> x(nx), dy(ny);
  for (auto& v : x) v = bf16_round(frand());
  for (auto& v : dy) v = bf16_round(frand() * 0.1f);
  std::vector<float> ref((size_t)B * Cin * Cout, 0.f);
  for (int b = 0; b < B; ++b)
    for (int p = 0; p < H * W; ++p)
      for (int ci = 0; ci < Cin; ++ci)
        for (int co = 0; co < Cout; ++co)
          ref[((size_t)b * Cin + ci) * Cout + co] += x[((size_t)b * H * W + p) * Cin + ci] * dy[((size_t)b * H * W + p) * Cout + co];
  Dev dx, ddy, ddw;
  auto xb = to_bf16(x); auto yb = to_bf16(dy);
  dx.alloc(nx * 2); ddy.alloc(ny * 2); ddw.alloc(ref.size() * 4);
  CK(cudaMemcpy(dx.p, xb.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ddy.p, yb.data(), ny * 2, cudaMemcpyHostToDevice));
  spyr_wgrad_desc d;
  memset(&d, 0, sizeof(d));
  d.B = B; d.H = H; d.W = W; d.Cin = Cin; d.Cout = Cout; d.ksize = 1;
  d.x = dx.p; d.dy = ddy.p; d.dw = (float*)ddw.p; d.per_image = 1;
  CK(cudaMemset(ddw.p, 0, ref.size() * 4));
  Dev dscratch;
  const long long need = spyr_conv2d_wgrad_scratch_floats(&d);
  if (need < 0) { printf("  wgrad scratch query: %s\n", spyr_last_error()); return 1; }
  if (need > 0) { dscratch.alloc((size_t)need * 4); d.scratch = (float*)dscratch.p; d.scratch_floats = need; }
  int rc = spyr_conv2d_wgrad(&d, 0);
  if (rc) { printf("  wgrad_pi rc=%d: %s\n", rc, spyr_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  wgrad_pi kernel error: %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<float> got(ref.size());
  CK(cudaMemcpy(got.data(), ddw.p, ref.size() * 4, cudaMemcpyDeviceToHost));
  const double err = rel_l2(got, ref);
  const bool ok = err < 1e-4;
  printf("%s wgrad_per_image B=%d %dx%d Cin=%d Cout=%d relL2=%.3e\n", ok ? "PASS" : "FAIL", B, H, W, Cin, Cout, err);
  return ok ? 0 : 1;
}

extern "C" int spyr_dbg_halo_probe(const void* x, const void* w, float* y, int B, int H, int W, int C, int Cout,
                                   int base_off_mode, void* stream);
static int test_halo(int B, int H, int W, int C, int Cout, int mode) {
  std::vector<float> x((size_t)B * H * W * C), w((size_t)9 * Cout * C);
  for (auto& v : x) v = bf16_round(frand());
  for (auto& v : w) v = bf16_round(frand() * 0.05f);
  std::vector<float> ref((size_t)B * H * W * Cout, 0.f);
  ref_conv(x, w, B, H, W, C, Cout, 3, ref);
  Dev dx, dw, dy;
  auto xb = to_bf16(x); auto wb = to_bf16(w);
  dx.alloc(xb.size() * 2); dw.alloc(wb.size() * 2); dy.alloc(ref.size() * 4);
  CK(cudaMemcpy(dx.p, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw.p, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
  int rc = spyr_dbg_halo_probe(dx.p, dw.p, (float*)dy.p, B, H, W, C, Cout, mode, 0);
  if (rc) { printf("  halo rc=%d: %s\n", rc, spyr_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  halo kernel error: %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<float> got(ref.size());
  CK(cudaMemcpy(got.data(), dy.p, ref.size() * 4, cudaMemcpyDeviceToHost));
  const double err = rel_l2(got, ref);
  const bool ok = err < 1e-5;
  printf("%s halo B=%d %dx%d C=%d Cout=%d base_off_mode=%d relL2=%.3e\n", ok ? "PASS" : "FAIL", B, H, W, C, Cout, mode, err);
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  int fails = 0;
  const char* mode = argc > 1 ? argv[1] : "all";
  if (!strcmp(mode, "all") || !strcmp(mode, "fprop")) {
    fails += test_fprop(2, 16, 16, 64, 64, 3, 1, false, 0, 0);
    fails += test_fprop(2, 16, 16, 64, 64, 1, 1, false, 0, 0);
    fails += test_fprop(1, 32, 32, 128, 128, 3, 1, false, 0, 0);
    fails += test_fprop(3, 8, 8, 128, 256, 3, 1, false, 0, 0);
    fails += test_fprop(5, 4, 4, 128, 512, 3, 1, true, 0, 0);
    fails += test_fprop(2, 32, 16, 256, 64, 3, 1, true, 0, 0);
    fails += test_fprop(2, 16, 16, 128, 128, 3, 3, true, 0, 0);
    fails += test_fprop(2, 16, 16, 32, 32, 3, 1, false, 0, 0);    // cin/cout < 64 -> TMA OOB fill
    fails += test_fprop(2, 16, 16, 64, 128, 3, 1, false, 0, 64);  // explicit N tile
    fails += test_fprop(20, 1, 1, 512, 365, 1, 1, false, 4, 128); // linear layer as 1x1, split-K, ragged Cout
    fails += test_fprop(2, 8, 8, 256, 128, 3, 1, false, 3, 0);
    // more tiles than SMs: resident weights reused across the tiles of a CTA, deep halo ring of the thin 1x1 layers
    fails += test_fprop(3, 128, 128, 64, 64, 3, 1, false, 0, 0);
    fails += test_fprop(3, 128, 128, 32, 64, 1, 1, true, 0, 0);
    fails += test_fprop(3, 128, 128, 64, 64, 3, 2, true, 0, 0);
    // conv_stack3.cu (Cout = 64, 3x3, W >= 128): ragged last column tile, 2-chunk and partial-chunk sources, 4-row map
    fails += test_fprop(2, 256, 256, 64, 64, 3, 1, true, 0, 0);
    fails += test_fprop(2, 128, 128, 128, 64, 3, 1, true, 0, 0);
    fails += test_fprop(1, 128, 128, 32, 64, 3, 1, false, 0, 0);
    fails += test_fprop(2, 4, 128, 64, 64, 3, 2, true, 0, 0);
    fails += test_fprop(1, 64, 256, 192, 64, 3, 1, false, 0, 0);
  }
  if (!strcmp(mode, "all") || !strcmp(mode, "wgrad")) {
    fails += test_wgrad(2, 16, 16, 64, 64, 3, 0, 0, 1);
    fails += test_wgrad(2, 16, 16, 64, 64, 1, 0, 0, 2);
    fails += test_wgrad(2, 16, 16, 128, 128, 3, 0, 0, 0);
    fails += test_wgrad(3, 8, 8, 128, 256, 3, 0, 0, 0);
    fails += test_wgrad(5, 4, 4, 64, 512, 3, 0, 0, 0);
    fails += test_wgrad(2, 32, 32, 256, 32, 1, 0, 0, 0);
    fails += test_wgrad(2, 64, 64, 64, 64, 3, 0, 0, 0);
    fails += test_wgrad(20, 1, 1, 512, 128, 1, 0, 0, 0);
    // halo-tiled kernel: 64-channel inputs pair taps, 128-channel blocks pair chunks; ragged Cout; several splits
    fails += test_wgrad(2, 16, 16, 64, 64, 3, 0, 0, 3);
    fails += test_wgrad(1, 32, 32, 128, 128, 3, 0, 0, 0);
    fails += test_wgrad(2, 32, 16, 256, 64, 3, 0, 0, 0);
    fails += test_wgrad(2, 16, 16, 128, 256, 3, 0, 0, 1);
    fails += test_wgrad(2, 16, 16, 64, 24, 3, 0, 0, 0);
    fails += test_wgrad(1, 64, 64, 384, 128, 3, 0, 0, 0);
  }
  if (!strcmp(mode, "all") || !strcmp(mode, "dgrad")) {
    fails += test_dgrad(2, 16, 16, 64, 64, 3, false);
    fails += test_dgrad(2, 16, 16, 128, 256, 3, true);
    fails += test_dgrad(3, 8, 8, 256, 128, 3, false);
    fails += test_dgrad(2, 16, 16, 64, 128, 1, true);
    fails += test_dgrad(2, 32, 32, 32, 64, 1, true);   // im2col'd first layer: 32 output channels
    fails += test_dgrad(5, 4, 4, 512, 768, 3, true);
    fails += test_dgrad(2, 16, 16, 8, 64, 1, true);    // padded 3-channel skip path
    fails += test_dgrad(3, 128, 128, 64, 64, 3, true); // resident MN-major weights, several tiles per CTA
    fails += test_per_image(3, 32, 32, 32, 256, 0);    // S = Q K^T
    fails += test_per_image(3, 32, 32, 256, 128, 1);   // O = P V
    fails += test_per_image(3, 32, 32, 128, 256, 0);   // dP = dO V^T
    fails += test_per_image(3, 32, 32, 256, 32, 1);    // dQ = dS K
    fails += test_per_image(2, 16, 16, 64, 16, 1);     // channel_factor 2 head dims
    fails += test_wgrad_per_image(3, 32, 32, 256, 32);  // dK
    fails += test_wgrad_per_image(3, 32, 32, 256, 128); // dV
    fails += test_wgrad_per_image(2, 16, 16, 64, 16);
  }
  if (!strcmp(mode, "halo")) {
    for (int m = 0; m < 2; ++m) {
      fails += test_halo(2, 16, 16, 64, 64, m);
      fails += test_halo(1, 32, 16, 128, 32, m);
    }
  }
  printf("native conv tests: %d failure(s)\n", fails);
  return fails ? 1 : 0;
}
