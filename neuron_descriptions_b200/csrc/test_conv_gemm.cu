// Standalone check of the tcgen05 conv/GEMM kernel against a double-precision CPU convolution.
// Build: make -C neuron_descriptions_b200/csrc test_conv_gemm ; run on a B200 (gpurun).
#include "conv_gemm.h"
#include "encoder.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

using namespace milan;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

static uint16_t f2bf(float f) {  // round-to-nearest-even
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return static_cast<uint16_t>(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static void split_vec(const std::vector<float>& x, std::vector<uint16_t>& hi, std::vector<uint16_t>& lo) {
  hi.resize(x.size());
  lo.resize(x.size());
  for (size_t i = 0; i < x.size(); ++i) {
    hi[i] = f2bf(x[i]);
    lo[i] = f2bf(x[i] - bf2f(hi[i]));
  }
}
template <class T>
static T* upload(const std::vector<T>& v) {
  T* d;
  CK(cudaMalloc(&d, v.size() * sizeof(T) + 256));
  CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

struct Case {
  const char* name;
  int N, H, W, Cin, Cout, ks, stride, relu, residual, split, f32out;
};

static int run_case(const Case& c, int num_sms, int reps = 5) {
  std::mt19937 rng(1234 + c.Cin * 7 + c.Cout + c.H);
  std::normal_distribution<float> nd(0.f, 1.f);
  const int Ho = c.H / c.stride, Wo = c.W / c.stride;
  const size_t in_elems = static_cast<size_t>(c.N) * c.H * c.W * c.Cin;
  const int taps = c.ks * c.ks;
  const size_t w_elems = static_cast<size_t>(c.Cout) * taps * c.Cin;
  const size_t out_elems = static_cast<size_t>(c.N) * Ho * Wo * c.Cout;
  std::vector<float> x(in_elems), w(w_elems), bias((c.Cout + 127) / 128 * 128, 0.f), res(out_elems);
  for (auto& v : x) v = nd(rng);
  const float wscale = 1.0f / std::sqrt(static_cast<float>(taps * c.Cin));
  for (auto& v : w) v = nd(rng) * wscale;
  for (int i = 0; i < c.Cout; ++i) bias[i] = nd(rng) * 0.1f;
  for (auto& v : res) v = nd(rng);
  std::vector<uint16_t> xh, xl, wh, wl, rh, rl;
  split_vec(x, xh, xl);
  split_vec(w, wh, wl);
  split_vec(res, rh, rl);
  // The CPU reference uses exactly the values the GPU sees.
  std::vector<float> xe(in_elems), we(w_elems), re(out_elems);
  for (size_t i = 0; i < in_elems; ++i) xe[i] = c.split ? bf2f(xh[i]) + bf2f(xl[i]) : bf2f(xh[i]);
  for (size_t i = 0; i < w_elems; ++i) we[i] = c.split ? bf2f(wh[i]) + bf2f(wl[i]) : bf2f(wh[i]);
  for (size_t i = 0; i < out_elems; ++i) re[i] = c.split ? bf2f(rh[i]) + bf2f(rl[i]) : bf2f(rh[i]);

  uint16_t *dxh = upload(xh), *dxl = upload(xl), *dwh = upload(wh), *dwl = upload(wl), *drh = upload(rh),
           *drl = upload(rl);
  float* dbias = upload(bias);
  uint16_t *doh, *dol;
  float* dof;
  CK(cudaMalloc(&doh, out_elems * 2 + 256));
  CK(cudaMalloc(&dol, out_elems * 2 + 256));
  CK(cudaMalloc(&dof, out_elems * 4 + 256));
  CK(cudaMemset(doh, 0xFF, out_elems * 2));
  CK(cudaMemset(dol, 0xFF, out_elems * 2));
  CK(cudaMemset(dof, 0xFF, out_elems * 4));

  ConvDesc d{c.N, c.H, c.W, c.Cin, c.Cout, c.ks, c.stride};
  ConvIO io{};
  io.in_hi = reinterpret_cast<__nv_bfloat16*>(dxh);
  io.in_lo = reinterpret_cast<__nv_bfloat16*>(dxl);
  io.w_hi = reinterpret_cast<__nv_bfloat16*>(dwh);
  io.w_lo = reinterpret_cast<__nv_bfloat16*>(dwl);
  io.bias = dbias;
  if (c.residual) {
    io.res_hi = reinterpret_cast<__nv_bfloat16*>(drh);
    io.res_lo = reinterpret_cast<__nv_bfloat16*>(drl);
  }
  if (!c.f32out) {
    io.out_hi = reinterpret_cast<__nv_bfloat16*>(doh);
    io.out_lo = reinterpret_cast<__nv_bfloat16*>(dol);
  }
  io.out_f32 = dof;
  io.relu = c.relu;
  ConvGemmParams p;
  int block_n = 0;
  int rc = build_conv_params(&p, d, io, c.split, &block_n);
  if (rc) {
    printf("[%s] build_conv_params failed rc=%d: %s\n", c.name, rc, tmap_last_error());
    return 1;
  }
  if (c.f32out) block_n = 128;
  if (c.f32out && c.Cout <= 64) { printf("[%s] bad case\n", c.name); return 1; }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  rc = launch_conv_gemm(p, block_n, c.split, c.f32out ? EPI_F32 : EPI_BF16, num_sms, 0);
  if (rc) { printf("[%s] launch failed rc=%d\n", c.name, rc); return 1; }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("[%s] kernel failed: %s\n", c.name, cudaGetErrorString(se)); return 1; }
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch_conv_gemm(p, block_n, c.split, c.f32out ? EPI_F32 : EPI_BF16, num_sms, 0);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;

  std::vector<uint16_t> oh(out_elems), ol(out_elems);
  std::vector<float> of(out_elems);
  CK(cudaMemcpy(oh.data(), doh, out_elems * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ol.data(), dol, out_elems * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(of.data(), dof, out_elems * 4, cudaMemcpyDeviceToHost));

  // CPU reference (sampled if large).
  const int pad = c.ks / 2;
  double max_err = 0, max_ref = 0;
  size_t checked = 0, bad = 0;
  const size_t stride_check = out_elems > 4000000 ? 97 : 1;
  for (size_t idx = 0; idx < out_elems; idx += stride_check) {
    const int co = idx % c.Cout;
    size_t pix = idx / c.Cout;
    const int ow = pix % Wo;
    const int oh_ = (pix / Wo) % Ho;
    const int n = pix / (static_cast<size_t>(Wo) * Ho);
    double acc = bias[co];
    for (int r = 0; r < c.ks; ++r) {
      const int ih = oh_ * c.stride + r - pad;
      if (ih < 0 || ih >= c.H) continue;
      for (int s = 0; s < c.ks; ++s) {
        const int iw = ow * c.stride + s - pad;
        if (iw < 0 || iw >= c.W) continue;
        const float* xp = &xe[((static_cast<size_t>(n) * c.H + ih) * c.W + iw) * c.Cin];
        const float* wp = &we[(static_cast<size_t>(co) * taps + r * c.ks + s) * c.Cin];
        for (int ci = 0; ci < c.Cin; ++ci) acc += static_cast<double>(xp[ci]) * wp[ci];
      }
    }
    if (c.residual && !c.f32out) acc += re[idx];
    if (c.relu) acc = acc > 0 ? acc : 0;
    double got;
    if (c.f32out) got = of[idx];
    else got = c.split ? static_cast<double>(bf2f(oh[idx])) + bf2f(ol[idx]) : bf2f(oh[idx]);
    const double err = std::fabs(got - acc);
    if (!(err <= 1e30)) { ++bad; }
    if (err > max_err) max_err = err;
    if (std::fabs(acc) > max_ref) max_ref = std::fabs(acc);
    ++checked;
  }
  const double tol = c.split ? 2e-4 : 6e-2;
  const double flops = 2.0 * out_elems * taps * c.Cin;
  const bool ok = bad == 0 && max_err <= tol * (max_ref > 1 ? max_ref : 1);
  printf("[%s] N=%d %dx%d Cin=%d Cout=%d k=%d s=%d split=%d f32=%d box=(%d,%d,%d) tiles=%d bn=%d : max_err=%.3e "
         "(max_ref=%.2f, nan=%zu, checked=%zu) %.3f ms %.1f TFLOP/s %s\n",
         c.name, c.N, c.H, c.W, c.Cin, c.Cout, c.ks, c.stride, c.split, c.f32out, p.box_w, p.box_h, p.box_n,
         p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles, block_n, max_err, max_ref, bad, checked, ms,
         flops / ms * 1e-9, ok ? "OK" : "FAIL");
  cudaFree(dxh); cudaFree(dxl); cudaFree(dwh); cudaFree(dwl); cudaFree(drh); cudaFree(drl);
  cudaFree(dbias); cudaFree(doh); cudaFree(dol); cudaFree(dof);
  return ok ? 0 : 1;
}

// 7x7 stride-2 stem (raw input rows + overlapping-window MMA descriptor) vs a direct CPU convolution.
static int run_stem(int N, int split, int num_sms) {
  std::mt19937 rng(99);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::uniform_int_distribution<int> ud(0, 255);
  const size_t img_elems = static_cast<size_t>(N) * 3 * 224 * 224;
  std::vector<uint8_t> img(img_elems);
  for (auto& v : img) v = static_cast<uint8_t>(ud(rng));
  std::vector<float> w(64 * 3 * 49), packed(64 * kStemKTotal);
  for (auto& v : w) v = nd(rng) * 0.08f;
  pack_stem_weights(w.data(), packed.data());
  std::vector<uint16_t> wh, wl;
  split_vec(packed, wh, wl);
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  uint8_t* dimg = upload(img);
  uint16_t *dwh = upload(wh), *dwl = upload(wl), *dph, *dpl, *doh, *dol;
  const size_t pad_elems = static_cast<size_t>(N) * kStemPadH * kStemPadW * 4;
  const size_t out_elems = static_cast<size_t>(N) * 112 * 112 * 64;
  CK(cudaMalloc(&dph, pad_elems * 2 + 256));
  CK(cudaMalloc(&dpl, pad_elems * 2 + 256));
  CK(cudaMalloc(&doh, out_elems * 2 + 256));
  CK(cudaMalloc(&dol, out_elems * 2 + 256));
  CK(cudaMemset(doh, 0xFF, out_elems * 2));
  CK(cudaMemset(dol, 0xFF, out_elems * 2));
  int rc = launch_stem_pack(dimg, 0, N, reinterpret_cast<__nv_bfloat16*>(dph), reinterpret_cast<__nv_bfloat16*>(dpl), mean,
                            stdv, split, 0);
  if (rc) { printf("[stem] pack launch failed %d\n", rc); return 1; }
  ConvGemmParams p;
  rc = build_stem_params(&p, N, reinterpret_cast<__nv_bfloat16*>(dph), reinterpret_cast<__nv_bfloat16*>(dpl),
                         reinterpret_cast<__nv_bfloat16*>(dwh), reinterpret_cast<__nv_bfloat16*>(dwl),
                         reinterpret_cast<__nv_bfloat16*>(doh), reinterpret_cast<__nv_bfloat16*>(dol), split);
  if (rc) { printf("[stem] build_stem_params failed rc=%d: %s\n", rc, tmap_last_error()); return 1; }
  rc = launch_conv_gemm(p, 64, split, EPI_BF16, num_sms, 0);
  if (rc) { printf("[stem] launch failed rc=%d\n", rc); return 1; }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("[stem] kernel failed: %s\n", cudaGetErrorString(se)); return 1; }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 3; ++i) launch_conv_gemm(p, 64, split, EPI_BF16, num_sms, 0);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= 3;
  std::vector<uint16_t> oh(out_elems), ol(out_elems);
  CK(cudaMemcpy(oh.data(), doh, out_elems * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ol.data(), dol, out_elems * 2, cudaMemcpyDeviceToHost));
  // CPU reference on the values the GPU sees
  std::vector<float> xn(img_elems);
  for (size_t i = 0; i < img_elems; ++i) {
    const int c = (i / (224 * 224)) % 3;
    const float x = static_cast<float>(img[i]) * 0.00392156862745098f;
    const float v = (x - mean[c]) / stdv[c];
    const uint16_t h = f2bf(v);
    xn[i] = split ? bf2f(h) + bf2f(f2bf(v - bf2f(h))) : bf2f(h);
  }
  std::vector<float> we(w.size());
  for (size_t i = 0; i < w.size(); ++i) {
    const uint16_t h = f2bf(w[i]);
    we[i] = split ? bf2f(h) + bf2f(f2bf(w[i] - bf2f(h))) : bf2f(h);
  }
  double max_err = 0, max_ref = 0;
  size_t checked = 0;
  for (size_t idx = 0; idx < out_elems; idx += 53) {
    const int co = idx % 64;
    size_t pix = idx / 64;
    const int ow = pix % 112, oh_ = (pix / 112) % 112, n = pix / (112 * 112);
    double acc = 0;
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 7; ++r) {
        const int ih = oh_ * 2 + r - 3;
        if (ih < 0 || ih >= 224) continue;
        for (int s = 0; s < 7; ++s) {
          const int iw = ow * 2 + s - 3;
          if (iw < 0 || iw >= 224) continue;
          acc += static_cast<double>(xn[((static_cast<size_t>(n) * 3 + c) * 224 + ih) * 224 + iw]) *
                 we[((co * 3 + c) * 7 + r) * 7 + s];
        }
      }
    const double got = split ? static_cast<double>(bf2f(oh[idx])) + bf2f(ol[idx]) : bf2f(oh[idx]);
    const double err = std::fabs(got - acc);
    if (!(err <= 1e30)) max_err = 1e30;
    if (err > max_err) max_err = err;
    if (std::fabs(acc) > max_ref) max_ref = std::fabs(acc);
    ++checked;
  }
  const bool ok = max_err <= (split ? 2e-4 : 6e-2) * (max_ref > 1 ? max_ref : 1);
  printf("[stem N=%d split=%d] box=(%d,%d,%d) tiles=%d : max_err=%.3e (max_ref=%.2f, checked=%zu) %.3f ms %.1f TFLOP/s(alg) %s\n",
         N, split, p.box_w, p.box_h, p.box_n, p.tiles_w * p.tiles_h * p.tiles_n, max_err, max_ref, checked, ms,
         2.0 * out_elems * 147 / ms * 1e-9, ok ? "OK" : "FAIL");
  cudaFree(dimg); cudaFree(dwh); cudaFree(dwl); cudaFree(dph); cudaFree(dpl); cudaFree(doh); cudaFree(dol);
  return ok ? 0 : 1;
}

// out = relu(W_a a + W_b b(strided) + bias): the fused conv3 + downsample tail of a stage's first bottleneck.
static int run_dual(const char* name, int N, int Ho, int Wo, int Ca, int Cb, int Cout, int stride, int num_sms) {
  std::mt19937 rng(77 + Ca + Cb + stride);
  std::normal_distribution<float> nd(0.f, 1.f);
  const int Hi = Ho * stride, Wi = Wo * stride;
  const size_t a_elems = static_cast<size_t>(N) * Ho * Wo * Ca, b_elems = static_cast<size_t>(N) * Hi * Wi * Cb;
  const size_t w_elems = static_cast<size_t>(Cout) * (Ca + Cb), out_elems = static_cast<size_t>(N) * Ho * Wo * Cout;
  std::vector<float> a(a_elems), b(b_elems), w(w_elems), bias((Cout + 127) / 128 * 128, 0.f);
  for (auto& v : a) v = nd(rng);
  for (auto& v : b) v = nd(rng);
  for (auto& v : w) v = nd(rng) / std::sqrt(static_cast<float>(Ca + Cb));
  for (int i = 0; i < Cout; ++i) bias[i] = nd(rng) * 0.1f;
  std::vector<uint16_t> ah, al, bh, bl, wh, wl;
  split_vec(a, ah, al);
  split_vec(b, bh, bl);
  split_vec(w, wh, wl);
  auto eff = [](const std::vector<uint16_t>& h, const std::vector<uint16_t>& l, size_t i) { return bf2f(h[i]) + bf2f(l[i]); };
  uint16_t *dah = upload(ah), *dal = upload(al), *dbh = upload(bh), *dbl = upload(bl), *dwh = upload(wh), *dwl = upload(wl);
  float* dbias = upload(bias);
  uint16_t *doh, *dol;
  CK(cudaMalloc(&doh, out_elems * 2 + 256));
  CK(cudaMalloc(&dol, out_elems * 2 + 256));
  CK(cudaMemset(doh, 0xFF, out_elems * 2));
  CK(cudaMemset(dol, 0xFF, out_elems * 2));
  DualConvIO io{};
  io.a_hi = reinterpret_cast<__nv_bfloat16*>(dah); io.a_lo = reinterpret_cast<__nv_bfloat16*>(dal);
  io.b_hi = reinterpret_cast<__nv_bfloat16*>(dbh); io.b_lo = reinterpret_cast<__nv_bfloat16*>(dbl);
  io.w_hi = reinterpret_cast<__nv_bfloat16*>(dwh); io.w_lo = reinterpret_cast<__nv_bfloat16*>(dwl);
  io.bias = dbias;
  io.out_hi = reinterpret_cast<__nv_bfloat16*>(doh); io.out_lo = reinterpret_cast<__nv_bfloat16*>(dol);
  io.relu = 1;
  ConvGemmParams p;
  int block_n = 0;
  int rc = build_dual_1x1_params(&p, N, Ho, Wo, Ca, Cb, Cout, stride, io, 1, &block_n);
  if (rc) { printf("[%s] build_dual_1x1_params failed rc=%d: %s\n", name, rc, tmap_last_error()); return 1; }
  rc = launch_conv_gemm(p, block_n, 1, EPI_BF16, num_sms, 0);
  if (rc) { printf("[%s] launch failed rc=%d\n", name, rc); return 1; }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("[%s] kernel failed: %s\n", name, cudaGetErrorString(se)); return 1; }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 5; ++i) launch_conv_gemm(p, block_n, 1, EPI_BF16, num_sms, 0);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= 5;
  std::vector<uint16_t> oh(out_elems), ol(out_elems);
  CK(cudaMemcpy(oh.data(), doh, out_elems * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ol.data(), dol, out_elems * 2, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0;
  size_t checked = 0;
  const size_t step = out_elems > 4000000 ? 97 : 1;
  for (size_t idx = 0; idx < out_elems; idx += step) {
    const int co = idx % Cout;
    const size_t pix = idx / Cout;
    const int ow = pix % Wo, oh_ = (pix / Wo) % Ho, n = pix / (static_cast<size_t>(Wo) * Ho);
    double acc = bias[co];
    for (int c = 0; c < Ca; ++c) acc += static_cast<double>(eff(ah, al, pix * Ca + c)) * eff(wh, wl, static_cast<size_t>(co) * (Ca + Cb) + c);
    const size_t bpix = (static_cast<size_t>(n) * Hi + oh_ * stride) * Wi + ow * stride;
    for (int c = 0; c < Cb; ++c) acc += static_cast<double>(eff(bh, bl, bpix * Cb + c)) * eff(wh, wl, static_cast<size_t>(co) * (Ca + Cb) + Ca + c);
    acc = acc > 0 ? acc : 0;
    const double got = static_cast<double>(bf2f(oh[idx])) + bf2f(ol[idx]);
    const double err = std::fabs(got - acc);
    if (!(err <= 1e30)) max_err = 1e30;
    if (err > max_err) max_err = err;
    if (std::fabs(acc) > max_ref) max_ref = std::fabs(acc);
    ++checked;
  }
  const bool ok = max_err <= 2e-4 * (max_ref > 1 ? max_ref : 1);
  printf("[%s] N=%d %dx%d Ca=%d Cb=%d Cout=%d s=%d box=(%d,%d,%d) : max_err=%.3e (max_ref=%.2f, checked=%zu) %.3f ms "
         "%.1f TFLOP/s %s\n", name, N, Ho, Wo, Ca, Cb, Cout, stride, p.box_w, p.box_h, p.box_n, max_err, max_ref, checked,
         ms, 2.0 * out_elems * (Ca + Cb) / ms * 1e-9, ok ? "OK" : "FAIL");
  cudaFree(dah); cudaFree(dal); cudaFree(dbh); cudaFree(dbl); cudaFree(dwh); cudaFree(dwl); cudaFree(dbias);
  cudaFree(doh); cudaFree(dol);
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
  const int sms = prop.multiProcessorCount;

  std::vector<Case> cases = {
      // name            N   H   W  Cin  Cout ks st relu res split f32
      {"gemm_small",     1,  1, 128,  64,  128, 1, 1, 0, 0, 0, 0},
      {"gemm_small_sp",  1,  1, 128,  64,  128, 1, 1, 0, 0, 1, 0},
      {"gemm_k256",      1,  1, 300, 256,  128, 1, 1, 0, 0, 1, 0},
      {"gemm_f32_tail",  1,  1, 200, 512,  300, 1, 1, 0, 0, 1, 1},
      {"acc_exact_k64",  1,  1, 256,   64,  256, 1, 1, 0, 0, 0, 1},
      {"acc_exact_k512", 1,  1, 256,  512,  256, 1, 1, 0, 0, 0, 1},
      {"acc_exact_k4544",1,  1, 256, 4544,  256, 1, 1, 0, 0, 0, 1},
      {"acc_split_k4544",1,  1, 256, 4544,  256, 1, 1, 0, 0, 1, 1},
      {"c1x1_56",        2, 56,  56,  64,  256, 1, 1, 1, 1, 1, 0},
      {"c1x1_ragged",    1,  1, 300, 256,  256, 1, 1, 1, 1, 1, 0},
      {"c1x1_ragged_f",  1,  1, 333, 128,  512, 1, 1, 1, 0, 0, 0},
      {"c1x1_cout64",    2, 56,  56, 256,   64, 1, 1, 1, 0, 1, 0},
      {"c3x3_56",        3, 56,  56,  64,   64, 3, 1, 1, 0, 1, 0},
      {"c3x3_28",        5, 28,  28, 128,  128, 3, 1, 1, 0, 1, 0},
      {"c3x3_14",        7, 14,  14, 256,  256, 3, 1, 1, 0, 1, 0},
      {"c3x3_7",         9,  7,   7, 512,  512, 3, 1, 1, 0, 1, 0},
      {"c3x3_s2_56",     3, 56,  56, 128,  128, 3, 2, 1, 0, 1, 0},
      {"c3x3_s2_14",     5, 14,  14, 512,  512, 3, 2, 1, 0, 1, 0},
      {"c1x1_s2_56",     3, 56,  56, 256,  512, 1, 2, 0, 0, 1, 0},
      {"c3x3_28_fast",   5, 28,  28, 128,  128, 3, 1, 1, 1, 0, 0},
      // throughput-sized (sampled check)
      {"perf_l3_1x1",  240, 14,  14, 1024, 256, 1, 1, 1, 0, 1, 0},
      {"perf_l3_3x3",  240, 14,  14, 256,  256, 3, 1, 1, 0, 1, 0},
      {"perf_l3_1x1b", 240, 14,  14, 256, 1024, 1, 1, 1, 1, 1, 0},
      {"perf_l3_3x3f", 240, 14,  14, 256,  256, 3, 1, 1, 0, 0, 0},
      {"perf_l1_3x3",  240, 56,  56,  64,   64, 3, 1, 1, 0, 1, 0},
      {"perf_l1_c3",   240, 56,  56,  64,  256, 1, 1, 1, 1, 1, 0},
      {"perf_l2_c3",   240, 28,  28, 128,  512, 1, 1, 1, 1, 1, 0},
      {"perf_l1_c1",   240, 56,  56, 256,   64, 1, 1, 1, 0, 1, 0},
      {"perf_l3_c3f",  240, 14,  14, 256, 1024, 1, 1, 1, 1, 0, 0},
  };
  int fails = 0;
  if (argc > 1 && std::string(argv[1]) == "l2") {
    // L2-residency sweep: the same problems at image counts whose tensors fit the 126 MB L2 (repeated launches on
    // one buffer set), next to the HBM-streaming sizes. Per-image time tells what a depth-first schedule could gain.
    const int ns[] = {12, 24, 48, 96, 240};
    const Case shapes[] = {
        {"l3_reduce", 0, 14, 14, 1024, 256, 1, 1, 1, 0, 1, 0}, {"l3_3x3", 0, 14, 14, 256, 256, 3, 1, 1, 0, 1, 0},
        {"l3_expand", 0, 14, 14, 256, 1024, 1, 1, 1, 1, 1, 0}, {"l2_reduce", 0, 28, 28, 512, 128, 1, 1, 1, 0, 1, 0},
        {"l2_3x3", 0, 28, 28, 128, 128, 3, 1, 1, 0, 1, 0},     {"l2_expand", 0, 28, 28, 128, 512, 1, 1, 1, 1, 1, 0},
        {"l1_reduce", 0, 56, 56, 256, 64, 1, 1, 1, 0, 1, 0},   {"l1_3x3", 0, 56, 56, 64, 64, 3, 1, 1, 0, 1, 0},
        {"l1_expand", 0, 56, 56, 64, 256, 1, 1, 1, 1, 1, 0},   {"l4_reduce", 0, 7, 7, 2048, 512, 1, 1, 1, 0, 1, 0},
        {"l4_3x3", 0, 7, 7, 512, 512, 3, 1, 1, 0, 1, 0},       {"l4_expand", 0, 7, 7, 512, 2048, 1, 1, 1, 1, 1, 0}};
    for (const Case& sh : shapes)
      for (int n : ns) {
        Case c = sh;
        c.N = n;
        fails += run_case(c, sms, 20);
        fflush(stdout);
      }
    printf("%s (%d failures)\n", fails ? "SOME FAILED" : "ALL OK", fails);
    return fails ? 1 : 0;
  }
  if (argc > 1 && std::string(argv[1]) == "stem") {  // the 240-image stem alone (ncu captures)
    fails += run_stem(240, 1, sms);
    return fails ? 1 : 0;
  }
  if (argc > 1 && std::string(argv[1]) == "res") {  // the residual (expand) convs alone, for A/B runs of their kernel
    const Case shapes[] = {
        {"l1_expand", 240, 56, 56, 64, 256, 1, 1, 1, 1, 1, 0},  {"l2_expand", 240, 28, 28, 128, 512, 1, 1, 1, 1, 1, 0},
        {"l3_expand", 240, 14, 14, 256, 1024, 1, 1, 1, 1, 1, 0}, {"l3_expand", 960, 14, 14, 256, 1024, 1, 1, 1, 1, 1, 0},
        {"l4_expand", 240, 7, 7, 512, 2048, 1, 1, 1, 1, 1, 0},   {"l4_expand", 960, 7, 7, 512, 2048, 1, 1, 1, 1, 1, 0},
        {"ragged_res", 1, 1, 333, 256, 384, 1, 1, 1, 1, 1, 0},   {"odd_tiles", 1, 1, 128 * 5 + 7, 64, 128, 1, 1, 1, 1, 1, 0}};
    const int pick = argc > 2 ? atoi(argv[2]) : -1;  // `res 3`: one case only (ncu captures)
    int index = 0;
    for (const Case& c : shapes) {
      if (pick >= 0 && index++ != pick) continue;
      fails += run_case(c, sms, pick >= 0 ? 2 : 10);
      fflush(stdout);
    }
    printf("%s (%d failures)\n", fails ? "SOME FAILED" : "ALL OK", fails);
    return fails ? 1 : 0;
  }
  int only = argc > 1 ? atoi(argv[1]) : -1;
  for (size_t i = 0; i < cases.size(); ++i) {
    if (only >= 0 && static_cast<int>(i) != only) continue;
    fails += run_case(cases[i], sms);
    fflush(stdout);
  }
  if (only < 0 || only >= 100) {
    fails += run_dual("dual_l1", 2, 56, 56, 64, 64, 256, 1, sms);
    fails += run_dual("dual_l2", 3, 28, 28, 128, 256, 512, 2, sms);
    fails += run_dual("dual_l3", 5, 14, 14, 256, 512, 1024, 2, sms);
    fails += run_dual("dual_l4", 7, 7, 7, 512, 1024, 2048, 2, sms);
    fails += run_dual("perf_dual_l4", 240, 7, 7, 512, 1024, 2048, 2, sms);
    fails += run_dual("perf_dual_l1", 240, 56, 56, 64, 64, 256, 1, sms);
    fails += run_dual("perf_dual_l2", 240, 28, 28, 128, 256, 512, 2, sms);
    fails += run_dual("perf_dual_l3", 240, 14, 14, 256, 512, 1024, 2, sms);
    fails += run_stem(3, 1, sms);
    fails += run_stem(2, 0, sms);
    fails += run_stem(240, 1, sms);
  }
  printf("%s (%d failures)\n", fails ? "SOME FAILED" : "ALL OK", fails);
  return fails ? 1 : 0;
}
