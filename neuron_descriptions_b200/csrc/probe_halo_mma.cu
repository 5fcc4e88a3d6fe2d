// Probe (not part of the library; `make probes`): can the nine taps of a 3x3 convolution read ONE halo tile in shared
// memory? The tile is what a TMA box (64 channels, 10 pixels, 18 rows) with the 128-byte swizzle would write: 180 rows
// of 128 B, 16-byte chunk j of row r stored at r * 128 + ((j ^ (r & 7)) << 4). Output pixel (y, x) of an 8 x 16 tile
// (MMA row m = 8 y + x) needs, for tap (dh, dw), halo row (y + dh) * 10 + (x + dw): a K-major descriptor starting
// (dh * 10 + dw) * 128 bytes into the tile with SBO = 1280 (one 8-row MMA group per image row). Neither the start
// nor SBO is a multiple of the 1024-byte swizzle period, so the question is whether the hardware derives the XOR
// phase from absolute address bits (then this just works) or relative to the descriptor start (then the descriptor's
// base-offset field, bits 49-51, has to carry it and a per-group phase shift could not be expressed at all).
// D = A_tap x B^T, 128 x 128 x 64, checked against the CPU for every tap and both conventions.
//   (built by `make probes`)
#include "conv_gemm.h"
#include "ptx.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

using namespace milan;

constexpr int kHaloW = 10, kHaloH = 18, kHaloRows = kHaloW * kHaloH;  // 180 rows of 128 B

__global__ void __launch_bounds__(128, 1) halo_kernel(const __grid_constant__ CUtensorMap tmap_b, const uint16_t* halo,
                                                      float* out, int dh, int dw, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;           // 180 x 128 B = 23040 B (room: 24 KB)
  uint8_t* b_s = smem + 24576;   // 128 x 128 B, TMA, SW128
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 24576 + 16384);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // swizzled copy of the halo tile, 16 bytes at a time (what TMA would have written)
  for (int i = threadIdx.x; i < kHaloRows * 8; i += blockDim.x) {
    const int r = i >> 3, j = i & 7;
    const uint4 v = reinterpret_cast<const uint4*>(halo)[i];
    *reinterpret_cast<uint4*>(a_s + r * 128 + ((j ^ (r & 7)) << 4)) = v;
  }
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 128);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 0) {  // warp-uniform issue
    mbar_arrive_expect_tx_elect(&bars[0], 16384);
    tma_load_2d_elect(b_s, &tmap_b, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tcgen05_fence_after();
    const uint32_t idesc = make_idesc_16bit(128, 128, 1u);
    const uint32_t start = smem_u32(a_s) + static_cast<uint32_t>(dh * kHaloW + dw) * 128u;
    uint64_t da = static_cast<uint64_t>((start >> 4) & 0x3FFF);
    da |= static_cast<uint64_t>((kHaloW * 128) >> 4) << 32;  // SBO = 1280
    da |= static_cast<uint64_t>(1) << 46;
    if (use_base_offset) da |= static_cast<uint64_t>((start >> 7) & 7) << 49;
    da |= static_cast<uint64_t>(2) << 61;                    // SWIZZLE_128B
    const uint64_t db = make_smem_desc_sw128(smem_u32(b_s));
    for (int k = 0; k < 4; ++k) umma_bf16_elect(tmem_base, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
    umma_commit_elect(&bars[1]);
  }
  __syncthreads();
  mbar_wait(&bars[1], 0);
  tcgen05_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 4; ++c) {
    uint32_t acc[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, acc);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[row * 128 + c * 32 + j] = __uint_as_float(acc[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

int main() {
  std::mt19937 rng(3);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<uint16_t> halo(kHaloRows * 64), b(128 * 64);
  for (auto& v : halo) v = f2bf(nd(rng));
  for (auto& v : b) v = f2bf(nd(rng));
  uint16_t *dhalo, *db;
  float* dout;
  cudaMalloc(&dhalo, halo.size() * 2);
  cudaMalloc(&db, b.size() * 2);
  cudaMalloc(&dout, 128 * 128 * 4);
  cudaMemcpy(dhalo, halo.data(), halo.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tb;
  if (make_tmap_2d(&tb, db, 64, 128, 128, 128)) {
    printf("tensor map failed: %s\n", tmap_last_error());
    return 1;
  }
  cudaFuncSetAttribute(halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48000);
  int ok_abs = 0, ok_rel = 0, taps = 0;
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int dh = 0; dh < 3; ++dh)
      for (int dw = 0; dw < 3; ++dw) {
        cudaMemset(dout, 0xFF, 128 * 128 * 4);
        halo_kernel<<<1, 128, 48000>>>(tb, dhalo, dout, dh, dw, use_bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("kernel failed: %s\n", cudaGetErrorString(e));
          return 2;
        }
        std::vector<float> out(128 * 128);
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        double max_err = 0;
        for (int m = 0; m < 128; ++m) {
          const int hr = (m / 8 + dh) * kHaloW + (m % 8 + dw);
          for (int n = 0; n < 128; ++n) {
            double acc = 0;
            for (int k = 0; k < 64; ++k) acc += static_cast<double>(bf2f(halo[hr * 64 + k])) * bf2f(b[n * 64 + k]);
            const double err = std::fabs(acc - out[m * 128 + n]);
            if (!(err <= max_err)) max_err = err;
          }
        }
        const bool ok = max_err < 1e-3;
        printf("tap (%d,%d) base_offset field %s: max_err %.3e %s\n", dh, dw, use_bo ? "set " : "zero", max_err,
               ok ? "OK" : "MISMATCH");
        (use_bo ? ok_rel : ok_abs) += ok ? 1 : 0;
        taps += use_bo ? 0 : 1;
      }
  printf("halo tile via shifted SW128 descriptors: %d/%d taps exact with base_offset = 0, %d/%d with it set\n", ok_abs,
         taps, ok_rel, taps);
  return 0;
}
