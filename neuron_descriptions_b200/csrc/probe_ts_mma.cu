// Probe (not part of the library; `make probes`): where can the A operand of tcgen05.mma come from? One
// 128 x 128 x 64 bf16 tile, D = A * B^T, checked against the CPU, in four ways:
//   0  A from the 128B-swizzled K-major shared-memory tile (what the conv kernel does)
//   1  A from TENSOR MEMORY, put there by tcgen05.cp.128x256b from that same tile (TS form)
//   2  A = overlapping 64-byte windows of ONE raw row: un-swizzled descriptor with LBO = 16 B, SBO = 128 B
//   3  the same with 8-row groups 176 B apart -- the layout the 7x7 stem now uses (make_smem_desc_stem_rows)
// All four are exact on B200, i.e. im2col can be done by the descriptor.
#include "conv_gemm.h"
#include "ptx.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

using namespace milan;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                       const __grid_constant__ CUtensorMap tmap_b, float* out,
                                                       const uint16_t* raw, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + 16384;
  uint8_t* raw_s = smem + 32768;  // modes 2/3: un-swizzled rows, 16 bytes per "pixel", overlapping 64-byte windows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768 + 4096);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (mode >= 2) {
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) reinterpret_cast<uint16_t*>(raw_s)[i] = raw[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], 32768);
    tma_load_2d(a_s, &tmap_a, &bars[0], 0, 0);
    tma_load_2d(b_s, &tmap_b, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tcgen05_fence_after();
    const uint32_t idesc = make_idesc_16bit(128, 128, 1u);
    const uint64_t da = make_smem_desc_sw128(smem_u32(a_s));
    const uint64_t db = make_smem_desc_sw128(smem_u32(b_s));
    const uint32_t tmem_d = tmem_base, tmem_a = tmem_base + 128;
    if (mode == 1) {
      for (int k = 0; k < 4; ++k) tmem_cp_128x256b(tmem_a + 8 * k, da + 2 * k);  // one K=16 slice = 8 columns
    }
    if (mode >= 2) {
      // K-major, no swizzle: 8 x 16-byte core-matrix rows are 16 B apart; the next K chunk (LBO) is ALSO 16 B on, so
      // row m reads raw[16 m, 16 m + 64): overlapping windows. SBO = pitch of an 8-row group (128 = one flat row of
      // pixels, 176 = 8-pixel groups with a 3-pixel halo each).
      uint64_t dr = static_cast<uint64_t>((smem_u32(raw_s) >> 4) & 0x3FFF);
      dr |= static_cast<uint64_t>(16 >> 4) << 16;
      dr |= static_cast<uint64_t>((mode == 2 ? 128 : 176) >> 4) << 32;
      dr |= static_cast<uint64_t>(1) << 46;
      for (int k = 0; k < 2; ++k) umma_bf16(tmem_d, dr + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
    }
    for (int k = 0; k < 4 && mode < 2; ++k) {
      if (mode == 0) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
      else umma_bf16_ts(tmem_d, tmem_a + 8 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bars[1]);
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
    tmem_dealloc(tmem_base, 256);
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
  std::mt19937 rng(1);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<uint16_t> a(128 * 64), b(128 * 64);
  for (auto& v : a) v = f2bf(nd(rng));
  for (auto& v : b) v = f2bf(nd(rng));
  uint16_t *da, *db;
  float* dout;
  cudaMalloc(&da, a.size() * 2);
  cudaMalloc(&db, b.size() * 2);
  cudaMalloc(&dout, 128 * 128 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, da, 64, 128, 128, 128) || make_tmap_2d(&tb, db, 64, 128, 128, 128)) {
    printf("tensor map failed: %s\n", tmap_last_error());
    return 1;
  }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  std::vector<uint16_t> raw(2048);
  for (auto& v : raw) v = f2bf(nd(rng));
  uint16_t* draw;
  cudaMalloc(&draw, raw.size() * 2);
  cudaMemcpy(draw, raw.data(), raw.size() * 2, cudaMemcpyHostToDevice);
  const char* names[4] = {"A from shared memory", "A from tensor memory via tcgen05.cp",
                          "A = overlapping 64-byte windows of one raw row (LBO 16, SBO 128)",
                          "A = overlapping windows, 8-row groups 176 B apart (LBO 16, SBO 176)"};
  int fails = 0;
  for (int mode = 0; mode < 4; ++mode) {
    cudaMemset(dout, 0xFF, 128 * 128 * 4);
    probe_kernel<<<1, 128, 40000>>>(ta, tb, dout, draw, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("mode %d: kernel failed: %s\n", mode, cudaGetErrorString(e));
      return 2;
    }
    std::vector<float> out(128 * 128);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double max_err = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        double acc = 0;
        if (mode < 2) {
          for (int k = 0; k < 64; ++k) acc += static_cast<double>(bf2f(a[m * 64 + k])) * bf2f(b[n * 64 + k]);
        } else {
          const int start = mode == 2 ? 8 * m : (m / 8) * 88 + (m % 8) * 8;
          for (int k = 0; k < 32; ++k) acc += static_cast<double>(bf2f(raw[start + k])) * bf2f(b[n * 64 + k]);
        }
        const double err = std::fabs(acc - out[m * 128 + n]);
        if (!(err <= max_err)) max_err = err;
      }
    printf("mode %d (%s): max_err %.3e %s\n", mode, names[mode], max_err, max_err < 1e-3 ? "OK" : "MISMATCH");
    fails += max_err < 1e-3 ? 0 : 1;
  }
  return fails;
}
