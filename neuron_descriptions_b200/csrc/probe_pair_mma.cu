// Probe (not part of the library; `make probes`): tcgen05.mma.cta_group::2 — a CTA pair (cluster of 2) computing one
// 256 x N x 64 tile, D = A * B^T. CTA r holds rows [128 r, 128 r + 128) of A and rows [N/2 r, N/2 r + N/2) of B in its
// own shared memory (same offsets in both CTAs, 128B swizzle); the leader CTA issues the MMAs, a multicast commit
// signals both CTAs, each CTA drains its own 128 TMEM lanes. Checks the result against the CPU and reports the issue
// rate: the point of the pair is that each SM reads only A + HALF of B per MMA.
#include "ptx.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

using namespace milan;

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {  // whole warp, in BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_elect(uint64_t* bar) {  // arrives on `bar` in both CTAs
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      ".reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t"
      "}"
      ::"r"(smem_u32(bar))
      : "memory");
}

// a: [256][64], b: [n][64] bf16 row-major in global memory; out: [256][n] fp32; cycles[cluster] for `iters` x 4 MMAs
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    pair_kernel(const uint16_t* a, const uint16_t* b, float* out, int n, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;          // 128 rows x 128 B
  uint8_t* b_s = smem + 16384;  // n/2 rows x 128 B (up to 16 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int half_n = n / 2;
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {  // A half, swizzled like TMA would
    const int r = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(a_s + r * 128 + ((j ^ (r & 7)) << 4)) =
        reinterpret_cast<const uint4*>(a)[(rank * 128 + r) * 8 + j];
  }
  for (int i = threadIdx.x; i < half_n * 8; i += blockDim.x) {  // B half
    const int r = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(b_s + r * 128 + ((j ^ (r & 7)) << 4)) =
        reinterpret_cast<const uint4*>(b)[(rank * half_n + r) * 8 + j];
  }
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_pair(tmem_ptr, 256);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs: operands in place, barriers initialised, TMEM allocated
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 0 && rank == 0) {  // leader CTA, warp-uniform issue
    const uint32_t idesc = make_idesc_16bit(256, n, 1u);
    const uint64_t da = make_smem_desc_sw128(smem_u32(a_s));
    const uint64_t db = make_smem_desc_sw128(smem_u32(b_s));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_pair_elect(tmem_base, da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
    }
    umma_commit_pair_elect(&bars[0]);
    mbar_wait(&bars[0], 0);
    if (lane == 0) cycles[blockIdx.x / 2] = clock64() - t0;
    __syncwarp();
  }
  mbar_wait(&bars[0], 0);  // multicast commit: signalled in both CTAs
  tcgen05_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < n / 32; ++c) {
    uint32_t acc[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, acc);
    tmem_ld_wait();
    if (blockIdx.x < 2)
      for (int j = 0; j < 32; ++j) out[(rank * 128 + row) * n + c * 32 + j] = __uint_as_float(acc[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may free TMEM / exit while the pair's MMAs or commits are in flight
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, 256);
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
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  std::mt19937 rng(5);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<uint16_t> a(256 * 64), b(256 * 64);
  for (auto& v : a) v = f2bf(nd(rng));
  for (auto& v : b) v = f2bf(nd(rng));
  uint16_t *da, *db;
  float* dout;
  long long* dcyc;
  cudaMalloc(&da, a.size() * 2);
  cudaMalloc(&db, b.size() * 2);
  cudaMalloc(&dout, 256 * 256 * 4);
  cudaMalloc(&dcyc, sms * sizeof(long long));
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 36000);
  for (int n : {128, 256}) {
    cudaMemset(dout, 0xFF, 256 * 256 * 4);
    pair_kernel<<<2, 128, 36000>>>(da, db, dout, n, 1, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("N=%d: kernel failed: %s\n", n, cudaGetErrorString(e));
      return 2;
    }
    std::vector<float> out(256 * n);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double max_err = 0;
    for (int m = 0; m < 256; ++m)
      for (int c = 0; c < n; ++c) {
        double acc = 0;
        for (int k = 0; k < 64; ++k) acc += static_cast<double>(bf2f(a[m * 64 + k])) * bf2f(b[c * 64 + k]);
        const double err = std::fabs(acc - out[m * n + c]);
        if (!(err <= max_err)) max_err = err;
      }
    printf("cta_group::2  256 x %d x 64: max_err %.3e %s\n", n, max_err, max_err < 1e-3 ? "OK" : "MISMATCH");
    const int grid = sms & ~1, iters = 2000;
    pair_kernel<<<grid, 128, 36000>>>(da, db, dout, n, iters, dcyc);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("N=%d: rate kernel failed: %s\n", n, cudaGetErrorString(e));
      return 3;
    }
    std::vector<long long> h(grid / 2);
    cudaMemcpy(h.data(), dcyc, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    printf("cta_group::2  256 x %d x 16 MMA on %d CTA pairs: %.1f cycles per MMA (each SM: 128 x %d, ideal %d)\n", n,
           grid / 2, static_cast<double>(mx) / (iters * 4), n, n / 2);
  }
  return 0;
}
