// Probe (not part of the library; `make probes`): cycles per tcgen05.mma (M = 128, K = 16, bf16) as a function of N, of
// where the A operand lives (shared memory vs tensor memory) and of HOW the instruction is issued, with nothing else
// touching shared memory. One CTA per SM. Findings on B200:
//   * issued from a single-thread region (`if (threadIdx.x == 0)`): 112-123 cycles for every N <= 128, SS or TS, one or
//     four accumulators; N = 256: 166 (SS) / 137 (TS). The floor is ptxas' ELECT / R2UR.BROADCAST / BRA.U.ANY loop
//     around every uniform-register operand of UTCHMMA in divergent code, not the tensor core.
//   * issued warp-uniformly (all 32 lanes run the loop, elect.sync picks the lane): 48 / 64 / 128 cycles for
//     N = 64 / 128 / 256 from shared memory (N = 64: the 128 B/clk shared-memory read rate), 32 / 64 / 128 with A in
//     tensor memory = the ideal N / 2. One runtime branch per MMA inside the loop costs ~45 cycles per MMA.
//   The last section (sustained TFLOP/s) still uses the single-thread form: it shows the floor under the power cap.
#include "ptx.cuh"

#include <cstdio>
#include <vector>

using namespace milan;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
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

// Warp-uniform issue: every lane of the warp runs the loop (so ptxas keeps descriptors in uniform registers without an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop per operand) and elect.sync inside umma_bf16_elect (ptx.cuh) picks the lane.
__device__ __forceinline__ void umma_bf16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TS is a template parameter on purpose: one runtime branch per MMA in this loop costs ~45 cycles per MMA.
template <bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel_uniform(int n, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + ((i * 2654435761u) & 0x00FF00FFu);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(tmem_ptr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x < 32) {  // whole warp, uniform
    const uint32_t idesc = make_idesc_16bit(128, n, 1u);
    const uint64_t da = make_smem_desc_sw128(smem_u32(a_s));
    const uint64_t db = make_smem_desc_sw128(smem_u32(b_s));
    const uint32_t tmem_a = tmem_base + 480;
    if (TS && threadIdx.x == 0)
      for (int k = 0; k < 4; ++k) tmem_cp_128x256b(tmem_a + 8 * k, da + 2 * k);
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS) umma_bf16_ts_elect(tmem_base, tmem_a + 8 * k, db + 2 * k, idesc, 1u);
        else umma_bf16_elect(tmem_base, da + 2 * k, db + 2 * k, idesc, 1u);
      }
    }
    if (threadIdx.x == 0) {
      umma_commit(&bars[0]);
      mbar_wait(&bars[0], 0);
      cycles[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// mode 0: SS; 1: TS; 2: SS with A and B un-swizzled "raw window" descriptors (LBO 16) for A
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int mode, int ndst, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;           // 128 x 64 bf16, 16 KB
  uint8_t* b_s = smem + 16384;   // up to 256 x 64 bf16, 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + ((i * 2654435761u) & 0x00FF00FFu);  // small positive bf16s
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(tmem_ptr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_16bit(128, n, 1u);
    const uint64_t da = make_smem_desc_sw128(smem_u32(a_s));
    const uint64_t db = make_smem_desc_sw128(smem_u32(b_s));
    const uint32_t tmem_a = tmem_base + 480;  // 32 columns: 128 x 64 bf16
    if (mode == 1)
      for (int k = 0; k < 4; ++k) tmem_cp_128x256b(tmem_a + 8 * k, da + 2 * k);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t tmem_d = tmem_base + (k % ndst) * n;  // ndst independent accumulators, round robin
        if (mode == 0) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, 1u);
        else umma_bf16_ts(tmem_d, tmem_a + 8 * k, db + 2 * k, idesc, 1u);
      }
    }
    umma_commit(&bars[0]);
    mbar_wait(&bars[0], 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sms * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52000);
  const int iters = 2000;
  for (int grid : {sms}) {
    for (int mode = 0; mode < 2; ++mode) {
      for (int n : {64, 128, 256})
      for (int ndst : {1, 2, 4}) {
        if (n * ndst > 448) continue;
        rate_kernel<<<grid, 128, 52000>>>(n, mode, ndst, 50, d);  // warm-up
        rate_kernel<<<grid, 128, 52000>>>(n, mode, ndst, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("failed: %s\n", cudaGetErrorString(e));
          return 1;
        }
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (auto v : h) mx = v > mx ? v : mx;
        printf("grid %3d  A from %s  N=%3d  %d accumulator(s) : %.1f cycles per 128xNx16 MMA (ideal %d)\n", grid,
               mode == 0 ? "smem" : "tmem", n, ndst, static_cast<double>(mx) / (iters * 4), n / 2);
      }
    }
  }
  cudaFuncSetAttribute(rate_kernel_uniform<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 52000);
  cudaFuncSetAttribute(rate_kernel_uniform<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 52000);
  for (int ts = 0; ts < 2; ++ts)
  for (int n : {64, 128, 256}) {
    if (ts) {
      rate_kernel_uniform<true><<<sms, 128, 52000>>>(n, 50, d);
      rate_kernel_uniform<true><<<sms, 128, 52000>>>(n, iters, d);
    } else {
      rate_kernel_uniform<false><<<sms, 128, 52000>>>(n, 50, d);
      rate_kernel_uniform<false><<<sms, 128, 52000>>>(n, iters, d);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("uniform failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    printf("warp-uniform issue (elect.sync)  A from %s  N=%3d : %.1f cycles per MMA (ideal %d)\n", ts ? "tmem" : "smem", n,
           static_cast<double>(mx) / (iters * 4), n / 2);
  }
  // Sustained throughput under the board's power cap: ~3 s of back-to-back launches per shape, last ~1 s timed.
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int n : {128, 256}) {
      const int it = 20000;  // 80 000 MMAs per launch: 5-7 ms
      for (int l = 0; l < 300; ++l) rate_kernel<<<sms, 128, 52000>>>(n, mode, 1, it, d);
      cudaEventRecord(e0);
      const int timed = 150;
      for (int l = 0; l < timed; ++l) rate_kernel<<<sms, 128, 52000>>>(n, mode, 1, it, d);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flop = 2.0 * 128 * n * 16 * 4.0 * it * timed * sms;
      std::vector<long long> h(sms);
      cudaMemcpy(h.data(), d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      printf("sustained  A from %s  N=%3d : %.0f TFLOP/s over %.0f ms (%.1f cycles per MMA, i.e. SM clock %.0f MHz)\n",
             mode == 0 ? "smem" : "tmem", n, flop / (ms * 1e-3) / 1e12, ms, static_cast<double>(h[0]) / (it * 4.0),
             static_cast<double>(h[0]) / (ms / timed * 1e-3) / 1e6);
    }
  }
  return 0;
}
