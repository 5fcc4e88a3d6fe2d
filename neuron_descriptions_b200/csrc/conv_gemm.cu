// tcgen05 implicit-GEMM convolution kernel (see conv_gemm.h for the contract).
//
// CTA = 11 warps, persistent over output tiles (static round-robin):
//   warp 0      : TMA producer  (A boxes per tap / k-block, B weight tiles) -> smem ring, mbarrier full/empty
//   warp 1      : TMEM allocator + tcgen05.mma issuer -> fp32 accumulators in TMEM (ring of 2-4 buffers)
//   warps 2..9  : epilogue, two groups of 4 warps (one warp per TMEM lane quarter and group; group g owns columns
//                 [32g, 32g+32) of every 64-column chunk): tcgen05.ld -> +bias (+residual) (ReLU) -> bf16 hi/lo
//                 split -> swizzled smem -> TMA store (EPI_BF16), or fp32 direct stores (EPI_F32)
//   warp 10     : residual prefetcher: TMA-loads the residual tile of each 64-column chunk into a 2-deep ring
// The issuing warps (0, 1, 10 and the storing warp 2) run their loops on all 32 lanes with uniform control flow and the
// `*_elect` helpers of ptx.cuh pick the lane that issues: inside an `if (lane == 0)` region ptxas wraps every
// uniform-register operand of UTCHMMA / UTMALDG in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~120 cycles per MMA).
// Two-level accumulation: the tensor core's fp32 accumulation truncates (measured: error grows linearly with the
// number of chained MMAs, 2.5e-5 abs at K=4544), so the MMA warp starts a fresh TMEM accumulator every
// `kb_per_chunk` k-blocks and the epilogue warps add the partials in fp32 registers (round-to-nearest). The TMEM
// ring (512 columns: 4 buffers, or 2 double-width ones in the WIDE form) lets partial drains and the final epilogue
// overlap the next chunks' MMAs.
#include "conv_gemm.h"
#include "ptx.cuh"

#include <atomic>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace milan {

namespace {

constexpr int kNumThreads = 352;
constexpr int kEpiThreads = 256;
constexpr int kTileBytes = kGemmBlockM * 64 * 2;  // 16 KiB: one 128-row x 64-col bf16 plane (epilogue tiles)
constexpr int kSmemBudget = 224 * 1024;

template <int BLOCK_N, bool SPLIT, int EPI, bool HAS_RES, int BK>
struct SmemLayout {
  // one A plane per stage. BK = 64: a 16 KiB SW128 tile per k-block. BK = 32 is the stem: ONE stage per TILE holding
  // the 38 raw padded-row segments (176 B each, 6688 B) that all seven filter rows of the tile read through
  // overlapping-window descriptors (filter row r starts r segments in; image rows are two segments apart).
  static constexpr int kABytes = BK == 32 ? 7168 : kGemmBlockM * BK * 2;
  static constexpr int kBBytes = BLOCK_N * BK * 2;
  static constexpr int kPlanes = SPLIT ? 2 : 1;
  // The 64-byte-k-block variant is the 7x7 stem: its whole weight matrix (7 k-blocks x 64 rows, 56 KB as hi + lo)
  // stays resident in shared memory for the life of the CTA, the stages carry the A windows only. Re-fetching it
  // for each of the 94 080 tiles of a 960-image step was a third of the kernel's L2 -> SM traffic.
  static constexpr bool kResidentB = BK == 32;
  static constexpr int kResidentKb = kResidentB ? 7 : 0;
  static constexpr int kResidentBytes = kResidentKb * kPlanes * kBBytes;
  static constexpr int kStageBytes = kPlanes * (kABytes + (kResidentB ? 0 : kBBytes));
  // one 64-column chunk: bf16 hi + lo planes for the TMA stores, or (EPI_HEAD) the same 32 KiB as 128 x 64 fp32 through
  // which the epilogue transposes its thread-per-row registers into row-contiguous global stores
  static constexpr int kStagingBytes = EPI == EPI_BF16 ? kPlanes * kTileBytes : (EPI == EPI_HEAD ? 2 * kTileBytes : 0);
  static constexpr int kResBytes = (EPI == EPI_BF16 && HAS_RES) ? 2 * kPlanes * kTileBytes : 0;  // 2-deep ring
  static constexpr int kStagesRaw = (kSmemBudget - kStagingBytes - kResBytes - kResidentBytes) / kStageBytes;
#ifndef MILAN_MAX_STAGES
#define MILAN_MAX_STAGES 8  // (-DMILAN_MAX_STAGES=n: pipeline-depth sensitivity experiments)
#endif
  static constexpr int kStages = kStagesRaw > MILAN_MAX_STAGES ? MILAN_MAX_STAGES : kStagesRaw;
  static constexpr int kBarrierBytes = 512;
  static constexpr int kTotalBytes =
      kStages * kStageBytes + kStagingBytes + kResBytes + kResidentBytes + kBarrierBytes + 1024;
  static_assert(kStages >= 2, "not enough shared memory for a pipelined main loop");
  static_assert(kTotalBytes <= 227 * 1024, "shared memory budget exceeded");
};

// byte offset of 16-byte chunk j of row r inside a [128][128 B] tile with the TMA/UMMA 128-byte swizzle
__device__ __forceinline__ uint32_t swz(int r, int j) { return static_cast<uint32_t>(r) * 128u + ((j ^ (r & 7)) << 4); }

// byte offset of 16-byte chunk j (0..15) of row r inside a [128][256 B] fp32 staging tile: the same XOR on the low three
// chunk bits, so that 8 consecutive rows writing the same chunk, and one row read across a warp, are conflict-free
__device__ __forceinline__ uint32_t swz256(int r, int j) { return static_cast<uint32_t>(r) * 256u + ((j ^ (r & 7)) << 4); }

template <int BLOCK_N, bool SPLIT, int EPI, bool HAS_RES, int BK, bool WIDE>
__global__ void __launch_bounds__(kNumThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p,
                                                                   const int* __restrict__ skip_flag) {
  if (skip_flag != nullptr && *skip_flag != 0) return;  // uniform: whole grid exits before touching any barrier
  using L = SmemLayout<BLOCK_N, SPLIT, EPI, HAS_RES, BK>;
  constexpr int kABytes = L::kABytes;
  constexpr int kStages = L::kStages;
  // WIDE (split mode, long-K problems without a residual): TWO MMAs per K step instead of three. The hi and lo planes
  // of B are adjacent in shared memory, so A_hi x [B_hi; B_lo]^T is one N = 2 * BLOCK_N instruction whose result is
  // [hi*hi | hi*lo] side by side, and A_lo x B_hi^T accumulates into the first half; the epilogue adds the halves.
  // Same FLOPs, but A_hi is fetched from shared memory once instead of twice (20 instead of 24 KB of operand reads per
  // K step at BLOCK_N = 128) and shared-memory bandwidth - MMA operand reads plus the TMA fill - is what bounds the
  // long-K convs. Short-K / residual problems keep three MMAs and the deeper accumulator ring:
  // their epilogue is the critical path and would only see the doubled TMEM reads.
  static_assert(!WIDE || SPLIT, "the wide form only exists in split mode");
  constexpr bool wide = WIDE;
  constexpr int acc_cols = WIDE ? 2 * BLOCK_N : BLOCK_N;  // TMEM columns per accumulator buffer
  constexpr int kTmemBufs = (512 / acc_cols) < 4 ? (512 / acc_cols) : 4;  // accumulator ring
  constexpr int tmem_bufs = kTmemBufs;
  constexpr uint32_t kTmemCols = kTmemBufs * acc_cols;  // 512 or 256 columns (power of two)
  constexpr int kChunks = BLOCK_N / 64;        // 64-column epilogue chunks per tile

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages * L::kStageBytes;
  uint8_t* res_smem = staging + L::kStagingBytes;
  uint8_t* resident_b = res_smem + L::kResBytes;  // [kResidentKb][hi | lo] weight k-blocks (stem variant only)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(resident_b + L::kResidentBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + kTmemBufs;
  uint64_t* res_full_bar = tmem_empty_bar + kTmemBufs;
  uint64_t* res_empty_bar = res_full_bar + 2;
  uint64_t* resident_bar = res_empty_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(resident_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int pl = 0; pl < 4; ++pl) {
      tma_prefetch_desc(&p.tmap_a[0][pl]);
      if (SPLIT) tma_prefetch_desc(&p.tmap_a[1][pl]);
    }
    tma_prefetch_desc(&p.tmap_b[0]);
    if (SPLIT) tma_prefetch_desc(&p.tmap_b[1]);
    if (EPI == EPI_BF16) {
      tma_prefetch_desc(&p.tmap_out[0]);
      if (SPLIT) tma_prefetch_desc(&p.tmap_out[1]);
      if (HAS_RES) {
        tma_prefetch_desc(&p.tmap_res[0]);
        if (SPLIT) tma_prefetch_desc(&p.tmap_res[1]);
      }
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kTmemBufs; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kEpiThreads);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&res_full_bar[s], 1);
      mbar_init(&res_empty_bar[s], kEpiThreads);
    }
    mbar_init(resident_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int cin_blocks = p.cin / BK;
  int num_kb = 0;
  for (int t = 0; t < p.num_taps; ++t) num_kb += p.tap_cb[t] > 0 ? p.tap_cb[t] : cin_blocks;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = m_tiles * p.n_tiles;
  const int kb_per_chunk = (p.kb_per_chunk > 0 && p.kb_per_chunk < num_kb) ? p.kb_per_chunk : num_kb;
  const int num_chunks = (num_kb + kb_per_chunk - 1) / kb_per_chunk;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    {  // all 32 lanes, uniform control flow; elect.sync inside the helpers picks the issuing lane
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = L::kPlanes * (p.a_box_bytes + (L::kResidentB ? 0 : L::kBBytes));
      if (L::kResidentB) {  // the whole weight matrix once (n_tiles == 1, num_kb <= kResidentKb: checked at launch)
        mbar_arrive_expect_tx_elect(resident_bar, static_cast<uint32_t>(num_kb) * L::kPlanes * L::kBBytes);
        for (int kb = 0; kb < num_kb; ++kb) {
          uint8_t* dst = resident_b + kb * (L::kPlanes * L::kBBytes);
          tma_load_2d_elect(dst, &p.tmap_b[0], resident_bar, kb * BK, 0);
          if (SPLIT) tma_load_2d_elect(dst + L::kBBytes, &p.tmap_b[1], resident_bar, kb * BK, 0);
        }
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.box_w, h0 = th * p.box_h, n0 = tn * p.box_n;
        if (BK == 32) {
          // stem: the tile's raw input once - padded rows 2 h0 .. 2 h0 + 37 (both parities of 19 row pairs), the 22
          // pixels its 8 output columns read. One box per plane instead of one per filter row: the filter rows overlap
          // (row 2h + r serves (h, r) and (h + 1, r - 2)), so this is a third of the bytes and of the narrow 176-byte
          // row requests, which is what bounded the stem (ncu: MMA warp starved, profiles/r02o_stem_full_summary.txt).
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * L::kStageBytes;
          mbar_arrive_expect_tx_elect(&full_bar[stage], L::kPlanes * p.a_box_bytes);
          tma_load_4d_elect(st, &p.tmap_a[0][0], &full_bar[stage], 8 * w0, 0, h0, n0);
          if (SPLIT) tma_load_4d_elect(st + kABytes, &p.tmap_a[1][0], &full_bar[stage], 8 * w0, 0, h0, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          continue;
        }
        int kcoord = 0;  // running K coordinate into the weight matrix
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int plane = p.tap_plane[tap];
          const int cw = w0 + p.tap_dw[tap];
          const int ch = h0 + p.tap_dh[tap];
          const int tap_blocks = p.tap_cb[tap] > 0 ? p.tap_cb[tap] : cin_blocks;
          for (int cb = 0; cb < tap_blocks; ++cb, kcoord += BK) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + stage * L::kStageBytes;
#ifdef MILAN_PROBE_SKIP  // fill-rate experiments: 1 = no B loads, 2 = no A loads, 3 = neither
            const bool probe_a = (MILAN_PROBE_SKIP & 2) == 0, probe_b = (MILAN_PROBE_SKIP & 1) == 0;
            mbar_arrive_expect_tx_elect(&full_bar[stage], L::kPlanes * ((probe_a ? p.a_box_bytes : 0u) +
                                                                   ((probe_b && !L::kResidentB) ? L::kBBytes : 0u)));
            if (!probe_a) {
            } else
#else
            constexpr bool probe_b = true;
            mbar_arrive_expect_tx_elect(&full_bar[stage], tx_bytes);
#endif
            {
              tma_load_4d_elect(st, &p.tmap_a[0][plane], &full_bar[stage], cb * BK, cw, ch, n0);
              if (SPLIT)
                tma_load_4d_elect(st + kABytes, &p.tmap_a[1][plane], &full_bar[stage], cb * BK, cw, ch, n0);
            }
            if (!L::kResidentB && probe_b) {
              uint8_t* sb = st + L::kPlanes * kABytes;
              tma_load_2d_elect(sb, &p.tmap_b[0], &full_bar[stage], kcoord, n_tile * BLOCK_N);
              if (SPLIT) tma_load_2d_elect(sb + L::kBBytes, &p.tmap_b[1], &full_bar[stage], kcoord, n_tile * BLOCK_N);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // All 32 lanes run this loop with uniform control flow; elect.sync inside the helpers picks the issuing lane
    // (see umma_bf16_elect in ptx.cuh: a divergent single-thread region costs ~120 cycles per MMA in operand shuffling).
    {
      const uint32_t idesc = make_idesc_16bit(kGemmBlockM, BLOCK_N, p.fp16_operands ? 0u : 1u);
      const uint32_t idesc_wide = make_idesc_16bit(kGemmBlockM, 2 * BLOCK_N, p.fp16_operands ? 0u : 1u);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t cc = 0;  // running accumulator-chunk counter: buffer = cc % kTmemBufs, phase = (cc / kTmemBufs) & 1
      if (L::kResidentB) mbar_wait(resident_bar, 0);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int kb = 0;
        for (int chunk = 0; chunk < num_chunks; ++chunk, ++cc) {
          const int as = cc % tmem_bufs;
          const uint32_t aphase = (cc / tmem_bufs) & 1;
          mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + as * acc_cols;
          const int kb_end = (kb + kb_per_chunk < num_kb) ? kb + kb_per_chunk : num_kb;
          const int kb_first = kb;
          for (; kb < kb_end; ++kb) {
            if (BK != 32 || kb == 0) {  // (stem: one stage per tile, filter row kb starts kb segments into it)
              mbar_wait(&full_bar[stage], phase);
              tcgen05_fence_after();
            }
            const uint32_t a_hi = smem_u32(smem + stage * L::kStageBytes) + (BK == 32 ? kb * kStemSegBytes : 0);
            const uint32_t b_hi = L::kResidentB ? smem_u32(resident_b + kb * (L::kPlanes * L::kBBytes))
                                                : a_hi + L::kPlanes * kABytes;
            // BK == 32 is the stem: A is not an im2col tile but the raw input rows, addressed as overlapping windows
            const uint64_t da_hi = BK == 32 ? make_smem_desc_stem_tile(a_hi) : make_smem_desc_k<BK>(a_hi);
            const uint64_t db_hi = make_smem_desc_k<BK>(b_hi);
            const uint64_t db_lo = make_smem_desc_k<BK>(b_hi + L::kBBytes);
            const uint64_t da_lo =
                BK == 32 ? make_smem_desc_stem_tile(a_hi + kABytes) : make_smem_desc_k<BK>(a_hi + kABytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t koff = 2 * k;  // 16 bf16 = 32 B = 2 x 16-byte units
              const uint32_t accumulate = (kb > kb_first || k > 0) ? 1u : 0u;
              if (wide) {  // db_hi spans the hi rows and, right behind them, the lo rows
                umma_bf16_elect(tmem_d, da_hi + koff, db_hi + koff, idesc_wide, accumulate);
                umma_bf16_elect(tmem_d, da_lo + koff, db_hi + koff, idesc, 1u);
              } else {
                umma_bf16_elect(tmem_d, da_hi + koff, db_hi + koff, idesc, accumulate);
                if (SPLIT) {
                  umma_bf16_elect(tmem_d, da_lo + koff, db_hi + koff, idesc, 1u);
                  umma_bf16_elect(tmem_d, da_hi + koff, db_lo + koff, idesc, 1u);
                }
              }
            }
            const bool stage_done = BK != 32 || kb == num_kb - 1;
            if (stage_done) umma_commit_elect(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
            if (kb == kb_end - 1) umma_commit_elect(&tmem_full_bar[as]);
            if (stage_done && ++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ------------------------------------------------------------ residual prefetcher
    if (EPI == EPI_BF16 && HAS_RES) {  // all lanes, uniform
      int rb = 0;
      uint32_t rphase = 0;
      const uint32_t tx_bytes = L::kPlanes * p.a_box_bytes;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        for (int c = 0; c < kChunks; ++c) {
          const int col0 = n_tile * BLOCK_N + c * 64;
          if (col0 >= p.cout) break;
          mbar_wait(&res_empty_bar[rb], rphase ^ 1);
          uint8_t* dst = res_smem + rb * (L::kPlanes * kTileBytes);
          mbar_arrive_expect_tx_elect(&res_full_bar[rb], tx_bytes);
          tma_load_4d_elect(dst, &p.tmap_res[0], &res_full_bar[rb], col0, tw * p.box_w, th * p.box_h, tn * p.box_n);
          if (SPLIT)
            tma_load_4d_elect(dst + kTileBytes, &p.tmap_res[1], &res_full_bar[rb], col0, tw * p.box_w, th * p.box_h,
                        tn * p.box_n);
          if (++rb == 2) { rb = 0; rphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int group = (warp - 2) >> 2;     // which 32-column half of each 64-column chunk
    const int row = quarter * 32 + lane;
    const bool leader = warp == 2;  // the warp whose elected lane issues the TMA stores (warp-uniform calls)
    uint32_t cc = 0;
    int rb = 0;
    uint32_t rphase = 0;
    // EPI_F32 only: row -> output pixel
    const int box_hw = p.box_w * p.box_h;
    const int dn = row / box_hw;
    const int rem = row - dn * box_hw;
    const int dh = rem / p.box_w;
    const int dw = rem - dh * p.box_w;
    const bool row_in_box = row < box_hw * p.box_n;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int tn = m_tile / (p.tiles_w * p.tiles_h);

      // ---- gather the accumulator: this thread owns columns c*64 + group*32 + [0,32) of row `row`
      float v[kChunks][32];
      for (int chunk = 0; chunk < num_chunks; ++chunk, ++cc) {
        const int as = cc % tmem_bufs;
        const uint32_t aphase = (cc / tmem_bufs) & 1;
        mbar_wait(&tmem_full_bar[as], aphase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * acc_cols + group * 32;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32(taddr + c * 64, acc);
          tmem_ld_wait();
          if (chunk == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c][j] = __uint_as_float(acc[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c][j] += __uint_as_float(acc[j]);
          }
          if (wide) {  // + the hi*lo half of the buffer
            tmem_ld_32x32(taddr + BLOCK_N + c * 64, acc);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c][j] += __uint_as_float(acc[j]);
          }
        }
        tcgen05_fence_before();
        mbar_arrive(&tmem_empty_bar[as]);  // partial drained: the MMA warp may reuse this buffer
      }

      if (EPI == EPI_BF16) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const int col0 = n_tile * BLOCK_N + c * 64;
          const bool active = col0 < p.cout;  // uniform across the CTA
          if (active && HAS_RES) mbar_wait(&res_full_bar[rb], rphase);
          // previous TMA store must have finished reading the staging tile before it is overwritten
          if (leader) tma_store_wait_read_elect<0>();
          named_bar_sync(1, kEpiThreads);
          if (active) {
            const uint8_t* rsrc = res_smem + rb * (L::kPlanes * kTileBytes);
            if (p.bias != nullptr) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0 + group * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                v[c][4 * j + 0] += b.x; v[c][4 * j + 1] += b.y; v[c][4 * j + 2] += b.z; v[c][4 * j + 3] += b.w;
              }
            }
            if (HAS_RES) {
#pragma unroll
              for (int pl = 0; pl < L::kPlanes; ++pl) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 r = *reinterpret_cast<const uint4*>(rsrc + pl * kTileBytes + swz(row, group * 4 + j));
                  v[c][8 * j + 0] += bf16_lo_to_f32(r.x); v[c][8 * j + 1] += bf16_hi_to_f32(r.x);
                  v[c][8 * j + 2] += bf16_lo_to_f32(r.y); v[c][8 * j + 3] += bf16_hi_to_f32(r.y);
                  v[c][8 * j + 4] += bf16_lo_to_f32(r.z); v[c][8 * j + 5] += bf16_hi_to_f32(r.z);
                  v[c][8 * j + 6] += bf16_lo_to_f32(r.w); v[c][8 * j + 7] += bf16_hi_to_f32(r.w);
                }
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[c][j] = fmaxf(v[c][j], 0.0f);
            }
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split_bf16x2(v[c][2 * j], v[c][2 * j + 1], hi[j], lo[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              *reinterpret_cast<uint4*>(staging + swz(row, group * 4 + j)) =
                  make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
              if (SPLIT)
                *reinterpret_cast<uint4*>(staging + kTileBytes + swz(row, group * 4 + j)) =
                    make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
            if (HAS_RES) {
              mbar_arrive(&res_empty_bar[rb]);  // residual chunk consumed
              if (++rb == 2) { rb = 0; rphase ^= 1; }
            }
          }
          fence_proxy_async();  // make the generic-proxy smem writes visible to the TMA engine
          named_bar_sync(1, kEpiThreads);
          if (leader && active) {
            tma_store_4d_elect(&p.tmap_out[0], staging, col0, tw * p.box_w, th * p.box_h, tn * p.box_n);
            if (SPLIT)
              tma_store_4d_elect(&p.tmap_out[1], staging + kTileBytes, col0, tw * p.box_w, th * p.box_h, tn * p.box_n);
            tma_store_commit_elect();
          }
        }
      } else if (EPI == EPI_LSTM) {
        // LSTM cell on the gate pre-activations this thread holds: columns 4u + g, g = (i, f, g, o) of unit u.
        const FusedEpilogue& fe = p.fe;
        const long long r = static_cast<long long>(tw) * p.box_w + row;  // flat GEMM: one output row per A row
        if (r < p.out_w) {
          const long long src = fe.src_row != nullptr ? fe.src_row[r] : r;
          const float* add_row = nullptr;
          if (fe.add_table != nullptr)
            add_row = fe.add_table +
                      (fe.add_index != nullptr ? fe.add_index[r * fe.add_stride] : fe.add_const) * fe.add_pitch;
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            const int col0 = n_tile * BLOCK_N + c * 64 + group * 32;
            if (col0 < p.cout) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                v[c][4 * j + 0] += b.x; v[c][4 * j + 1] += b.y; v[c][4 * j + 2] += b.z; v[c][4 * j + 3] += b.w;
              }
              if (add_row != nullptr) {
                const float4* a4 = reinterpret_cast<const float4*>(add_row + col0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b = __ldg(a4 + j);
                  v[c][4 * j + 0] += b.x; v[c][4 * j + 1] += b.y; v[c][4 * j + 2] += b.z; v[c][4 * j + 3] += b.w;
                }
              }
              const int u0 = col0 >> 2;  // first of this chunk's 8 hidden units
              const float4* cp = reinterpret_cast<const float4*>(fe.c_in + src * fe.hidden + u0);
              const float4 c_a = cp[0], c_b = cp[1];
              const float c_prev[8] = {c_a.x, c_a.y, c_a.z, c_a.w, c_b.x, c_b.y, c_b.z, c_b.w};
              float c_new[8], h_new[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float gi = v[c][4 * j], gf = v[c][4 * j + 1], gg = v[c][4 * j + 2], go = v[c][4 * j + 3];
                c_new[j] = sigmoid_fast(gf) * c_prev[j] + sigmoid_fast(gi) * tanh_fast(gg);
                h_new[j] = sigmoid_fast(go) * tanh_fast(c_new[j]);
              }
              float4* co = reinterpret_cast<float4*>(fe.c_out + r * fe.hidden + u0);
              co[0] = make_float4(c_new[0], c_new[1], c_new[2], c_new[3]);
              co[1] = make_float4(c_new[4], c_new[5], c_new[6], c_new[7]);
              if (fe.h_f32 != nullptr) {
                float4* ho = reinterpret_cast<float4*>(fe.h_f32 + r * fe.hidden + u0);
                ho[0] = make_float4(h_new[0], h_new[1], h_new[2], h_new[3]);
                ho[1] = make_float4(h_new[4], h_new[5], h_new[6], h_new[7]);
              }
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) split_fp16x2(h_new[2 * j], h_new[2 * j + 1], hi[j], lo[j]);
              *reinterpret_cast<uint4*>(fe.h_hi + r * fe.h_pitch + u0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              if (SPLIT) *reinterpret_cast<uint4*>(fe.h_lo + r * fe.h_pitch + u0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
      } else if (EPI == EPI_HEAD) {
        const FusedEpilogue& fe = p.fe;
        const long long r = static_cast<long long>(tw) * p.box_w + row;
        const bool valid = r < p.out_w;
        // One 64-column chunk of the tile -> global memory, row-contiguous: every thread holds 32 columns of ITS row, so
        // storing from registers writes 16 bytes to each of 32 different rows per instruction (120 MB of such stores per
        // beam step made the head GEMM 57 % slower per tile than the LM head, which stores nothing). Through the
        // staging tile instead: each warp then writes whole 256-byte row segments. All epilogue threads call this
        // (two named barriers); rows past M and columns past `col_limit` are masked at the store.
        auto store_rows = [&](const float (&vals)[32], float* dst, long long pitch, int col0, int col_limit) {
          named_bar_sync(1, kEpiThreads);  // the previous chunk has been read back
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(staging + swz256(row, group * 8 + j)) =
                make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
          named_bar_sync(1, kEpiThreads);
          const int ew = warp - 2;  // 0..7: rows [16 ew, 16 ew + 16) of the tile
          const int col = col0 + 2 * lane;
          const long long row0 = static_cast<long long>(tw) * p.box_w + ew * 16;
#pragma unroll 4
          for (int rr = 0; rr < 16; ++rr) {
            const int rl = ew * 16 + rr;
            const float2 x = *reinterpret_cast<const float2*>(staging + swz256(rl, lane >> 1) + (lane & 1) * 8);
            if (row0 + rr < p.out_w) {
              float* o = dst + (row0 + rr) * pitch + col;
              if (col + 1 < col_limit) *reinterpret_cast<float2*>(o) = x;
              else if (col < col_limit) *o = x.x;
            }
          }
        };
        const int tile_col0 = n_tile * BLOCK_N + group * 32;  // chunk c adds 64 c
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {  // + bias (padded to whole tiles)
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + tile_col0 + c * 64);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[c][4 * j + 0] += b.x; v[c][4 * j + 1] += b.y; v[c][4 * j + 2] += b.z; v[c][4 * j + 3] += b.w;
          }
        }
        if (n_tile < fe.vocab_tiles) {
          // vocabulary logits: optional store, softmax partial of this thread's 64 columns, optional target gather
          float mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            const int col0 = tile_col0 + c * 64;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (col0 + j >= fe.vocab) v[c][j] = -INFINITY;
              mx = fmaxf(mx, v[c][j]);
            }
          }
          float sum = 0.f;
          if (mx > -INFINITY) {
#pragma unroll
            for (int c = 0; c < kChunks; ++c)
#pragma unroll
              for (int j = 0; j < 32; ++j) sum += expf(v[c][j] - mx);
          }
          if (valid) {
            fe.partials[r * (2 * fe.vocab_tiles) + n_tile * 2 + group] = make_float2(mx, sum);
            if (fe.target != nullptr) {
              const int rel = static_cast<int>(fe.target[r * fe.target_stride]) - tile_col0;  // column within this thread's view
              if (rel >= 0 && rel < 96 && (rel & 32) == 0) {  // [0, 32) -> chunk 0, [64, 96) -> chunk 1
                float t = 0.f;
#pragma unroll
                for (int c = 0; c < kChunks; ++c)
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (rel == c * 64 + j) t = v[c][j];
                fe.tgt_logit[r] = t;
              }
            }
          }
          if (fe.logits != nullptr) {  // uniform
#pragma unroll
            for (int c = 0; c < kChunks; ++c)
              store_rows(v[c], fe.logits, fe.ld_logits, n_tile * BLOCK_N + c * 64, fe.vocab);
          }
        } else {
          const bool is_q = n_tile < fe.vocab_tiles + fe.q_tiles;
          const int base = is_q ? (n_tile - fe.vocab_tiles) * BLOCK_N : (n_tile - fe.vocab_tiles - fe.q_tiles) * BLOCK_N;
          const int limit = is_q ? fe.q_cols : fe.gate_cols;
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            if (base + c * 64 < limit) {  // uniform
              if (!is_q) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[c][j] = sigmoid_fast(v[c][j]);
              }
              store_rows(v[c], is_q ? fe.q_out : fe.g_out, is_q ? fe.q_pitch : fe.g_pitch, base + c * 64, limit);
            }
          }
        }
      } else {  // EPI_F32: direct fp32 stores, 128 B contiguous per thread per 32-column chunk
        const int w = tw * p.box_w + dw, h = th * p.box_h + dh, n = tn * p.box_n + dn;
        const bool valid = row_in_box && w < p.out_w && h < p.out_h && n < p.out_n;
        const long long pix = (static_cast<long long>(n) * p.out_h + h) * p.out_w + w;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const int col0 = n_tile * BLOCK_N + c * 64 + group * 32;
          if (valid && col0 < p.cout) {
            if (p.bias != nullptr) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                v[c][4 * j + 0] += b.x; v[c][4 * j + 1] += b.y; v[c][4 * j + 2] += b.z; v[c][4 * j + 3] += b.w;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[c][j] = fmaxf(v[c][j], 0.0f);
            }
            float* o = p.out_f32 + pix * p.ldc + col0;
            if (col0 + 32 <= p.cout) {
              float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                o4[j] = make_float4(v[c][4 * j], v[c][4 * j + 1], v[c][4 * j + 2], v[c][4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.cout) o[j] = v[c][j];
            }
          }
        }
      }
    }
    if (EPI == EPI_BF16 && leader) tma_store_wait_all_elect<0>();  // all stores complete before the CTA exits
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

std::atomic<long long> g_launches{0};
std::atomic<long long> g_all_launches{0};

template <int BLOCK_N, bool SPLIT, int EPI, bool HAS_RES, int BK, bool WIDE>
int launch_impl(const ConvGemmParams& p, int num_sms, cudaStream_t stream, const int* skip_flag) {
  using L = SmemLayout<BLOCK_N, SPLIT, EPI, HAS_RES, BK>;
  auto kernel = conv_gemm_kernel<BLOCK_N, SPLIT, EPI, HAS_RES, BK, WIDE>;
  static bool configured = false;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotalBytes);
      if (e != cudaSuccess) return static_cast<int>(e);
      configured = true;
    }
  }
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  if (total_tiles <= 0) return 0;
  if (L::kResidentB) {  // resident weights: one N tile, every k-block of it in the resident region
    int num_kb = 0;
    for (int t = 0; t < p.num_taps; ++t) num_kb += p.tap_cb[t] > 0 ? p.tap_cb[t] : p.cin / BK;
    if (p.n_tiles != 1 || num_kb > L::kResidentKb) return static_cast<int>(cudaErrorInvalidValue);
  }
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  kernel<<<grid, kNumThreads, L::kTotalBytes, stream>>>(p, skip_flag);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  g_all_launches.fetch_add(1, std::memory_order_relaxed);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

long long conv_gemm_launch_count() { return g_launches.load(); }
long long total_launch_count() { return g_all_launches.load(); }
void note_launch(int n) { g_all_launches.fetch_add(n, std::memory_order_relaxed); }
void count_conv_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  g_all_launches.fetch_add(1, std::memory_order_relaxed);
}

int launch_conv_gemm(const ConvGemmParams& p, int block_n, int split, int epilogue, int num_sms,
                     cudaStream_t stream, const int* skip_flag) {
  const bool res = p.has_res != 0;
  const int bk = p.block_k == 32 ? 32 : 64;
  // Two wide MMAs per K step (kernel comment) when the main loop, not the epilogue, is the long pole.
  bool wide = false;
  if (split != 0 && !res) {
    int num_kb = 0;
    for (int t = 0; t < p.num_taps; ++t) num_kb += p.tap_cb[t] > 0 ? p.tap_cb[t] : p.cin / bk;
    static const int min_k = [] {
      const char* e = getenv("MILAN_WIDE_SPLIT_MIN_K");  // experiment knob; 0 disables the wide form
      return e != nullptr ? atoi(e) : 256;  // measured: 256 > 512 > 1024 > off (bench.py, same box)
    }();
    // fp32-output GEMMs (decoder / LM) store 4 bytes per element from the epilogue: only the K = 4544 LSTM input GEMM is
    // main-loop-bound (ncu, one decode step: K = 512 GEMMs 53 -> 62 us with the wide form, K = 4544 172 -> 160 us)
    const int need_k = epilogue != EPI_BF16 ? 8 * min_k : min_k;
    // The stem (K = 224) too, since its A operand comes as one box per tile: it is bound by how fast one warp can
    // issue N = 64 MMAs (~70 cycles of descriptor arithmetic each against 48 of tensor time), and the wide form
    // issues 28 instead of 42 per tile: 0.334 -> 0.309 ms per 240 images (with per-filter-row boxes it was TMA-bound
    // and three MMAs were marginally better, 0.370 vs 0.382).
    wide = min_k > 0 && (num_kb * bk >= need_k || p.stem_mode);
  }
  static const bool pair_mode = [] {
    // cta_group::2 pairs for the long-K bf16 convs (conv_gemm_pair.cu): default since round 2 (full parity suite green,
    // same-box bench 1048 -> 1081 neurons/s, profiles/r02a_*); MILAN_PAIR=0 falls back to single-CTA tiles for A/B runs
    const char* e = getenv("MILAN_PAIR");
    return e == nullptr || atoi(e) != 0;
  }();
  if (pair_mode && wide && block_n == 128 && epilogue == EPI_BF16 && bk == 64 && p.has_b_half && !p.stem_mode)
    return launch_conv_gemm_pair(p, num_sms, stream, skip_flag);
  static const bool pair_res_mode = [] {
    // the residual (1x1 expand) convs on CTA pairs with the in-place io buffers; MILAN_PAIR_RES=0 for A/B runs
    const char* e = getenv("MILAN_PAIR_RES");
    return e == nullptr || atoi(e) != 0;
  }();
  // K >= 256 only (layer3 / layer4): same box, 240 images, single CTA -> pair: K = 512 82 -> 69 us, K = 256 97 -> 92 us,
  // but K = 128 151 -> 159 us and K = 64 278 -> 293 us - the short-K expands are HBM-bound (6.3 of 6.55 TB/s) and only
  // pay for the pair's extra synchronisation (gpurun_out/r02h_res_ab.txt -> profiles/r02h_expand_pair_ab.txt)
  if (pair_mode && pair_res_mode && res && split != 0 && block_n == 128 && epilogue == EPI_BF16 && bk == 64 &&
      p.has_b_half && !p.stem_mode && p.cout % 128 == 0 && p.num_taps == 1 && p.cin >= 256)
    return launch_conv_gemm_pair(p, num_sms, stream, skip_flag);
#define MILAN_DISPATCH(BN, SP, EP, RS, BKV, WD)                                                          \
  if (block_n == BN && (split != 0) == SP && epilogue == EP && res == RS && bk == BKV && wide == WD)     \
    return launch_impl<BN, SP, EP, RS, BKV, WD>(p, num_sms, stream, skip_flag);
  MILAN_DISPATCH(128, true, EPI_BF16, false, 64, false)
  MILAN_DISPATCH(128, true, EPI_BF16, false, 64, true)
  MILAN_DISPATCH(128, true, EPI_BF16, true, 64, false)
  MILAN_DISPATCH(128, false, EPI_BF16, false, 64, false)
  MILAN_DISPATCH(128, false, EPI_BF16, true, 64, false)
  MILAN_DISPATCH(64, true, EPI_BF16, false, 64, false)
  MILAN_DISPATCH(64, true, EPI_BF16, false, 64, true)
  MILAN_DISPATCH(64, false, EPI_BF16, false, 64, false)
  MILAN_DISPATCH(64, true, EPI_BF16, true, 64, false)    // BasicBlock conv2 of layer1 (resnet18/34): 64 outputs + residual
  MILAN_DISPATCH(64, false, EPI_BF16, true, 64, false)
  MILAN_DISPATCH(64, true, EPI_BF16, false, 32, true)    // stem: raw-row A windows, 64-byte k-blocks of B (SWIZZLE_64B)
  MILAN_DISPATCH(64, true, EPI_BF16, false, 32, false)
  MILAN_DISPATCH(64, false, EPI_BF16, false, 32, false)
  MILAN_DISPATCH(128, true, EPI_F32, false, 64, false)
  MILAN_DISPATCH(128, true, EPI_F32, false, 64, true)
  MILAN_DISPATCH(128, false, EPI_F32, false, 64, false)
  MILAN_DISPATCH(128, true, EPI_LSTM, false, 64, false)
  MILAN_DISPATCH(128, true, EPI_LSTM, false, 64, true)
  MILAN_DISPATCH(128, true, EPI_HEAD, false, 64, false)
#undef MILAN_DISPATCH
  return static_cast<int>(cudaErrorInvalidValue);
}

// ---------------------------------------------------------------- tensor maps
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
char g_tmap_err[384] = "";

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}
}  // namespace

const char* tmap_last_error() { return g_tmap_err; }

int make_tmap_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    snprintf(g_tmap_err, sizeof g_tmap_err, "cuTensorMapEncodeTiled entry point unavailable");
    return -1;
  }
  cuuint64_t d[5];
  cuuint64_t s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = 1;
    if (i + 1 < rank) s[i] = strides_bytes[i];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), d, s,
                  b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 0    ? CU_TENSOR_MAP_SWIZZLE_NONE
                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                        : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    int n = snprintf(g_tmap_err, sizeof g_tmap_err, "cuTensorMapEncodeTiled(rank %d) failed: %d base=%p dims=(", rank,
                     static_cast<int>(r), base);
    for (int i = 0; i < rank && n < static_cast<int>(sizeof g_tmap_err) - 32; ++i)
      n += snprintf(g_tmap_err + n, sizeof g_tmap_err - n, "%llu,", (unsigned long long)dims[i]);
    n += snprintf(g_tmap_err + n, sizeof g_tmap_err - n, ") strides=(");
    for (int i = 0; i + 1 < rank && n < static_cast<int>(sizeof g_tmap_err) - 32; ++i)
      n += snprintf(g_tmap_err + n, sizeof g_tmap_err - n, "%llu,", (unsigned long long)strides_bytes[i]);
    n += snprintf(g_tmap_err + n, sizeof g_tmap_err - n, ") box=(");
    for (int i = 0; i < rank && n < static_cast<int>(sizeof g_tmap_err) - 32; ++i)
      n += snprintf(g_tmap_err + n, sizeof g_tmap_err - n, "%u,", box[i]);
    snprintf(g_tmap_err + n, sizeof g_tmap_err - n, ")");
    return static_cast<int>(r);
  }
  return 0;
}

int make_tmap_4d(CUtensorMap* out, const void* base, uint64_t c, uint64_t w, uint64_t h, uint64_t n,
                 uint64_t stride_w_bytes, uint64_t stride_h_bytes, uint64_t stride_n_bytes, uint32_t box_w,
                 uint32_t box_h, uint32_t box_n) {
  const uint64_t dims[4] = {c, w, h, n};
  const uint64_t strides[3] = {stride_w_bytes, stride_h_bytes, stride_n_bytes};
  const uint32_t box[4] = {static_cast<uint32_t>(kGemmBlockK), box_w, box_h, box_n};
  return make_tmap_nd(out, base, 4, dims, strides, box);
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t k, uint64_t rows, uint64_t pitch_bytes,
                 uint32_t box_rows, int block_k) {
  const uint64_t dims[2] = {k, rows};
  const uint64_t strides[1] = {pitch_bytes};
  const uint32_t box[2] = {static_cast<uint32_t>(block_k), box_rows};
  return make_tmap_nd(out, base, 2, dims, strides, box, block_k * 2);
}

void choose_box(int W, int H, int N, int* bw, int* bh, int* bn) {
  long long best_tiles = -1;
  int bbw = 1, bbh = 1, bbn = 1;
  for (int w = 1; w <= W && w <= kGemmBlockM; ++w) {
    for (int h = 1; h <= H && w * h <= kGemmBlockM; ++h) {
      int n = kGemmBlockM / (w * h);
      if (n > N) n = N;
      if (n > 256) n = 256;
      if (n < 1) continue;
      const long long tiles =
          static_cast<long long>((W + w - 1) / w) * ((H + h - 1) / h) * ((N + n - 1) / n);
      // Prefer fewer tiles; on ties prefer wider rows (longer contiguous runs in memory).
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && w > bbw)) {
        best_tiles = tiles;
        bbw = w; bbh = h; bbn = n;
      }
    }
  }
  *bw = bbw; *bh = bbh; *bn = bbn;
}

}  // namespace milan
