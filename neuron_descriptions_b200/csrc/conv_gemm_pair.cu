// CTA-pair (tcgen05 cta_group::2) variant of the implicit-GEMM convolution kernel for the bf16 convs with 128-wide
// N tiles: split (hi/lo) operands, bf16 hi/lo output through TMA stores, 64-wide k-blocks; with (RES) or without a
// residual. Default path for these convs (MILAN_PAIR=0 disables it, see launch_conv_gemm): DESIGN.md section 8c.
//
// A cluster of two CTAs (the two SMs of a TPC) computes two vertically adjacent M tiles of the same N tile as ONE
// 256 x 128 MMA per K step: CTA r loads its own 128 pixels of A (hi + lo) and rows [64 r, 64 r + 64) of the B tile
// (hi + lo); the leader (rank 0) issues tcgen05.mma.cta_group::2, which reads A from each CTA's own shared memory and
// B half from each, so per SM and K step the tensor core's operand reads fall from 24 to 18 KB and the TMA fill from
// 16 to 12 KB - shared-memory bandwidth is what bounds these convs.
//   * every TMA load of both CTAs (.cta_group::2) completes on the LEADER's full barrier (peer bit of the barrier
//     address cleared), which expects the bytes of both CTAs;
//   * tcgen05.commit.multicast frees the operand stage / publishes the accumulator in both CTAs;
//   * each CTA drains its own 128 TMEM lanes; the epilogue threads of both CTAs arrive on the leader's tmem_empty.
// Roles per CTA are those of conv_gemm.cu (warp 0 producer, warp 1 MMA - leader only -, warps 2..9 epilogue); all
// issue loops are warp-uniform (elect.sync).
//
// RES (the 1x1 expand convs, `relu(bn3(conv3(t)) + x)`): shared-memory bandwidth bounds these too - per 128 x 128
// tile the single-CTA kernel moves 896 KB through shared memory (operand fill 256, MMA operand reads 384, residual
// in + out staging 256) against 3072 tensor-core cycles; the pair halves B (fill 64, reads 288). The epilogue works IN
// PLACE on two 32 KB io buffers (one 64-column chunk as hi + lo planes each): the io warp (warp 10) TMA-loads the
// residual chunk into a buffer, the epilogue threads add their accumulators to the elements they own, split the sum
// back into the same bytes and arrive on `out_ready`; the io warp stores the buffer and, once that store has read it,
// refills it with the residual two chunks ahead. No separate residual ring and no CTA-wide named barriers; the 64 KB
// saved buy a third operand stage (the expands are bound by bytes in flight).
#include "conv_gemm.h"
#include "ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace milan {
namespace {

constexpr int kNumThreads = 352;
constexpr int kEpiThreads = 256;
constexpr int kBlockN = 128;
constexpr int kBK = 64;
constexpr int kABytes = kGemmBlockM * kBK * 2;        // 16 KB: one A plane
constexpr int kBHalfBytes = (kBlockN / 2) * kBK * 2;  // 8 KB: this CTA's half of one B plane
constexpr int kStageBytes = 2 * kABytes + 2 * kBHalfBytes;  // 48 KB
constexpr int kTileBytes = kGemmBlockM * 64 * 2;  // 16 KB: one 64-column output plane
constexpr int kStagingBytes = 2 * kTileBytes;     // hi + lo planes of one 64-column chunk
constexpr int kTmemBufs = 4;
constexpr uint32_t kTmemCols = kTmemBufs * kBlockN;  // 512
// STAGES operand stages of 48 KB + IO buffers of 32 KB. Without a residual: 4 + 1 (store staging). With one the io
// buffers are the in-place residual / result buffers: 3 + 2. (Measured and rejected, same box, 240 images: 2 + 4 -
// deeper residual prefetch, shallower operand ring - is 4-25 % slower on every expand conv, and pulling residual
// chunks into L2 ahead of the io buffers with cp.async.bulk.prefetch.tensor is neutral at 4 chunks and 10-35 % slower
// at 8: these convs already keep HBM as busy as their 128-byte-wide boxes allow.)
constexpr int kMaxIo = 4;
template <bool RES, int STAGES, int IO>
struct PairLayout {
  static constexpr int kStages = STAGES;
  static constexpr int kIoBufs = IO;
  static constexpr int kSmemBytes = kStages * kStageBytes + kIoBufs * kStagingBytes + 512 + 1024;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
  static_assert(IO <= kMaxIo && (RES || IO == 1), "io buffer count");
};
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: -> the leader

__device__ __forceinline__ uint32_t swz(int r, int j) { return static_cast<uint32_t>(r) * 128u + ((j ^ (r & 7)) << 4); }

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {  // one warp of EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA loads into THIS CTA's shared memory whose bytes complete on the LEADER's barrier at the same offset
__device__ __forceinline__ void tma_load_2d_pair_elect(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];\n\t"
      "}"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
      "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair_elect(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1,
                                                       int c2, int c3) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2];\n\t"
      "}"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
      "r"(c1), "r"(c2), "r"(c3)
      : "memory");
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
__device__ __forceinline__ void umma_commit_pair_elect(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs
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
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // the leader CTA's barrier at this offset
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// The 64-column chunks of this CTA's tiles, in the order the epilogue produces them (the io warp of the RES variant
// walks the same sequence twice: once loading residuals two chunks ahead, once storing results).
struct ChunkCursor {
  int tile, c, total_tiles, num_pairs, rank;
  __device__ ChunkCursor(int pair_id, int total_tiles_, int num_pairs_, int rank_)
      : tile(pair_id), c(0), total_tiles(total_tiles_), num_pairs(num_pairs_), rank(rank_) {}
  __device__ bool valid() const { return tile < total_tiles; }
  __device__ void next() {
    if (++c == kBlockN / 64) { c = 0; tile += num_pairs; }
  }
  __device__ void coords(const ConvGemmParams& p, int* col0, int* w0, int* h0, int* n0) const {
    const int n_tile = tile % p.n_tiles;
    const int m_tile = 2 * (tile / p.n_tiles) + rank;
    *col0 = n_tile * kBlockN + c * 64;
    *w0 = (m_tile % p.tiles_w) * p.box_w;
    *h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.box_h;
    *n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.box_n;
  }
};

template <bool RES, int STAGES, int IO>
__global__ void __launch_bounds__(kNumThreads, 1) conv_gemm_pair_kernel(const __grid_constant__ ConvGemmParams p,
                                                                        const int* __restrict__ skip_flag) {
  if (skip_flag != nullptr && *skip_flag != 0) return;  // uniform over the grid
  using L = PairLayout<RES, STAGES, IO>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages * kStageBytes;  // RES: io buffers [2][hi | lo]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + L::kIoBufs * kStagingBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + kTmemBufs;
  uint64_t* res_full_bar = tmem_empty_bar + kTmemBufs;  // [IO] RES: residual chunk landed in io buffer b (this CTA)
  uint64_t* out_ready_bar = res_full_bar + kMaxIo;      // [IO] RES: every epilogue thread finished io buffer b
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(out_ready_bar + kMaxIo);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_rank());

  if (warp == 0 && lane == 0) {
    for (int pl = 0; pl < 4; ++pl) {
      tma_prefetch_desc(&p.tmap_a[0][pl]);
      tma_prefetch_desc(&p.tmap_a[1][pl]);
    }
    tma_prefetch_desc(&p.tmap_b_half[0]);
    tma_prefetch_desc(&p.tmap_b_half[1]);
    tma_prefetch_desc(&p.tmap_out[0]);
    tma_prefetch_desc(&p.tmap_out[1]);
    if (RES) {
      tma_prefetch_desc(&p.tmap_res[0]);
      tma_prefetch_desc(&p.tmap_res[1]);
    }
    for (int s = 0; s < IO; ++s) {
      mbar_init(&res_full_bar[s], 1);
      mbar_init(&out_ready_bar[s], kEpiThreads);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);   // used in the leader only: its producer's arrive + the bytes of both CTAs
      mbar_init(&empty_bar[s], 1);  // multicast commit of the leader
    }
    for (int s = 0; s < kTmemBufs; ++s) {
      mbar_init(&tmem_full_bar[s], 1);                  // multicast commit of the leader
      mbar_init(&tmem_empty_bar[s], 2 * kEpiThreads);   // used in the leader only: epilogue threads of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_ptr_smem, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated in both
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int cin_blocks = p.cin / kBK;
  int num_kb = 0;
  for (int t = 0; t < p.num_taps; ++t) num_kb += p.tap_cb[t] > 0 ? p.tap_cb[t] : cin_blocks;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int m_pairs = (m_tiles + 1) / 2;
  const int total_tiles = m_pairs * p.n_tiles;  // pair tiles: two M tiles x one N tile
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int kb_per_chunk = (p.kb_per_chunk > 0 && p.kb_per_chunk < num_kb) ? p.kb_per_chunk : num_kb;
  const int num_chunks = (num_kb + kb_per_chunk - 1) / kb_per_chunk;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_bytes = 2u * (2u * p.a_box_bytes + 2u * kBHalfBytes);  // both CTAs' A (hi, lo) and B halves
    for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = 2 * (tile / p.n_tiles) + rank;  // may be one past the end: TMA zero-fills / clips it
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int tn = m_tile / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.box_w, h0 = th * p.box_h, n0 = tn * p.box_n;
      int kcoord = 0;
      for (int tap = 0; tap < p.num_taps; ++tap) {
        const int plane = p.tap_plane[tap];
        const int cw = w0 + p.tap_dw[tap];
        const int ch = h0 + p.tap_dh[tap];
        const int tap_blocks = p.tap_cb[tap] > 0 ? p.tap_cb[tap] : cin_blocks;
        for (int cb = 0; cb < tap_blocks; ++cb, kcoord += kBK) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * kStageBytes;
          if (rank == 0) mbar_arrive_expect_tx_elect(&full_bar[stage], tx_bytes);
          tma_load_4d_pair_elect(st, &p.tmap_a[0][plane], &full_bar[stage], cb * kBK, cw, ch, n0);
          tma_load_4d_pair_elect(st + kABytes, &p.tmap_a[1][plane], &full_bar[stage], cb * kBK, cw, ch, n0);
          uint8_t* sb = st + 2 * kABytes;
          const int brow = n_tile * kBlockN + rank * (kBlockN / 2);
          tma_load_2d_pair_elect(sb, &p.tmap_b_half[0], &full_bar[stage], kcoord, brow);
          tma_load_2d_pair_elect(sb + kBHalfBytes, &p.tmap_b_half[1], &full_bar[stage], kcoord, brow);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = make_idesc_16bit(2 * kGemmBlockM, kBlockN, 1u);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t cc = 0;
      for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
        int kb = 0;
        for (int chunk = 0; chunk < num_chunks; ++chunk, ++cc) {
          const int as = cc % kTmemBufs;
          const uint32_t aphase = (cc / kTmemBufs) & 1;
          mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + as * kBlockN;
          const int kb_end = (kb + kb_per_chunk < num_kb) ? kb + kb_per_chunk : num_kb;
          const int kb_first = kb;
          for (; kb < kb_end; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            const uint32_t a_hi = smem_u32(smem + stage * kStageBytes);
            const uint32_t b_hi = a_hi + 2 * kABytes;
            const uint64_t da_hi = make_smem_desc_sw128(a_hi);
            const uint64_t da_lo = make_smem_desc_sw128(a_hi + kABytes);
            const uint64_t db_hi = make_smem_desc_sw128(b_hi);
            const uint64_t db_lo = make_smem_desc_sw128(b_hi + kBHalfBytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t koff = 2 * k;
              umma_bf16_pair_elect(tmem_d, da_hi + koff, db_hi + koff, idesc, (kb > kb_first || k > 0) ? 1u : 0u);
              umma_bf16_pair_elect(tmem_d, da_lo + koff, db_hi + koff, idesc, 1u);
              umma_bf16_pair_elect(tmem_d, da_hi + koff, db_lo + koff, idesc, 1u);
            }
            umma_commit_pair_elect(&empty_bar[stage]);  // frees this stage in both CTAs
            if (kb == kb_end - 1) umma_commit_pair_elect(&tmem_full_bar[as]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ------------------------------------------------------------ io warp (RES): residual in, result out
    if (RES) {  // all lanes, uniform
      ChunkCursor ld(pair_id, total_tiles, num_pairs, rank), st(pair_id, total_tiles, num_pairs, rank);
      const uint32_t tx_bytes = 2u * p.a_box_bytes;  // hi + lo planes of one 64-column box
      auto load_residual = [&](const ChunkCursor& cur, int b) {
        int col0, w0, h0, n0;
        cur.coords(p, &col0, &w0, &h0, &n0);
        uint8_t* dst = staging + b * kStagingBytes;
        mbar_arrive_expect_tx_elect(&res_full_bar[b], tx_bytes);
        tma_load_4d_elect(dst, &p.tmap_res[0], &res_full_bar[b], col0, w0, h0, n0);
        tma_load_4d_elect(dst + kTileBytes, &p.tmap_res[1], &res_full_bar[b], col0, w0, h0, n0);
      };
      for (int b = 0; b < IO && ld.valid(); ++b, ld.next()) load_residual(ld, b);
      for (uint32_t n = 0; st.valid(); ++n, st.next()) {
        const int b = n % IO;
        mbar_wait(&out_ready_bar[b], (n / IO) & 1);  // the writers fenced their stores for the async proxy
        int col0, w0, h0, n0;
        st.coords(p, &col0, &w0, &h0, &n0);
        tma_store_4d_elect(&p.tmap_out[0], staging + b * kStagingBytes, col0, w0, h0, n0);
        tma_store_4d_elect(&p.tmap_out[1], staging + b * kStagingBytes + kTileBytes, col0, w0, h0, n0);
        tma_store_commit_elect();
        if (ld.valid()) {  // refill this buffer as soon as the store has read it; the epilogue is on another one
          tma_store_wait_read_elect<0>();
          load_residual(ld, b);
          ld.next();
        }
      }
      tma_store_wait_all_elect<0>();
    }
    __syncwarp();
  } else if (warp >= 2 && warp < 10) {
    // ------------------------------------------------------------ epilogue (both CTAs, own M tile)
    const int quarter = warp & 3;
    const int group = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const bool leader = warp == 2;
    uint32_t cc = 0;
    uint32_t io_count = 0;  // RES: chunks this CTA has produced (io buffer = count % IO, barrier phase = count / IO)
    for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = 2 * (tile / p.n_tiles) + rank;
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int tn = m_tile / (p.tiles_w * p.tiles_h);
      float v[2][32];
      for (int chunk = 0; chunk < num_chunks; ++chunk, ++cc) {
        const int as = cc % kTmemBufs;
        const uint32_t aphase = (cc / kTmemBufs) & 1;
        mbar_wait(&tmem_full_bar[as], aphase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * kBlockN + group * 32;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
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
        }
        tcgen05_fence_before();
        mbar_arrive_leader(&tmem_empty_bar[as]);  // the leader's MMA warp waits for both CTAs' drains
      }
      if constexpr (RES) {
#pragma unroll
        for (int c = 0; c < 2; ++c, ++io_count) {
          const int col0 = n_tile * kBlockN + c * 64;  // cout is a multiple of 128 here (checked at launch)
          const int b = io_count % IO;
          uint8_t* io = staging + b * kStagingBytes;
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0 + group * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = __ldg(b4 + j);
              v[c][4 * j + 0] += bb.x; v[c][4 * j + 1] += bb.y; v[c][4 * j + 2] += bb.z; v[c][4 * j + 3] += bb.w;
            }
          }
          mbar_wait(&res_full_bar[b], (io_count / IO) & 1);
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 r = *reinterpret_cast<const uint4*>(io + pl * kTileBytes + swz(row, group * 4 + j));
              v[c][8 * j + 0] += bf16_lo_to_f32(r.x); v[c][8 * j + 1] += bf16_hi_to_f32(r.x);
              v[c][8 * j + 2] += bf16_lo_to_f32(r.y); v[c][8 * j + 3] += bf16_hi_to_f32(r.y);
              v[c][8 * j + 4] += bf16_lo_to_f32(r.z); v[c][8 * j + 5] += bf16_hi_to_f32(r.z);
              v[c][8 * j + 6] += bf16_lo_to_f32(r.w); v[c][8 * j + 7] += bf16_hi_to_f32(r.w);
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
          for (int j = 0; j < 4; ++j) {  // the very bytes this thread just read: in place
            *reinterpret_cast<uint4*>(io + swz(row, group * 4 + j)) =
                make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            *reinterpret_cast<uint4*>(io + kTileBytes + swz(row, group * 4 + j)) =
                make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the TMA store the io warp issues
          mbar_arrive(&out_ready_bar[b]);
        }
      } else {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col0 = n_tile * kBlockN + c * 64;
        const bool active = col0 < p.cout;
        if (leader) tma_store_wait_read_elect<0>();
        named_bar_sync(1, kEpiThreads);
        if (active) {
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0 + group * 32);
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
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_bf16x2(v[c][2 * j], v[c][2 * j + 1], hi[j], lo[j]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            *reinterpret_cast<uint4*>(staging + swz(row, group * 4 + j)) =
                make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            *reinterpret_cast<uint4*>(staging + kTileBytes + swz(row, group * 4 + j)) =
                make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
        }
        fence_proxy_async();
        named_bar_sync(1, kEpiThreads);
        if (leader && active) {  // a tile past the end of M is clipped away entirely by the tensor map
          tma_store_4d_elect(&p.tmap_out[0], staging, col0, tw * p.box_w, th * p.box_h, tn * p.box_n);
          tma_store_4d_elect(&p.tmap_out[1], staging + kTileBytes, col0, tw * p.box_w, th * p.box_h, tn * p.box_n);
          tma_store_commit_elect();
        }
      }
      }
    }
    if (!RES && leader) tma_store_wait_all_elect<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA frees TMEM or exits while the pair's MMAs, commits or remote arrives are in flight
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace

namespace {
template <bool RES, int STAGES, int IO>
int launch_pair_impl(const ConvGemmParams& p, int num_sms, cudaStream_t stream, const int* skip_flag) {
  using L = PairLayout<RES, STAGES, IO>;
  static bool configured[64] = {};  // cudaFuncSetAttribute is per device
  static std::mutex mu;
  {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (dev < 0 || dev >= 64) return static_cast<int>(cudaErrorInvalidDevice);
    std::lock_guard<std::mutex> lock(mu);
    if (!configured[dev]) {
      e = cudaFuncSetAttribute(conv_gemm_pair_kernel<RES, STAGES, IO>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmemBytes);
      if (e != cudaSuccess) return static_cast<int>(e);
      configured[dev] = true;
    }
  }
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_pairs = ((m_tiles + 1) / 2) * p.n_tiles;
  if (total_pairs <= 0) return 0;
  int grid = 2 * total_pairs;
  if (grid > (num_sms & ~1)) grid = num_sms & ~1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = L::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_gemm_pair_kernel<RES, STAGES, IO>, p, skip_flag);
  count_conv_launch();
  return static_cast<int>(e);
}
}  // namespace

int launch_conv_gemm_pair(const ConvGemmParams& p, int num_sms, cudaStream_t stream, const int* skip_flag) {
  if (p.has_res) {
    if (p.cout % kBlockN != 0) return static_cast<int>(cudaErrorInvalidValue);  // whole 64-column chunks only
    return launch_pair_impl<true, 3, 2>(p, num_sms, stream, skip_flag);
  }
  return launch_pair_impl<false, 4, 1>(p, num_sms, stream, skip_flag);
}

}  // namespace milan
