// Persistent, warp-specialised tcgen05 GEMM for sm_100a with pluggable fused epilogues.
//
//   C[M,N] = A[M,K] * B[K,N]      bf16 operands, fp32 accumulation in TMEM
//
// One CTA per SM (grid = min(#tiles, #SMs)), 384 threads:
//   warp 0      TMA producer   -- cp.async.bulk.tensor tiles of A and B into a ring of
//                                 128B-swizzled shared-memory stages (mbarrier full/empty)
//   warp 1      MMA issuer     -- one elected lane issues tcgen05.mma (M=128, N=BN, K=16);
//                                 tcgen05.commit releases smem stages and publishes the tile
//   warps 2..3  idle (complete warpgroup 0 so that it can release registers with setmaxnreg)
//   warps 4..11 epilogue       -- tcgen05.ld the fp32 accumulator (thread == tile row),
//                                 apply the fused epilogue functor, write global memory.
//                                 Two warps per TMEM lane quarter alternate 32-column chunks; each chunk is
//                                 transposed through shared memory so the functor sees row-contiguous
//                                 float4 fragments and all its global traffic is fully coalesced.
// Optional split-K (weight gradients: M,N small, K = batch): work items are (tile, k-slice) pairs
// and the epilogue accumulates with red.global.add into a zero-initialised output.
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i
// overlaps the MMAs of tile i+1.
//
// Operand layouts (both supported for A and for B, selected at compile time):
//   K-major : matrix stored with K contiguous  (A as [M,K], B as [N,K])
//   MN-major: matrix stored with M/N contiguous (A as [K,M], B as [K,N])
// so forward (x @ W, W stored [in,out] = B MN-major), input-gradient (g @ W^T = B K-major)
// and weight-gradient (act^T @ g = A and B MN-major) GEMMs all read the Flax layouts
// directly with no transposed copies.
//
// Out-of-bounds rows / K tails are zero-filled by TMA; the epilogue is only invoked for
// rows < M and 4-column fragments that start below N (N must be a multiple of 4).
#pragma once

#include <cmath>
#include <type_traits>

#include "mfac_common.cuh"

namespace mfac {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_EPI_WARPS = 8;  // two per TMEM lane quarter, alternating 32-column chunks
// Warpgroup 0 = {TMA producer, MMA issuer, 2 idle warps}; warpgroups 1..2 = epilogue.  After setup warpgroup 0 gives
// most of its registers to the epilogue warps (setmaxnreg), which keep two chunks of functor operands in flight.
constexpr int GEMM_THREADS = 128 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_REGS_LAUNCH = 168;  // 65536 / 384 rounded down to a multiple of 8
constexpr int GEMM_REGS_CTRL = 56;
constexpr int GEMM_REGS_EPI = 224;     // 128 * 40 + 256 * 232 == 384 * 168

constexpr int GEMM_SMEM_TOTAL = 232448 - 1024 /*align slack*/ - 256 /*barriers*/;  // 227 KB per CTA

template <int BN, bool TMA_EPI = false>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // per-warp epilogue staging: one 32x32 fp32 transpose buffer, or two 32x64 bf16 tiles for the TMA-store epilogue
  static constexpr int STAGING_BYTES = GEMM_EPI_WARPS * (TMA_EPI ? 8192 : 4096);
  static constexpr int STAGES_MAX = (GEMM_SMEM_TOTAL - STAGING_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_MAX > 6 ? 6 : STAGES_MAX;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");
};

// Epilogue contract.  The accumulator arrives thread-per-row (tcgen05.ld 32x32b: lane == tile row); a direct
// global access in that layout touches 16-byte pieces of 32 different rows per instruction (partial sectors,
// 2-3 TB/s at best).  Each epilogue warp therefore transposes its 32x32 fp32 chunk through a private, XOR-swizzled
// 4 KB shared-memory buffer and hands the functor row-contiguous fragments:
//     epi.prefetch(m0, n0, bn, M, N, etid)      L2 prefetch of the tile's global operands, one tile ahead
//     epi.load_col(col, cregs)                  per-column operands (bias), once per chunk
//     epi.load(row, col, regs)                  global reads of the fragment (issued one chunk ahead of use)
//     epi.frag(row, col, float4 acc, regs, cregs)   row < M, col % 4 == 0, 8 consecutive lanes cover 32 columns of a row
// so every global load/store the functor issues is a full 32-byte sector (bf16) or 128-byte line (fp32).
struct GemmShape {
  int M, N, K;
  int splits;        // split-K factor (1 = none); work item = tile * splits + split
  int kb_per_split;  // 64-wide k-blocks per split
  int no_prefetch;   // L2 prefetch of the epilogue operands: 0 = first tile + one tile ahead, 1 = first tile only, 2 = none (default)
  int reverse_m;     // 1 = walk the row tiles from the last to the first (the rows the previous kernel touched last are
                     // the ones still in L2)
};
// Sweep direction of the next row-parallel kernel (GEMMs without split-K and the vectorised row kernels): consecutive
// kernels of a chain alternate between ascending and descending rows, so that each one starts on the rows its producer
// touched last -- the part of the producer's output that is still in the 126 MB L2 (the step's tensors are 40-270 MB each).
// Results do not depend on the direction.  MFAC_NO_SWEEP_ALTERNATE=1 keeps every kernel ascending.
// The direction state is per calling thread and is reset by every C-ABI entry point (sweep_reset), so the traversal order
// of a call's kernels depends on that call alone, never on what the thread launched before.
inline int& sweep_state() {
  static thread_local int cur = 0;
  return cur;
}
inline void sweep_reset() { sweep_state() = 0; }
inline int sweep_next() {
  static const int on = getenv("MFAC_NO_SWEEP_ALTERNATE") ? 0 : 1;
  if (!on) return 0;
  return sweep_state() ^= 1;
}
// The L2 prefetch of the next tile's epilogue operands paid while the epilogues were latency-bound; with the current
// register prefetch it only adds DRAM traffic (lines fetched early are evicted before use: 371 MB read against 294 MB of
// unique bytes in the tangent block-output GEMM) -- measured at 18944 rows: block_out 58 -> 51 us, block_out_tangent
// 98 -> 83 us, step 9.02 -> 8.81 ms without it.  MFAC_EPI_PREFETCH=2 switches it back on.
// (first-tile-only prefetch: 8.88 ms, none: 8.81 ms; 1024 and 4096 rows are indifferent.)
inline int epi_prefetch_off() {
  static const int off = getenv("MFAC_EPI_PREFETCH") ? 2 - atoi(getenv("MFAC_EPI_PREFETCH")) : 2;   // env 2 -> all, 1 -> first tile, 0 -> none
  return off < 0 ? 0 : (off > 2 ? 2 : off);
}

// Output tensor maps of the TMA-store epilogues (bf16 outputs, 32 x 32 boxes, 64-byte swizzle)
struct EpiTmaps {
  CUtensorMap m[2];
};
int make_tmap_bf16_sw(CUtensorMap* out, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                      int box_outer, int swizzle_bytes);

// TMA-store epilogue for functors with ONE plain bf16 output and a per-column vector operand (Epi::kTmaStore): the
// accumulator stays in its native layout (lane == row), the functor turns 2 x 32 columns into packed bf16, eight
// 16-byte shared-memory stores lay a 32 x 64 tile down in the 128-byte-swizzled layout the tensor map expects, and ONE
// elected lane hands it to the TMA engine (full 128-byte lines; two staging tiles per warp alternate).  Compared with
// the transposed path: 16 instead of 128 shared-memory wavefronts per 64 columns and no global store instructions at
// all -- the LSU/L1 path was the bottleneck of the store-bound K = 128 GEMMs.  Out-of-range rows are clipped by the
// tensor map; N must be a multiple of 64.
template <int BN, class Epi>
__device__ __forceinline__ void epilogue_tile_tma(const Epi& epi, const GemmShape& shape, const EpiTmaps& maps, uint8_t* stage,
                                                  uint32_t taddr, int row0, int n0, int half, int lane, uint64_t* tfull,
                                                  uint32_t aph, uint32_t& parity) {
  constexpr int NCH = BN / 128;                    // 64-column chunks handled by this warp: c = half + 2 j
  mbar_wait(tfull, aph);
  tc_fence_after();
  const int sw = lane & 7;                         // 128-byte swizzle: 16-byte piece p of row r sits at p ^ (r & 7)
#pragma unroll
  for (int j = 0; j < (NCH > 0 ? NCH : 1); ++j) {
    const int c = half + 2 * j;
    const int col0 = n0 + c * 64;
    if (c * 64 >= BN || col0 >= shape.N) break;    // warp-uniform
    float acc[64];
    tmem_ld_32x32(taddr + c * 64, *reinterpret_cast<float(*)[32]>(&acc[0]));
    tmem_ld_32x32(taddr + c * 64 + 32, *reinterpret_cast<float(*)[32]>(&acc[32]));
    tmem_ld_wait();
    uint32_t o[32];
    epi.compute(col0, *reinterpret_cast<const float(*)[32]>(&acc[0]), *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
    epi.compute(col0 + 32, *reinterpret_cast<const float(*)[32]>(&acc[32]), *reinterpret_cast<uint32_t(*)[16]>(&o[16]));
    if (lane == 0) tma_store_wait_read<1>();       // the tile written two chunks ago has been read by the engine
    __syncwarp();
    uint8_t* buf = stage + (parity & 1) * 4096;
    uint4* rowp = reinterpret_cast<uint4*>(buf + lane * 128);
#pragma unroll
    for (int p = 0; p < 8; ++p) rowp[p ^ sw] = make_uint4(o[4 * p], o[4 * p + 1], o[4 * p + 2], o[4 * p + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&maps.m[0], buf, col0, row0);
      tma_store_commit();
    }
    ++parity;
  }
}

// Functors with Epi::kColSum also want the column sums of what they store (bias gradients).
template <class E, class = void>
struct epi_colsum : std::false_type {};
template <class E>
struct epi_colsum<E, std::void_t<decltype(E::kColSum)>> : std::bool_constant<E::kColSum> {};

// Functors with Epi::kNarrowTiles also get a 64-column tile instantiation: GEMMs of a few hundred rows are bound by what ONE
// SM can pull from L2 (a 128 x 128 x K tile reads 2 x 128 x K operand elements), so the chain GEMMs of the small-batch step
// spread over twice as many SMs with half the B operand each.
template <class E, class = void>
struct epi_narrow : std::false_type {};
template <class E>
struct epi_narrow<E, std::void_t<decltype(E::kNarrowTiles)>> : std::bool_constant<E::kNarrowTiles> {};

// One epilogue warp's share of one output tile: 32 rows (its TMEM lane quarter) x every other 32-column chunk.
// FULL = the tile lies entirely inside the matrix (no row / column predicates in the hot loop).
// The functor's global operands are fetched (coalesced) TWO chunks ahead into two register sets: the first two
// chunks' while this tile's MMAs are still running, chunk j+2's as soon as chunk j has been written out.
template <int BN, bool FULL, class Epi>
__device__ __forceinline__ void epilogue_tile(const Epi& epi, const GemmShape& shape, float4* st, uint32_t taddr, int row0,
                                              int n0, int half, int lane, uint64_t* tfull, uint32_t aph) {
  constexpr int NCH = BN / 64;              // chunks handled by this warp: c = half + 2 j
  const int fr = lane >> 3, fp = lane & 7;  // fragment i of this lane: row row0 + 4 i + fr, columns col0 + 4 fp ..
  if constexpr (Epi::kPrefetchDepth == 0) {
    // store-only functors with bulky per-fragment code (gradient scatter): rolled loop, small instruction footprint
    typename Epi::Regs nr;
    typename Epi::ColRegs nc;
    mbar_wait(tfull, aph);
    tc_fence_after();
#pragma unroll 1
    for (int c = half; c < BN / 32; c += 2) {
      const int col0 = n0 + c * 32;
      if (col0 >= shape.N) break;  // warp-uniform
      float acc[32];
      tmem_ld_32x32(taddr + c * 32, acc);
      tmem_ld_wait();
      __syncwarp();
#pragma unroll
      for (int p = 0; p < 8; ++p)
        st[lane * 8 + (p ^ (lane & 7))] = make_float4(acc[4 * p], acc[4 * p + 1], acc[4 * p + 2], acc[4 * p + 3]);
      __syncwarp();
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + fr;
        const float4 u = st[r * 8 + (fp ^ (r & 7))];
        if (row0 + r < shape.M && col0 + 4 * fp < shape.N) epi.frag(row0 + r, col0 + 4 * fp, u, nr, nc);
      }
    }
    return;
  } else {
  constexpr int DEPTH = Epi::kPrefetchDepth < NCH ? (Epi::kPrefetchDepth > 0 ? Epi::kPrefetchDepth : 1) : NCH;  // register sets in flight
  typename Epi::Regs regs[DEPTH][8];
  typename Epi::ColRegs cregs[DEPTH];
#pragma unroll
  for (int j = 0; j < DEPTH; ++j) {
    const int col0 = n0 + (half + 2 * j) * 32;
    if (FULL || col0 + 4 * fp < shape.N) {   // N may be any multiple of 4: the last chunk can be partial
      epi.load_col(col0 + 4 * fp, cregs[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (FULL || row0 + 4 * i + fr < shape.M) epi.load(row0 + 4 * i + fr, col0 + 4 * fp, regs[j][i]);
    }
  }
  mbar_wait(tfull, aph);
  tc_fence_after();
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c = half + 2 * j;
    const int col0 = n0 + c * 32;
    if (!FULL && col0 >= shape.N) break;  // warp-uniform
    float acc[32];
    tmem_ld_32x32(taddr + c * 32, acc);
    tmem_ld_wait();
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 8; ++p)
      st[lane * 8 + (p ^ (lane & 7))] = make_float4(acc[4 * p], acc[4 * p + 1], acc[4 * p + 2], acc[4 * p + 3]);
    __syncwarp();
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + fr;
      const float4 u = st[r * 8 + (fp ^ (r & 7))];
      if (FULL || (row0 + r < shape.M && col0 + 4 * fp < shape.N)) {
        if constexpr (epi_colsum<Epi>::value) {
          const float4 v = epi.frag(row0 + r, col0 + 4 * fp, u, regs[j % DEPTH][i], cregs[j % DEPTH]);
          csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
        } else {
          epi.frag(row0 + r, col0 + 4 * fp, u, regs[j % DEPTH][i], cregs[j % DEPTH]);
        }
      }
    }
    if constexpr (epi_colsum<Epi>::value) {
      // lanes l, l^8, l^16, l^24 hold the same 4 columns of different rows: fold them, one atomic float4 per 32 rows
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o);
        csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
        csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o);
        csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
      }
      if (fr == 0 && (FULL || col0 + 4 * fp < shape.N)) epi.col_add(col0 + 4 * fp, csum);
    }
    if (j + DEPTH < NCH) {
      const int coln = col0 + 64 * DEPTH;
      if (FULL || coln + 4 * fp < shape.N) {
        epi.load_col(coln + 4 * fp, cregs[j % DEPTH]);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (FULL || row0 + 4 * i + fr < shape.M) epi.load(row0 + 4 * i + fr, coln + 4 * fp, regs[j % DEPTH][i]);
      }
    }
  }
  }
}

template <int BN, bool A_MN, bool B_MN, bool FULL, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ EpiTmaps tmOut, GemmShape shape, Epi epi) {
  using Cfg = GemmCfg<BN, Epi::kTmaStore>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sStage = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = ceil_div(shape.M, GEMM_BM);
  const int n_tiles = ceil_div(shape.N, BN);
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks_total = ceil_div(shape.K, GEMM_BK);
  const int num_items = num_tiles * shape.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], GEMM_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (descriptor prefetch, barrier init, TMEM allocation) touched no global
  // memory and may run under the tail of the previous kernel in the stream; let our own successor start its set-up, then
  // wait until the predecessor has completed and its writes are visible.  (Both are no-ops for a plain launch.)
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_CTRL));
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t kit = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int tile = item / shape.splits, sp = item % shape.splits;
        const int m0 = (shape.reverse_m ? m_tiles - 1 - tile / n_tiles : tile / n_tiles) * GEMM_BM;
        const int n0 = (tile % n_tiles) * BN;
        const int kb0 = sp * shape.kb_per_split;
        const int kb1 = min(k_blocks_total, kb0 + shape.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          uint8_t* a = sA + s * Cfg::A_BYTES;
          uint8_t* b = sB + s * Cfg::B_BYTES;
          const int k0 = kb * GEMM_BK;
          if constexpr (A_MN) {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d(a + j * 8192, &tmA, &full_bar[s], m0 + 64 * j, k0);
          } else {
            tma_load_2d(a, &tmA, &full_bar[s], k0, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, k0);
          } else {
            tma_load_2d(b, &tmB, &full_bar[s], k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_CTRL));
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, A_MN, B_MN);
      // Shared-memory descriptors of stage 0, built once: the single issuing thread is on the critical path of narrow tiles
      // (a 64-column k-block is 128 clocks of tensor work), so the k loop only adds to the start-address field.
      // K-major: 16 elements = 32 B inside the 128-B swizzle row; MN-major: 16 k-rows of 128 B = 2048 B (SBO apart).
      const uint64_t a_desc0 = A_MN ? umma_smem_desc_sw128(smem_u32(sA), 8192, 1024) : umma_smem_desc_sw128(smem_u32(sA), 16, 1024);
      const uint64_t b_desc0 = B_MN ? umma_smem_desc_sw128(smem_u32(sB), 8192, 1024) : umma_smem_desc_sw128(smem_u32(sB), 16, 1024);
      constexpr int a_kk = (A_MN ? 2048 : 32) >> 4, b_kk = (B_MN ? 2048 : 32) >> 4;
      uint32_t kit = 0, it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int sp = item % shape.splits;
        const int kb0 = sp * shape.kb_per_split;
        const int kb1 = min(k_blocks_total, kb0 + shape.kb_per_split);
        const uint32_t as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          // stage s, k-step kk: only the 14-bit start-address field (bytes >> 4) of the stage-0 descriptors moves
          const uint64_t ad = a_desc0 + (uint64_t)(s * (int)(Cfg::A_BYTES >> 4));
          const uint64_t bd = b_desc0 + (uint64_t)(s * (int)(Cfg::B_BYTES >> 4));
#pragma unroll
          for (int kk = 0; kk < GEMM_BK / 16; ++kk)
            umma_bf16(tmem_d, ad + (uint64_t)(kk * a_kk), bd + (uint64_t)(kk * b_kk), idesc, (kb > kb0 || kk != 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);  // smem stage reusable once these MMAs retire
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
      }
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_CTRL));
  } else {
    // ===================== epilogue (warps 4..11) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_EPI));
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;        // which of the quarter's two warps: even / odd chunks
    const int etid = threadIdx.x - 128;      // 0 .. 32 * GEMM_EPI_WARPS - 1
    float4* st = reinterpret_cast<float4*>(sStage + (warp - 4) * (Cfg::STAGING_BYTES / GEMM_EPI_WARPS));
    uint32_t it = 0, tma_parity = 0;
    // The functor's global operands of the first tile are pulled into L2 while its MMAs run; every later
    // tile's are requested one tile ahead (the register prefetch inside a tile then only sees L2 latency).
    if ((int)blockIdx.x < num_items) {
      const int tile = blockIdx.x / shape.splits;
      if (shape.no_prefetch < 2) epi.prefetch((tile / n_tiles) * GEMM_BM, (tile % n_tiles) * BN, BN, shape.M, shape.N, etid);
    }
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int tile = item / shape.splits;
      const uint32_t as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int m0 = (shape.reverse_m ? m_tiles - 1 - tile / n_tiles : tile / n_tiles) * GEMM_BM;
      const int n0 = (tile % n_tiles) * BN;
      if (item + (int)gridDim.x < num_items) {
        const int nt = (item + gridDim.x) / shape.splits;
        if (nt != tile && !shape.no_prefetch) epi.prefetch((nt / n_tiles) * GEMM_BM, (nt % n_tiles) * BN, BN, shape.M, shape.N, etid);
      }
      const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(quarter * 32) << 16);
      if constexpr (Epi::kTmaStore)
        epilogue_tile_tma<BN>(epi, shape, tmOut, reinterpret_cast<uint8_t*>(st), taddr, m0 + quarter * 32, n0, half, lane,
                              &tfull_bar[as], aph, tma_parity);
      else
        epilogue_tile<BN, FULL>(epi, shape, st, taddr, m0 + quarter * 32, n0, half, lane, &tfull_bar[as], aph);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    if constexpr (Epi::kTmaStore) {
      if (lane == 0) tma_store_wait_all();   // the staging buffers must outlive the engine's reads
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2).  A 2-wide cluster owns a 256 x 256 output tile: each CTA stages its own 128 rows of A
// and HALF of the B tile (128 of the 256 columns), the leader's MMA thread issues tcgen05.mma.cta_group::2 (M = 256)
// which reads both CTAs' shared memory and writes each CTA's 128 x 256 accumulator into its own TMEM, and each CTA runs
// the usual epilogue on its half.  Per SM this halves the B-operand fill and read traffic (32 instead of 48 KB per
// k-block: six stages instead of four) -- the compute-bound GEMMs were limited by operand delivery, not by the tensor
// pipe.  Protocol (after CUTLASS's sm100 2-SM pipeline):
//   full[s]   lives in the leader; the leader's producer arms it with the pair's byte count, both CTAs' TMA loads
//             complete_tx on it (cp.async.bulk.tensor ... cta_group::2 with the leader's barrier address)
//   empty[s]  one per CTA; tcgen05.commit.cta_group::2 ... multicast arrives on both when the stage's MMAs retire
//   tfull[a]  one per CTA, same multicast commit: the accumulator of buffer a is complete in both TMEMs
//   tempty[a] lives in the leader and counts the epilogue warps of BOTH CTAs (the peer's arrive remotely)
// ---------------------------------------------------------------------------------------
// Work decomposition of one CTA pair.  Uniform mode (shape.splits >= 1): items (tile, k-slice) strided over the pairs.
// Stream-K mode (shape.splits == 0, weight gradients): the tiles' k-blocks form one line of tiles * k_blocks entries that is
// cut into equal contiguous ranges, one per pair; a range crosses at most one tile boundary per k_blocks entries, so every
// pair does the same number of MMAs (no wave quantisation) and the number of partial tiles that go through the atomic
// epilogue drops from tiles * splits to about pairs + tiles.
struct Seg { int tile, kb0, kb1; };
struct SegIter {
  int item, stride, num_items, splits, kbps, kbt, tiles;   // uniform
  int g, g_end;                                            // stream-K
  __device__ __forceinline__ SegIter(const GemmShape& sh, int tiles_, int kbt_, int pair, int npairs)
      : item(pair), stride(npairs), num_items(tiles_ * sh.splits), splits(sh.splits), kbps(sh.kb_per_split), kbt(kbt_),
        tiles(tiles_), g(0), g_end(0) {
    if (splits == 0) {
      const int total = tiles * kbt, q = (total + npairs - 1) / npairs;
      g = pair * q < total ? pair * q : total;
      g_end = g + q < total ? g + q : total;
    }
  }
  __device__ __forceinline__ bool next(Seg& o) {
    if (splits == 0) {
      if (g >= g_end) return false;
      o.tile = g / kbt;
      o.kb0 = g - o.tile * kbt;
      o.kb1 = o.kb0 + (g_end - g) < kbt ? o.kb0 + (g_end - g) : kbt;
      g += o.kb1 - o.kb0;
      return true;
    }
    if (item >= num_items) return false;
    // k-slice major: the pairs of one wave work on the SAME k-slice of all the tiles, so every operand panel of that
    // slice is fetched from HBM once and shared through L2 by the tiles of its row / column (tile-major order re-read the
    // B panels once per tile row: 494 MB of DRAM traffic for 200 MB of operands at 37888 rows)
    const int sp = item / tiles;
    o.tile = item - sp * tiles;
    o.kb0 = sp * kbps;
    o.kb1 = o.kb0 + kbps < kbt ? o.kb0 + kbps : kbt;
    item += stride;
    return true;
  }
  // tile of the segment after the current one, or -1 (uniform mode only: the look-ahead drives the L2 operand prefetch)
  __device__ __forceinline__ int peek_tile() const { return (splits != 0 && item < num_items) ? item % tiles : -1; }
};

constexpr int GEMM2_BN = 256;
struct Gemm2Cfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;          // this CTA's 128 rows
  static constexpr int B_BYTES = (GEMM2_BN / 2) * GEMM_BK * 2;   // this CTA's 128 of the 256 columns
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = GEMM_EPI_WARPS * 4096;
  static constexpr int STAGES = 6;
  static constexpr int TMEM_COLS = 2 * GEMM2_BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 256;
  static_assert(STAGES * STAGE_BYTES + STAGING_BYTES <= GEMM_SMEM_TOTAL, "shared memory budget");
};

template <bool A_MN, bool B_MN, bool FULL, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmShape shape,
                         Epi epi) {
  using Cfg = Gemm2Cfg;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BN = GEMM2_BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sStage = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();           // 0 = leader
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_tiles = ceil_div(shape.M, 2 * GEMM_BM);  // 256-row pair tiles
  const int n_tiles = ceil_div(shape.N, BN);
  const int k_blocks_total = ceil_div(shape.K, GEMM_BK);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * GEMM_EPI_WARPS);   // both CTAs' epilogue warps (only the leader's copy is used)
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();                                // barriers of both CTAs initialised before any remote access
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (descriptor prefetch, barrier init, TMEM allocation) touched no global
  // memory and may run under the tail of the previous kernel in the stream; let our own successor start its set-up, then
  // wait until the predecessor has completed and its writes are visible.  (Both are no-ops for a plain launch.)
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_CTRL));
    if (lane == 0) {
      uint32_t kit = 0;
      SegIter segs(shape, m_tiles * n_tiles, k_blocks_total, pair, npairs);
      Seg sg;
      while (segs.next(sg)) {
        const int tile = sg.tile;
        const int m0 = (shape.reverse_m ? m_tiles - 1 - tile / n_tiles : tile / n_tiles) * 2 * GEMM_BM + (int)rank * GEMM_BM;   // this CTA's rows
        const int n0 = (tile % n_tiles) * BN + (int)rank * (BN / 2);           // this CTA's half of the columns
        const int kb0 = sg.kb0, kb1 = sg.kb1;
        for (int kb = kb0; kb < kb1; ++kb, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t lbar = mapa_u32(&full_bar[s], 0);
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
          uint8_t* a = sA + s * Cfg::A_BYTES;
          uint8_t* b = sB + s * Cfg::B_BYTES;
          const int k0 = kb * GEMM_BK;
          if constexpr (A_MN) {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d_pair(a + j * 8192, &tmA, lbar, m0 + 64 * j, k0);
          } else {
            tma_load_2d_pair(a, &tmA, lbar, k0, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 2 / 64; ++j) tma_load_2d_pair(b + j * 8192, &tmB, lbar, n0 + 64 * j, k0);
          } else {
            tma_load_2d_pair(b, &tmB, lbar, k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_CTRL));
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, BN, A_MN, B_MN);
      // Shared-memory descriptors of stage 0, built once: the single issuing thread is on the critical path of narrow tiles
      // (a 64-column k-block is 128 clocks of tensor work), so the k loop only adds to the start-address field.
      // K-major: 16 elements = 32 B inside the 128-B swizzle row; MN-major: 16 k-rows of 128 B = 2048 B (SBO apart).
      const uint64_t a_desc0 = A_MN ? umma_smem_desc_sw128(smem_u32(sA), 8192, 1024) : umma_smem_desc_sw128(smem_u32(sA), 16, 1024);
      const uint64_t b_desc0 = B_MN ? umma_smem_desc_sw128(smem_u32(sB), 8192, 1024) : umma_smem_desc_sw128(smem_u32(sB), 16, 1024);
      constexpr int a_kk = (A_MN ? 2048 : 32) >> 4, b_kk = (B_MN ? 2048 : 32) >> 4;
      uint32_t kit = 0, it = 0;
      SegIter segs(shape, m_tiles * n_tiles, k_blocks_total, pair, npairs);
      Seg sg;
      for (; segs.next(sg); ++it) {
        const int kb0 = sg.kb0, kb1 = sg.kb1;
        const uint32_t as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          // stage s, k-step kk: only the 14-bit start-address field (bytes >> 4) of the stage-0 descriptors moves
          const uint64_t ad = a_desc0 + (uint64_t)(s * (int)(Cfg::A_BYTES >> 4));
          const uint64_t bd = b_desc0 + (uint64_t)(s * (int)(Cfg::B_BYTES >> 4));
#pragma unroll
          for (int kk = 0; kk < GEMM_BK / 16; ++kk)
            umma_bf16_pair(tmem_d, ad + (uint64_t)(kk * a_kk), bd + (uint64_t)(kk * b_kk), idesc, (kb > kb0 || kk != 0) ? 1u : 0u);
          umma_commit_pair(&empty_bar[s]);   // both CTAs' stage s reusable once these MMAs retire
        }
        umma_commit_pair(&tfull_bar[as]);    // accumulator complete in both TMEMs
      }
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_CTRL));
  } else {
    // ===================== epilogue (warps 4..11 of both CTAs) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_EPI));
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    const int etid = threadIdx.x - 128;
    float4* st = reinterpret_cast<float4*>(sStage + (warp - 4) * 4096);
    uint32_t it = 0;
    const uint32_t lead_tempty0 = mapa_u32(&tempty_bar[0], 0), lead_tempty1 = mapa_u32(&tempty_bar[1], 0);
    SegIter segs(shape, m_tiles * n_tiles, k_blocks_total, pair, npairs);
    Seg sg;
    {
      const int tile = segs.peek_tile();
      if (tile >= 0 && shape.no_prefetch < 2)
        epi.prefetch((tile / n_tiles) * 2 * GEMM_BM + (int)rank * GEMM_BM, (tile % n_tiles) * BN, BN, shape.M, shape.N, etid);
    }
    for (; segs.next(sg); ++it) {
      const int tile = sg.tile;
      const uint32_t as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int m0 = (shape.reverse_m ? m_tiles - 1 - tile / n_tiles : tile / n_tiles) * 2 * GEMM_BM + (int)rank * GEMM_BM;
      const int n0 = (tile % n_tiles) * BN;
      const int nt = segs.peek_tile();
      if (nt >= 0 && nt != tile && !shape.no_prefetch)
        epi.prefetch((nt / n_tiles) * 2 * GEMM_BM + (int)rank * GEMM_BM, (nt % n_tiles) * BN, BN, shape.M, shape.N, etid);
      const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(quarter * 32) << 16);
      epilogue_tile<BN, FULL>(epi, shape, st, taddr, m0 + quarter * 32, n0, half, lane, &tfull_bar[as], aph);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(as ? lead_tempty1 : lead_tempty0);
    }
  }

  // neither CTA may leave (or free TMEM) while its partner can still read its shared memory or signal its barriers
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

// Debug/triage kernel: same contract and epilogues, plain SIMT fp32 accumulation.
// Selected only through mfac_debug_set_simt_gemm(1); never part of a product path.
struct SimtOperand {
  const __nv_bfloat16* p;
  int64_t ld;
  int mn_major;
};
template <class Epi>
__global__ void gemm_simt_kernel(SimtOperand A, SimtOperand B, GemmShape shape, Epi epi) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int groups = shape.N / 32;
  const int row = (int)(gid / groups);
  const int col0 = (int)(gid % groups) * 32;
  if (row >= shape.M) return;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  for (int k = 0; k < shape.K; ++k) {
    const float a = __bfloat162float(A.mn_major ? A.p[(int64_t)k * A.ld + row] : A.p[(int64_t)row * A.ld + k]);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float b = __bfloat162float(B.mn_major ? B.p[(int64_t)k * B.ld + col0 + j] : B.p[(int64_t)(col0 + j) * B.ld + k]);
      acc[j] = fmaf(a, b, acc[j]);
    }
  }
  if constexpr (Epi::kTmaStore) {
    epi.store_row(row, col0, acc);   // debug path of the TMA-store functors: plain stores
  } else {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      typename Epi::Regs regs;
      typename Epi::ColRegs cregs;
      epi.load_col(col0 + j, cregs);
      epi.load(row, col0 + j, regs);
      epi.frag(row, col0 + j, make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]), regs, cregs);
    }
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct GemmOperandDesc {
  const void* ptr;   // bf16
  int64_t ld;        // leading dimension in elements (stride between stored rows)
  bool mn_major;     // false: stored [MN, K]; true: stored [K, MN]
};

// Encodes a 2-D bf16 tensor map (128B swizzle). inner/outer are the stored extents.
int make_tmap_bf16(CUtensorMap* out, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                   int box_outer);
bool simt_gemm_enabled();
void count_launch();
void* profile_begin(int family, double work, cudaStream_t s, const char* label = nullptr, int M = 0, int N = 0, int K = 0);
void profile_end(void* token, cudaStream_t s);

// GEMM launches carry the programmatic-stream-serialization attribute: a GEMM that follows another GEMM starts its
// set-up as soon as the predecessor's CTAs have passed theirs (MFAC_NO_PDL=1 launches plainly).
inline bool pdl_enabled() {
  static const bool on = getenv("MFAC_NO_PDL") == nullptr;
  return on;
}
template <class Kern, class... Args>
inline void launch_pdl(Kern kern, dim3 grid, dim3 threads, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = threads;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, args...);
}

template <int BN, bool A_MN, bool B_MN, class Epi>
int launch_gemm_bn(const GemmOperandDesc& A, const GemmOperandDesc& B, int M, int N, int K, const Epi& epi,
                   cudaStream_t stream, int splits) {
  using Cfg = GemmCfg<BN, Epi::kTmaStore>;
  CUtensorMap tmA, tmB;
  if (A_MN) {
    MFAC_OK(make_tmap_bf16(&tmA, A.ptr, M, K, A.ld, 64, 64));
  } else {
    MFAC_OK(make_tmap_bf16(&tmA, A.ptr, K, M, A.ld, 64, GEMM_BM));
  }
  if (B_MN) {
    MFAC_OK(make_tmap_bf16(&tmB, B.ptr, N, K, B.ld, 64, 64));
  } else {
    MFAC_OK(make_tmap_bf16(&tmB, B.ptr, K, N, B.ld, 64, BN));
  }
  // FULL: every tile lies inside the matrix, so the epilogue carries no row / column predicates (and half the code).
  const bool full = (M % GEMM_BM == 0) && (N % BN == 0);
  auto kern = full ? gemm_tcgen05_kernel<BN, A_MN, B_MN, true, Epi> : gemm_tcgen05_kernel<BN, A_MN, B_MN, false, Epi>;
  static PerDeviceOnce configured;  // one per template instantiation
  if (configured.need()) {
    MFAC_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, A_MN, B_MN, true, Epi>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    MFAC_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, A_MN, B_MN, false, Epi>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured.done();
  }
  const int tiles = ceil_div(M, GEMM_BM) * ceil_div(N, BN);
  const int k_blocks = ceil_div(K, GEMM_BK);
  if (splits > k_blocks) splits = k_blocks;
  if (splits < 1) splits = 1;
  const int kbps = ceil_div(k_blocks, splits);
  splits = ceil_div(k_blocks, kbps);  // no empty slice
  const int items = tiles * splits;
  const int grid = items < num_sms() ? items : num_sms();
  GemmShape shape{M, N, K, splits, kbps, epi_prefetch_off(), splits != 1 ? 0 : sweep_next()};
  EpiTmaps maps;
  if constexpr (Epi::kTmaStore) {
    MFAC_OK(epi.make_maps(maps, M, N));
  } else {
    maps.m[0] = tmA;   // unused by the kernel; any valid descriptor
    maps.m[1] = tmA;
  }
  void* prof = profile_begin(MFAC_PROF_GEMM, 2.0 * (double)M * (double)N * (double)K, stream, Epi::name, M, N, K);
  launch_pdl(kern, dim3((unsigned)grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, maps, shape, epi);
  profile_end(prof, stream);
  count_launch();
  return launch_status();
}

// ---- intra-call concurrency (runtime.cu) ------------------------------------------------------------------------------
// Small batches leave most of the machine idle: a 128-row GEMM occupies 10 of 148 SMs.  Entry points may therefore fork
// independent kernel chains onto the library's own side streams and join them back before they return, so the caller
// still sees ONE stream-ordered operation.  Streams and events are created once per (thread, device); forks are plain
// event record / wait pairs, which stream capture turns into graph edges (GraphedTrainStep replays them as a DAG).
struct ForkCtx {
  static constexpr int kStreams = 6;
  static constexpr int kEvents = 1024;
  cudaStream_t side[kStreams];
  cudaEvent_t ev[kEvents];
  int next = 0;
  cudaEvent_t next_event() { next = (next + 1) % kEvents; return ev[next]; }
};
ForkCtx* fork_ctx();          // nullptr when the streams / events cannot be created
int concurrency_max_rows();   // batches up to this many rows use the concurrent schedules (MFAC_CONC_MAX_ROWS, default 4096; 0 = off)
// everything enqueued on `from` so far happens before whatever is enqueued on `to` from now on
inline int stream_after(ForkCtx* fc, cudaStream_t from, cudaStream_t to) {
  cudaEvent_t e = fc->next_event();
  MFAC_CUDA_OK(cudaEventRecord(e, from));
  MFAC_CUDA_OK(cudaStreamWaitEvent(to, e, 0));
  return MFAC_SUCCESS;
}

void phase_mark(int id, cudaStream_t s);   // runtime.cu: debug event at a phase boundary (no-op unless mfac_debug_phase_marks(1))
bool pair_gemm_enabled();   // runtime.cu: on unless MFAC_NO_PAIR_GEMM is set / mfac_debug_set_pair_gemm(0)
bool stream_k_enabled();    // runtime.cu: on unless MFAC_NO_STREAM_K is set

// CTA pairs pay for compute-bound shapes that fill the machine with whole 256 x 256 tiles.
inline bool pair_gemm_pays(int M, int N, int K, bool split_k) {
  if (N % GEMM2_BN != 0 || M < 2 * GEMM_BM || K < 512) return false;
  const int items = ceil_div(M, 2 * GEMM_BM) * (N / GEMM2_BN);
  const int pairs = num_sms() / 2;
  if (split_k) return items * 4 >= pairs;            // the split search below fills the waves
  const double waves = (double)items / pairs;
  return items >= pairs && waves / std::ceil(waves) >= 0.8;
}

template <bool A_MN, bool B_MN, class Epi>
int launch_gemm_pair(const GemmOperandDesc& A, const GemmOperandDesc& B, int M, int N, int K, const Epi& epi, cudaStream_t stream,
                     bool split_k) {
  using Cfg = Gemm2Cfg;
  CUtensorMap tmA, tmB;
  if (A_MN) {
    MFAC_OK(make_tmap_bf16(&tmA, A.ptr, M, K, A.ld, 64, 64));
  } else {
    MFAC_OK(make_tmap_bf16(&tmA, A.ptr, K, M, A.ld, 64, GEMM_BM));
  }
  if (B_MN) {
    MFAC_OK(make_tmap_bf16(&tmB, B.ptr, N, K, B.ld, 64, 64));
  } else {
    MFAC_OK(make_tmap_bf16(&tmB, B.ptr, K, N, B.ld, 64, GEMM2_BN / 2));
  }
  const bool full = (M % (2 * GEMM_BM) == 0) && (N % GEMM2_BN == 0);
  auto kern = full ? gemm_tcgen05_pair_kernel<A_MN, B_MN, true, Epi> : gemm_tcgen05_pair_kernel<A_MN, B_MN, false, Epi>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    MFAC_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_pair_kernel<A_MN, B_MN, true, Epi>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    MFAC_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_pair_kernel<A_MN, B_MN, false, Epi>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured.done();
  }
  const int tiles = ceil_div(M, 2 * GEMM_BM) * ceil_div(N, GEMM2_BN);
  const int k_blocks = ceil_div(K, GEMM_BK);
  const int pairs = num_sms() / 2;
  int splits = 1;
  if (split_k) {
    double best = -1.0;
    for (int sp = 1; sp <= 32 && sp <= k_blocks; ++sp) {
      if (sp > 1 && k_blocks / sp < 4) break;
      const int items = tiles * ceil_div(k_blocks, ceil_div(k_blocks, sp));
      const double waves = (double)items / pairs;
      const double eff = waves / std::ceil(waves) - 0.004 * sp;
      if (eff > best) { best = eff; splits = sp; }
    }
  }
  int kbps = ceil_div(k_blocks, splits);
  splits = ceil_div(k_blocks, kbps);
  int items = tiles * splits;
  const double uniform_waves = (double)items / pairs;
  if (split_k && stream_k_enabled() && (int64_t)tiles * k_blocks >= 8 * (int64_t)pairs &&
      uniform_waves / std::ceil(uniform_waves) < 0.97 && kbps < 32) {
    // stream-K (see SegIter): every pair gets the same number of k-blocks.  Only where the best uniform slicing leaves a
    // ragged last wave: concurrent uniform items walk the same k-slice of neighbouring tiles and share those operand
    // panels in L2, stream-K ranges do not (measured at 18944 rows: 1280x1024 47 us uniform / 57 us stream-K,
    // 1280x1280 76 us uniform (4.73 waves) / 64 us stream-K).  Below ~8 k-blocks per pair the ranges are too short to pay
    // (1024 rows: 2.42 ms/step uniform, 2.59 stream-K; 2048 rows: 2.95 / 2.84); with long uniform slices (>= 32 k-blocks,
    // 37888 rows) the k-slice-major uniform order wins again (14.5 vs 14.7 ms/step): its waves share operand panels in L2.
    splits = 0;
    kbps = 0;
    items = pairs;
  }
  const int grid = 2 * (items < pairs ? items : pairs);
  GemmShape shape{M, N, K, splits, kbps, epi_prefetch_off(), splits != 1 ? 0 : sweep_next()};
  void* prof = profile_begin(MFAC_PROF_GEMM, 2.0 * (double)M * (double)N * (double)K, stream, Epi::name, M, N, K);
  launch_pdl(kern, dim3((unsigned)grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, shape, epi);
  profile_end(prof, stream);
  count_launch();
  return launch_status();
}

// C = A * B with the given epilogue.  N must be a multiple of 4; K and M are arbitrary
// (TMA zero-fills), leading dimensions must be multiples of 8 elements (16-byte TMA strides).
// split_k = true lets the launcher slice K so that (tiles x slices) covers the machine; the epilogue
// must then accumulate atomically (Epi::kAtomic) into a zero-initialised output.
template <bool A_MN, bool B_MN, class Epi>
int launch_gemm(const GemmOperandDesc& A, const GemmOperandDesc& B, int M, int N, int K, const Epi& epi,
                cudaStream_t stream, int force_bn = 0, bool split_k = false) {
  if (M <= 0 || N <= 0 || K <= 0) return MFAC_ERR_BAD_SHAPE;
  if (N % 4 != 0 || A.ld % 8 != 0 || B.ld % 8 != 0) return MFAC_ERR_UNSUPPORTED;
  if (A.mn_major != A_MN || B.mn_major != B_MN) return MFAC_ERR_UNSUPPORTED;
  if (simt_gemm_enabled()) {
    SimtOperand a{reinterpret_cast<const __nv_bfloat16*>(A.ptr), A.ld, A_MN ? 1 : 0};
    SimtOperand b{reinterpret_cast<const __nv_bfloat16*>(B.ptr), B.ld, B_MN ? 1 : 0};
    const int64_t threads = (int64_t)M * (N / 32);
    gemm_simt_kernel<Epi><<<(unsigned)ceil_div<int64_t>(threads, 128), 128, 0, stream>>>(a, b, GemmShape{M, N, K, 1, 0, 0, 0}, epi);
    count_launch();
    return launch_status();
  }
  // Tile choice: BN=256 halves the per-MMA shared-memory traffic (96 vs 128 B/clk) but only
  // pays when it does not cost a wave; prefer it when N divides and the tile count still
  // covers the machine.
  int bn = force_bn;
  if (bn == 0) {
    // rounds of the persistent grid x relative cost of one tile (a 128x256 tile does twice the work of a 128x128 one
    // at ~13 % better tensor efficiency): e.g. 160 tiles of 256 on 148 SMs are two rounds, 320 tiles of 128 only three
    // half-size ones.
    const int sms = num_sms();
    const int t256 = ceil_div(M, GEMM_BM) * ceil_div(N, 256), t128 = ceil_div(M, GEMM_BM) * ceil_div(N, 128);
    const double c256 = (double)ceil_div(t256, sms) * 1.74, c128 = (double)ceil_div(t128, sms);
    bn = (N % 256 == 0 && c256 <= c128) ? 256 : 128;
  }
  int splits = 1;
  if (split_k) {
    // Weight gradients: few output tiles, K = batch.  Pick the slice count whose (tiles x slices) work
    // items fill whole waves of the machine best, with at least 4 k-blocks per item.
    bn = force_bn ? force_bn : 128;   // (256-wide tiles measured slower for the [C, 2I+D] modulation gradient: 9.16 vs 9.02 ms/step)
    const int tiles = ceil_div(M, GEMM_BM) * ceil_div(N, bn);
    const int k_blocks = ceil_div(K, GEMM_BK);
    double best = -1.0;
    for (int sp = 1; sp <= 32 && sp <= k_blocks; ++sp) {
      if (sp > 1 && k_blocks / sp < 4) break;
      const int items = tiles * ceil_div(k_blocks, ceil_div(k_blocks, sp));
      const double waves = (double)items / num_sms();
      const double eff = waves / std::ceil(waves) - 0.004 * sp;  // mild preference for fewer atomics
      if (eff > best) { best = eff; splits = sp; }
    }
  }
  if constexpr (!Epi::kTmaStore) {
    if (!force_bn && pair_gemm_enabled() && pair_gemm_pays(M, N, K, split_k)) return launch_gemm_pair<A_MN, B_MN, Epi>(A, B, M, N, K, epi, stream, split_k);
  }
  if constexpr (epi_narrow<Epi>::value && !Epi::kTmaStore) {
    // small M: the 128-wide tiling leaves most SMs idle and each busy one L2-bandwidth bound -> 64-wide tiles
    if (!force_bn && !split_k && N % 64 == 0 && ceil_div(M, GEMM_BM) * ceil_div(N, 128) * 4 <= num_sms())
      return launch_gemm_bn<64, A_MN, B_MN, Epi>(A, B, M, N, K, epi, stream, 1);
  }
  if (bn == 256) return launch_gemm_bn<256, A_MN, B_MN, Epi>(A, B, M, N, K, epi, stream, splits);
  return launch_gemm_bn<128, A_MN, B_MN, Epi>(A, B, M, N, K, epi, stream, splits);
}

// Plain epilogue used by the test hook and the weight-gradient GEMMs.
struct EpiStoreF32 {
  static constexpr const char* name = "store_f32";
  static constexpr int kPrefetchDepth = 1;
  static constexpr bool kTmaStore = false;
  struct Regs {};
  struct ColRegs {};
  float* C;
  int64_t ldc;
  __device__ __forceinline__ void prefetch(int, int, int, int, int, int) const {}
  __device__ __forceinline__ void load_col(int, ColRegs&) const {}
  __device__ __forceinline__ void load(int, int, Regs&) const {}
  __device__ __forceinline__ void frag(int row, int col, float4 acc, const Regs&, const ColRegs&) const {
    *reinterpret_cast<float4*>(C + (int64_t)row * ldc + col) = acc;
  }
};

}  // namespace mfac
