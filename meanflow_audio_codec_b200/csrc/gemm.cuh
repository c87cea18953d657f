// Persistent, warp-specialised tcgen05 GEMM for sm_100a with pluggable fused epilogues.
//
//   C[M,N] = A[M,K] * B[K,N]      bf16 operands, fp32 accumulation in TMEM
//
// One CTA per SM (grid = min(#tiles, #SMs)), 192 threads:
//   warp 0      TMA producer   -- cp.async.bulk.tensor tiles of A and B into a ring of
//                                 128B-swizzled shared-memory stages (mbarrier full/empty)
//   warp 1      MMA issuer     -- one elected lane issues tcgen05.mma (M=128, N=BN, K=16);
//                                 tcgen05.commit releases smem stages and publishes the tile
//   warps 2..5  epilogue       -- tcgen05.ld the fp32 accumulator (thread == tile row),
//                                 apply the fused epilogue functor, write global memory
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
// rows < M and 32-column groups that start below N (N must be a multiple of 32).
#pragma once

#include "mfac_common.cuh"

namespace mfac {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_SMEM_BUDGET = 196608;  // bytes of operand stages

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = GEMM_SMEM_BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");
};

struct GemmShape {
  int M, N, K;
};

template <int BN, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmShape shape,
                    Epi epi) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
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
  const int k_blocks = ceil_div(shape.K, GEMM_BK);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t kit = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * GEMM_BM;
        const int n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++kit) {
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
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, A_MN, B_MN);
      uint32_t kit = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + s * Cfg::B_BYTES);
#pragma unroll
          for (int kk = 0; kk < GEMM_BK / 16; ++kk) {
            // K-major: advance 16 elements = 32 B inside the 128-B swizzle row.
            // MN-major: advance 16 k-rows of 128 B = 2048 B (two 8-row swizzle atoms, SBO apart).
            const uint64_t ad = A_MN ? umma_smem_desc_sw128(a_addr + kk * 2048, 8192, 1024)
                                     : umma_smem_desc_sw128(a_addr + kk * 32, 16, 1024);
            const uint64_t bd = B_MN ? umma_smem_desc_sw128(b_addr + kk * 2048, 8192, 1024)
                                     : umma_smem_desc_sw128(b_addr + kk * 32, 16, 1024);
            umma_bf16(tmem_d, ad, bd, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // smem stage reusable once these MMAs retire
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * GEMM_BM;
      const int n0 = (tile % n_tiles) * BN;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float acc[32];
        tmem_ld_32x32(taddr + c * 32, acc);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row < shape.M && col0 < shape.N) epi(row, col0, acc);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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
  epi(row, col0, acc);
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
void* profile_begin(int family, double work, cudaStream_t s);
void profile_end(void* token, cudaStream_t s);

template <int BN, bool A_MN, bool B_MN, class Epi>
int launch_gemm_bn(const GemmOperandDesc& A, const GemmOperandDesc& B, int M, int N, int K, const Epi& epi,
                   cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
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
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, Epi>;
  static bool configured = false;  // one per template instantiation
  if (!configured) {
    MFAC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = ceil_div(M, GEMM_BM) * ceil_div(N, BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  GemmShape shape{M, N, K};
  void* prof = profile_begin(MFAC_PROF_GEMM, 2.0 * (double)M * (double)N * (double)K, stream);
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, shape, epi);
  profile_end(prof, stream);
  count_launch();
  return launch_status();
}

// C = A * B with the given epilogue.  N must be a multiple of 32; K and M are arbitrary
// (TMA zero-fills), leading dimensions must be multiples of 8 elements (16-byte TMA strides).
template <bool A_MN, bool B_MN, class Epi>
int launch_gemm(const GemmOperandDesc& A, const GemmOperandDesc& B, int M, int N, int K, const Epi& epi,
                cudaStream_t stream, int force_bn = 0) {
  if (M <= 0 || N <= 0 || K <= 0) return MFAC_ERR_BAD_SHAPE;
  if (N % 32 != 0 || A.ld % 8 != 0 || B.ld % 8 != 0) return MFAC_ERR_UNSUPPORTED;
  if (A.mn_major != A_MN || B.mn_major != B_MN) return MFAC_ERR_UNSUPPORTED;
  if (simt_gemm_enabled()) {
    SimtOperand a{reinterpret_cast<const __nv_bfloat16*>(A.ptr), A.ld, A_MN ? 1 : 0};
    SimtOperand b{reinterpret_cast<const __nv_bfloat16*>(B.ptr), B.ld, B_MN ? 1 : 0};
    const int64_t threads = (int64_t)M * (N / 32);
    gemm_simt_kernel<Epi><<<(unsigned)ceil_div<int64_t>(threads, 128), 128, 0, stream>>>(a, b, GemmShape{M, N, K}, epi);
    count_launch();
    return launch_status();
  }
  // Tile choice: BN=256 halves the per-MMA shared-memory traffic (96 vs 128 B/clk) but only
  // pays when it does not cost a wave; prefer it when N divides and the tile count still
  // covers the machine.
  int bn = force_bn;
  if (bn == 0) {
    const int tiles256 = ceil_div(M, GEMM_BM) * ceil_div(N, 256);
    bn = (N % 256 == 0 && tiles256 >= num_sms()) ? 256 : 128;
  }
  if (bn == 256) return launch_gemm_bn<256, A_MN, B_MN, Epi>(A, B, M, N, K, epi, stream);
  return launch_gemm_bn<128, A_MN, B_MN, Epi>(A, B, M, N, K, epi, stream);
}

// Plain epilogue used by the test hook and the weight-gradient GEMMs.
struct EpiStoreF32 {
  float* C;
  int64_t ldc;
  __device__ __forceinline__ void operator()(int row, int col0, float (&acc)[32]) const {
    store_f32x32(C + (int64_t)row * ldc + col0, acc);
  }
};

}  // namespace mfac
