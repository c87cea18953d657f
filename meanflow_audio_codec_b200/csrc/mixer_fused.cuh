// Fused channel-mixing MLP of the MLP-Mixer block (models/mlp_mixer.py:89-93):
//     u3[tok, :] = W2^T gelu(W1^T a2[tok, :] + b1) + b2 + u[tok, :]            a2, u3: [B * tokens, CH], hidden width Hc
// As two GEMMs the hidden tensor [B * tokens, Hc] (1 GB in bf16 at B = 256, tokens = 1024, Hc = 2048) is written to HBM by the
// first and read back by the second, per block.  Here it never leaves the SM: per 128-token tile the first product's 64-column
// chunks go TMEM -> registers (bias, GELU) -> a 128B-swizzled K-major shared-memory tile WRITTEN BY THE EPILOGUE WARPS, which the
// tensor core then consumes as the A operand of the second product (accumulated over the chunks in a second TMEM region).
//   warp 0      TMA: both weight matrices once per CTA (64 + 64 KB at Hc = 2048), then one [128 x CH] activation tile per tile
//   warp 1 / 2  tcgen05.mma issuers of the first / second product (four TMEM accumulator stages, four hidden-tile stages)
//   warps 4..19 epilogue / A-operand producers: four TEAMS of four warps (TMEM lane quarter = warp & 3); team k owns stage k of
//               the accumulator and of the hidden tile and handles chunks k, k + 4, ..., so four chunks are in flight and a
//               warp's TMEM-load -> GELU -> STS -> fence.proxy.async -> arrive latency chain overlaps the other teams' work
// History (profiles/r02_channel_mix_fused.txt): the first version had ONE thread issuing both products and ran at one 64-column
// chunk per ~1400 clocks whatever the epilogue cost (GELU removed: -4 %) -- with K = 16 per product the tensor pipe wants a new
// instruction every few dozen clocks and ~330 SASS instructions per chunk on one thread could not supply them.
// Both weight operands land in shared memory as [CH rows x 64 columns] 128B-swizzled boxes: W1 [CH, Hc] read as the MN-major B
// operand of MMA1 (K = CH), the transposed copy W2^T [CH, Hc] as the K-major B operand of MMA2 (N = CH).
#pragma once

#include "epilogues.cuh"
#include "gemm.cuh"

namespace mfac {

constexpr int CM_EPI_WARPS = 16;
constexpr int CM_THREADS = 128 + 32 * CM_EPI_WARPS;
constexpr int CM_CHUNK = 64;      // hidden columns per chunk
constexpr int CM_STAGES = 4;      // acc1 (TMEM) and hidden-tile (shared memory) stages
constexpr int CM_ACC2_COL = CM_STAGES * CM_CHUNK;   // acc2: 2 x 32 columns behind the acc1 stages
constexpr int CM_TMEM_COLS = 512;
constexpr int CM_TEAMS = CM_EPI_WARPS / 4;
static_assert(CM_TEAMS == CM_STAGES, "team k owns stage k");

struct ChannelMixArgs {
  const float* b1;        // [Hc]
  const float* b2;        // [CH]
  const float* res;       // [Mtok, CH] fp32 (u)
  __nv_bfloat16* out;     // [Mtok, CH] bf16 (u3)
  int Mtok, Hc;
};

inline size_t channel_mix_smem(int Hc) {
  return (size_t)(Hc / CM_CHUNK) * 2048 * 2 + 2 * 16384 + CM_STAGES * 16384 + 1024 + 512;
}

template <int CH>
__global__ void __launch_bounds__(CM_THREADS, 1)
channel_mix_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                         const __grid_constant__ CUtensorMap tmW2t, ChannelMixArgs a) {
  static_assert(CH == 16, "one K = 16 step for MMA1 and N = 16 for MMA2");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int NC = a.Hc / CM_CHUNK;
  uint8_t* sW1 = smem;                         // NC boxes of [CH x 64] (2 KB each)
  uint8_t* sW2 = sW1 + NC * 2048;
  uint8_t* sX = sW2 + NC * 2048;               // 2 x [128 x 64] (K padded with zeros by TMA)
  uint8_t* sH = sX + 2 * 16384;                // CM_STAGES x [128 x 64] hidden chunk, written by the epilogue warps
  uint64_t* bars = reinterpret_cast<uint64_t*>(sH + CM_STAGES * 16384);
  uint64_t* w_full = bars;                     // 1
  uint64_t* x_full = bars + 1;                 // 2
  uint64_t* x_empty = bars + 3;                // 2
  uint64_t* acc2_full = bars + 5;              // 2
  uint64_t* acc2_empty = bars + 7;             // 2
  uint64_t* acc1_full = bars + 9;              // CM_STAGES each from here on
  uint64_t* acc1_empty = acc1_full + CM_STAGES;
  uint64_t* h_full = acc1_empty + CM_STAGES;
  uint64_t* h_empty = h_full + CM_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_empty + CM_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = ceil_div(a.Mtok, GEMM_BM);
  const int my_tiles = blockIdx.x < tiles ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2t);
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
      mbar_init(&acc2_full[s], 1);
      mbar_init(&acc2_empty[s], 4);              // the 4 warps that read acc2
    }
    for (int s = 0; s < CM_STAGES; ++s) {
      mbar_init(&acc1_full[s], 1);
      mbar_init(&acc1_empty[s], 4);              // the four warps of the team that owns the stage
      mbar_init(&h_full[s], 4);
      mbar_init(&h_empty[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, CM_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, (uint32_t)NC * 2048u * 2u);
      for (int c = 0; c < NC; ++c) {
        tma_load_2d(sW1 + c * 2048, &tmW1, w_full, c * CM_CHUNK, 0);
        tma_load_2d(sW2 + c * 2048, &tmW2t, w_full, c * CM_CHUNK, 0);
      }
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t & 1;
        mbar_wait(&x_empty[s], ((t >> 1) & 1) ^ 1);
        mbar_expect_tx(&x_full[s], 16384);
        tma_load_2d(sX + s * 16384, &tmX, &x_full[s], 0, ((int)blockIdx.x + t * (int)gridDim.x) * GEMM_BM);   // columns >= CH: zero fill
      }
    }
  } else if (warp == 1) {
    // MMA1 issuer: hidden chunk = X W1[:, chunk] into accumulator stage g mod CM_STAGES (chunks numbered over this CTA's tiles).
    // The two issuers are separate warps with their descriptors precomputed: with K = 16 per product the tensor pipe needs a new
    // instruction every few dozen clocks, and ONE thread issuing both products (~330 SASS instructions per chunk) was the
    // kernel's bottleneck (profiles/r02_channel_mix_fused.txt).
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(GEMM_BM, CM_CHUNK, false, true);
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint64_t w1_desc = umma_smem_desc_sw128(smem_u32(sW1), 2048, 1024);
      uint32_t g = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const uint32_t xs = t & 1;
        mbar_wait(&x_full[xs], (t >> 1) & 1);
        tc_fence_after();
        const uint64_t x_desc = umma_smem_desc_sw128(smem_u32(sX + xs * 16384), 16, 1024);
        for (int c = 0; c < NC; ++c, ++g) {
          const uint32_t b = g % CM_STAGES;
          mbar_wait(&acc1_empty[b], ((g / CM_STAGES) & 1) ^ 1);
          tc_fence_after();
          umma_bf16(tmem_base + b * CM_CHUNK, x_desc, w1_desc + (uint64_t)(c * (2048 >> 4)), idesc1, 0u);
          umma_commit(&acc1_full[b]);
        }
        umma_commit(&x_empty[xs]);   // all MMA1s of this tile issued: the activation tile is free once they retire
      }
    }
  } else if (warp == 2) {
    // MMA2 issuer: out += H_chunk W2[chunk, :] into acc2 stage (tile & 1)
    if (lane == 0) {
      constexpr uint32_t idesc2 = umma_idesc_bf16(GEMM_BM, CH, false, false);
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint64_t w2_desc = umma_smem_desc_sw128(smem_u32(sW2), 16, 1024);
      const uint64_t h_desc = umma_smem_desc_sw128(smem_u32(sH), 16, 1024);
      uint32_t g = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const uint32_t as = t & 1;
        mbar_wait(&acc2_empty[as], ((t >> 1) & 1) ^ 1);   // the result two tiles back has been read out of this acc2 stage
        tc_fence_after();
        const uint32_t d_addr = tmem_base + CM_ACC2_COL + as * 32;
        for (int c = 0; c < NC; ++c, ++g) {
          const uint32_t b = g % CM_STAGES;
          mbar_wait(&h_full[b], (g / CM_STAGES) & 1);
          tc_fence_after();
          const uint64_t ad = h_desc + (uint64_t)(b * (16384 >> 4)), bd = w2_desc + (uint64_t)(c * (2048 >> 4));
#pragma unroll
          for (int kk = 0; kk < CM_CHUNK / 16; ++kk)
            umma_bf16(d_addr, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc2, (c > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&h_empty[b]);
        }
        umma_commit(&acc2_full[as]);
      }
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3, team = (warp - 4) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)NC;
    // K-major 128B-swizzled A tile: row r at r * 128 B, 16-byte piece p at p ^ (r & 7)
    uint4* rowp = reinterpret_cast<uint4*>(sH + team * 16384 + row_in_tile * 128);
    const uint32_t acc1_addr = tmem_base + lane_sel + team * CM_CHUNK;
    // 16 accumulator columns -> bias, GELU, bf16 -> pieces 2 j, 2 j + 1 of this lane's row of the hidden tile
    auto gelu_store = [&](const float (&acc)[16], const float* bias, int j) {
      uint32_t o[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 g = gelu_fast4(add4(make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]), ldg_f4(bias + 4 * q)));
        o[2 * q] = pack_bf16(g.x, g.y);
        o[2 * q + 1] = pack_bf16(g.z, g.w);
      }
      rowp[(2 * j) ^ (row_in_tile & 7)] = make_uint4(o[0], o[1], o[2], o[3]);
      rowp[(2 * j + 1) ^ (row_in_tile & 7)] = make_uint4(o[4], o[5], o[6], o[7]);
    };
    uint32_t ph = 0, t = 0, c = team;
    while (c >= (uint32_t)NC && NC > 0) { c -= NC; ++t; }
    for (uint32_t n = team; n < total; n += CM_TEAMS, ph ^= 1) {
      const float* bias = a.b1 + c * CM_CHUNK;
      mbar_wait(&acc1_full[team], ph);
      tc_fence_after();
      // the chunk's 64 columns in four 16-column loads, each in flight under the GELU of the one before
      float a0[16], a1[16];
      tmem_ld_32x16(acc1_addr, a0);
      tmem_ld_wait();
      tmem_ld_32x16(acc1_addr + 16, a1);
      mbar_wait(&h_empty[team], ph ^ 1);   // MMA2 of this team's previous chunk has retired: the hidden tile is free
      gelu_store(a0, bias, 0);
      tmem_ld_wait();
      tmem_ld_32x16(acc1_addr + 32, a0);
      gelu_store(a1, bias + 16, 1);
      tmem_ld_wait();
      tmem_ld_32x16(acc1_addr + 48, a1);
      gelu_store(a0, bias + 32, 2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc1_empty[team]);
      gelu_store(a1, bias + 48, 3);
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_full[team]);
      if (c == (uint32_t)NC - 1) {
        // the team that produced a tile's last chunk writes the tile: acc2 + b2 + residual -> bf16 (16 real columns)
        const int as = t & 1;
        mbar_wait(&acc2_full[as], (t >> 1) & 1);
        tc_fence_after();
        float r2[16];
        tmem_ld_32x16(tmem_base + lane_sel + CM_ACC2_COL + as * 32, r2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc2_empty[as]);
        const int64_t row = (int64_t)((int)blockIdx.x + (int)t * (int)gridDim.x) * GEMM_BM + row_in_tile;
        if (row < a.Mtok) {
          uint32_t w[CH / 2];
#pragma unroll
          for (int q = 0; q < CH / 4; ++q) {
            const float4 r = *reinterpret_cast<const float4*>(a.res + row * CH + 4 * q), bb = ldg_f4(a.b2 + 4 * q);
            w[2 * q] = pack_bf16(r2[4 * q] + bb.x + r.x, r2[4 * q + 1] + bb.y + r.y);
            w[2 * q + 1] = pack_bf16(r2[4 * q + 2] + bb.z + r.z, r2[4 * q + 3] + bb.w + r.w);
          }
          uint4* dst = reinterpret_cast<uint4*>(a.out + row * CH);
          dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      c += CM_TEAMS;
      while (c >= (uint32_t)NC) { c -= NC; ++t; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, CM_TMEM_COLS);
  }
}

// W2 [Hc, CH] (Flax [in, out]) -> W2^T [CH, Hc]
__global__ void transpose_w2_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ wt, int Hc, int CH) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Hc * CH) {
    const int h = i / CH, c = i - h * CH;
    wt[(int64_t)c * Hc + h] = w[i];
  }
}

// returns MFAC_ERR_UNSUPPORTED when the geometry is outside the fused kernel (the caller then runs the two GEMMs)
// geometry the fused kernel covers (the workspace plan skips the [tokens, Hc] hidden tensor when it does)
inline bool channel_mix_fused_ok(int CH, int Hc) {
  static const bool off = getenv("MFAC_NO_FUSED_CHANNEL_MIX") != nullptr;
  return !off && CH == 16 && Hc % CM_CHUNK == 0 && Hc >= CM_CHUNK && channel_mix_smem(Hc) <= 227 * 1024;
}
inline int channel_mix_fused(const __nv_bfloat16* a2, const MfacDense& ch1, const MfacDense& ch2, __nv_bfloat16* w2t_scratch,
                             const float* res, __nv_bfloat16* out, int64_t Mtok, int CH, int Hc, cudaStream_t s) {
  if (!channel_mix_fused_ok(CH, Hc) || Mtok > 0x7fffffff) return MFAC_ERR_UNSUPPORTED;
  transpose_w2_kernel<<<(unsigned)ceil_div(Hc * CH, 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(ch2.w), w2t_scratch, Hc, CH);
  count_launch();
  CUtensorMap tmX, tmW1, tmW2t;
  MFAC_OK(make_tmap_bf16(&tmX, a2, CH, Mtok, CH, 64, GEMM_BM));      // box wider than the 16 real columns: zero fill
  MFAC_OK(make_tmap_bf16(&tmW1, ch1.w, Hc, CH, Hc, 64, CH));
  MFAC_OK(make_tmap_bf16(&tmW2t, w2t_scratch, Hc, CH, Hc, 64, CH));
  const size_t smem = channel_mix_smem(Hc);
  static PerDeviceOnce configured;
  if (configured.need()) {
    MFAC_CUDA_OK(cudaFuncSetAttribute(channel_mix_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.done();
  }
  const int tiles = ceil_div((int)Mtok, GEMM_BM);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  ChannelMixArgs args{ch1.b, ch2.b, res, out, (int)Mtok, Hc};
  void* prof = profile_begin(MFAC_PROF_GEMM, 4.0 * (double)Mtok * CH * Hc, s, "channel_mix_fused", (int)Mtok, CH, Hc);
  launch_pdl(channel_mix_fused_kernel<16>, dim3((unsigned)grid), dim3(CM_THREADS), smem, s, tmX, tmW1, tmW2t, args);
  profile_end(prof, s);
  count_launch();
  return launch_status();
}

}  // namespace mfac
