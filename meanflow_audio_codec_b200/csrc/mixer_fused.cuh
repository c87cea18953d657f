// Fused channel-mixing MLP of the MLP-Mixer block (models/mlp_mixer.py:89-93):
//     u3[tok, :] = W2^T gelu(W1^T a2[tok, :] + b1) + b2 + u[tok, :]            a2, u3: [B * tokens, CH], hidden width Hc
// As two GEMMs the hidden tensor [B * tokens, Hc] (1 GB in bf16 at B = 256, tokens = 1024, Hc = 2048) is written to HBM by the
// first and read back by the second, per block.  Here it never leaves the SM: per 128-token tile the first product's 64-column
// chunks go TMEM -> registers (bias, GELU) -> a 128B-swizzled K-major shared-memory tile WRITTEN BY THE EPILOGUE WARPS, which the
// tensor core then consumes as the A operand of the second product (accumulated over the chunks in a second TMEM region).
//   warp 0      TMA: both weight matrices once per CTA (64 + 64 KB at Hc = 2048), then one [128 x CH] activation tile per tile
//   warp 1      tcgen05.mma issuer: MMA1(c + 1) is in flight while the epilogue works on chunk c and MMA2(c) waits for its tile
//   warps 4..11 epilogue / A-operand producers (TMEM lane quarter = warp & 3, 32 of the chunk's 64 columns each)
// Both weight operands land in shared memory as [CH rows x 64 columns] 128B-swizzled boxes: W1 [CH, Hc] read as the MN-major B
// operand of MMA1 (K = CH), the transposed copy W2^T [CH, Hc] as the K-major B operand of MMA2 (N = CH).
#pragma once

#include "epilogues.cuh"
#include "gemm.cuh"

namespace mfac {

constexpr int CM_THREADS = 384;
constexpr int CM_CHUNK = 64;      // hidden columns per chunk
constexpr int CM_TMEM_COLS = 256; // acc1: 2 x 64 columns at 0 / 64; acc2: 32 columns at 128

struct ChannelMixArgs {
  const float* b1;        // [Hc]
  const float* b2;        // [CH]
  const float* res;       // [Mtok, CH] fp32 (u)
  __nv_bfloat16* out;     // [Mtok, CH] bf16 (u3)
  int Mtok, Hc;
};

inline size_t channel_mix_smem(int Hc) { return (size_t)(Hc / CM_CHUNK) * 2048 * 2 + 2 * 16384 + 2 * 16384 + 1024 + 256; }

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
  uint8_t* sH = sX + 2 * 16384;                // 2 x [128 x 64] hidden chunk, written by the epilogue warps
  uint64_t* bars = reinterpret_cast<uint64_t*>(sH + 2 * 16384);
  uint64_t* w_full = bars;                     // 1
  uint64_t* x_full = bars + 1;                 // 2
  uint64_t* x_empty = bars + 3;                // 2
  uint64_t* acc1_full = bars + 5;              // 2
  uint64_t* acc1_empty = bars + 7;             // 2
  uint64_t* h_full = bars + 9;                 // 2
  uint64_t* h_empty = bars + 11;               // 2
  uint64_t* acc2_full = bars + 13;             // 1
  uint64_t* acc2_empty = bars + 14;            // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = ceil_div(a.Mtok, GEMM_BM);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2t);
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
      mbar_init(&acc1_full[s], 1);
      mbar_init(&acc1_empty[s], 8);   // the 8 epilogue warps
      mbar_init(&h_full[s], 8);
      mbar_init(&h_empty[s], 1);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 4);         // the 4 warps that read acc2
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
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(&x_empty[s], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&x_full[s], 16384);
        tma_load_2d(sX + s * 16384, &tmX, &x_full[s], 0, tile * GEMM_BM);   // columns >= CH are zero-filled
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(GEMM_BM, CM_CHUNK, false, true);   // hidden chunk = X W1[:, chunk]
      constexpr uint32_t idesc2 = umma_idesc_bf16(GEMM_BM, CH, false, false);        // out += H_chunk W2[chunk, :]
      mbar_wait(w_full, 0);
      tc_fence_after();
      uint32_t it = 0, n1 = 0, n2 = 0;   // tiles done, MMA1 chunks issued, MMA2 chunks issued (global counters -> barrier phases)
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(&x_full[s], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t x_addr = smem_u32(sX + s * 16384);
        auto mma1 = [&](int c) {
          const uint32_t b = n1 & 1;
          mbar_wait(&acc1_empty[b], ((n1 >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint64_t ad = umma_smem_desc_sw128(x_addr, 16, 1024);
          const uint64_t bd = umma_smem_desc_sw128(smem_u32(sW1 + c * 2048), 2048, 1024);
          umma_bf16(tmem_base + b * CM_CHUNK, ad, bd, idesc1, 0u);
          umma_commit(&acc1_full[b]);
          ++n1;
        };
        mma1(0);
        mbar_wait(acc2_empty, (it & 1) ^ 1);   // the previous tile's result has been read out of acc2
        tc_fence_after();
        for (int c = 0; c < NC; ++c) {
          if (c + 1 < NC) mma1(c + 1);
          else umma_commit(&x_empty[s]);       // all MMA1s of this tile issued: the activation tile is free once they retire
          const uint32_t b = n2 & 1;
          mbar_wait(&h_full[b], (n2 >> 1) & 1);
          tc_fence_after();
          const uint32_t h_addr = smem_u32(sH + b * 16384), w2_addr = smem_u32(sW2 + c * 2048);
#pragma unroll
          for (int kk = 0; kk < CM_CHUNK / 16; ++kk) {
            const uint64_t ad = umma_smem_desc_sw128(h_addr + kk * 32, 16, 1024);
            const uint64_t bd = umma_smem_desc_sw128(w2_addr + kk * 32, 16, 1024);
            umma_bf16(tmem_base + 128, ad, bd, idesc2, (c > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&h_empty[b]);
          ++n2;
        }
        umma_commit(acc2_full);
      }
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3, hf = (warp - 4) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    uint32_t it = 0, n = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      for (int c = 0; c < NC; ++c, ++n) {
        const uint32_t b = n & 1;
        mbar_wait(&acc1_full[b], (n >> 1) & 1);
        tc_fence_after();
        float acc[32];
        tmem_ld_32x32(tmem_base + lane_sel + b * CM_CHUNK + hf * 32, acc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[b]);
        uint32_t o[16];
        const float* bias = a.b1 + c * CM_CHUNK + hf * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = add4(make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]), ldg_f4(bias + 4 * q));
          const float4 g = gelu_fast4(v);
          o[2 * q] = pack_bf16(g.x, g.y);
          o[2 * q + 1] = pack_bf16(g.z, g.w);
        }
        mbar_wait(&h_empty[b], ((n >> 1) & 1) ^ 1);   // MMA2 of the chunk that used this tile two chunks ago has retired
        // K-major 128B-swizzled A tile: row r at r * 128 B, 16-byte piece p at p ^ (r & 7); this lane owns pieces 4 hf .. 4 hf + 3
        uint4* rowp = reinterpret_cast<uint4*>(sH + b * 16384 + row_in_tile * 128);
#pragma unroll
        for (int p = 0; p < 4; ++p)
          rowp[(4 * hf + p) ^ (row_in_tile & 7)] = make_uint4(o[4 * p], o[4 * p + 1], o[4 * p + 2], o[4 * p + 3]);
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_full[b]);
      }
      // tile result: acc2 + b2 + residual -> bf16 (the hf == 0 warps own the 16 real columns)
      if (hf == 0) {
        mbar_wait(acc2_full, it & 1);
        tc_fence_after();
        float acc[32];
        tmem_ld_32x32(tmem_base + lane_sel + 128, acc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc2_empty);
        const int64_t row = (int64_t)tile * GEMM_BM + row_in_tile;
        if (row < a.Mtok) {
          uint32_t o[CH / 2];
#pragma unroll
          for (int q = 0; q < CH / 4; ++q) {
            const float4 r = *reinterpret_cast<const float4*>(a.res + row * CH + 4 * q), bb = ldg_f4(a.b2 + 4 * q);
            o[2 * q] = pack_bf16(acc[4 * q] + bb.x + r.x, acc[4 * q + 1] + bb.y + r.y);
            o[2 * q + 1] = pack_bf16(acc[4 * q + 2] + bb.z + r.z, acc[4 * q + 3] + bb.w + r.w);
          }
          uint4* dst = reinterpret_cast<uint4*>(a.out + row * CH);
          dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
          dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
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
inline int channel_mix_fused(const __nv_bfloat16* a2, const MfacDense& ch1, const MfacDense& ch2, __nv_bfloat16* w2t_scratch,
                             const float* res, __nv_bfloat16* out, int64_t Mtok, int CH, int Hc, cudaStream_t s) {
  static const bool off = getenv("MFAC_NO_FUSED_CHANNEL_MIX") != nullptr;
  if (off || CH != 16 || Hc % CM_CHUNK != 0 || Hc < CM_CHUNK || channel_mix_smem(Hc) > 227 * 1024 || Mtok > 0x7fffffff)
    return MFAC_ERR_UNSUPPORTED;
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
