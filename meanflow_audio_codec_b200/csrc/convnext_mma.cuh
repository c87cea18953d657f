// ConvNeXt block of the convolutional velocity network (models/conv_flow.py:69-115, 176-193) on warp-level tensor-core MMAs,
// one CTA per sample, for the reference geometry (16 channels, S x S image with S a multiple of 16):
//     xs[S,S,16] fp32 -> LN(16) -> FiLM -> [3x3 SAME conv + b -> LN(16) -> 1x1 (16 -> 32) + b -> GELU -> GRN -> 1x1 (32 -> 16) + b]
//                     * layer_scale + skip -> bf16 [S*S*16]
// The op is HBM-bound by its roofline (96 KB per sample and block against 7 MFLOP) but the thread-per-pixel fp32 version spent
// 1.08 ms per launch at 2048 samples (35x its HBM time) on shared-memory weight reads.  Here every product is an implicit GEMM
// over 16-pixel row segments (M = 16 pixels, K = 16 input channels per tap, N = 8 output channels per MMA):
//   * the LN + FiLM'd image lives in shared memory as bf16 [P*P pixels][16 channels] (zero halo); the A fragment of tap (dy, dx)
//     is ONE ldmatrix.x4 at a shifted pixel address (the two 16-byte channel halves of a pixel are XOR-swizzled by bit 2 of the
//     pixel index so that the eight rows of an 8x8 matrix fall into distinct banks);
//   * the 3x3 weights sit in shared memory already in B-fragment order (one LDS.128 per tap and lane), the 1x1 weights in registers;
//   * the accumulator layout of two adjacent N = 8 tiles IS the A-fragment layout of the next product, so conv -> LN -> expand ->
//     GELU -> (GRN) -> contract chains through registers; the GELU'd hidden activations wait for the sample-wide GRN statistic as
//     packed bf16 fragments in shared memory, each lane writing and later reading only its own words (conflict-free; keeping
//     them in registers -- 64 per thread at S = 32 -- spilled 0.85 GB per launch to local memory);
//   * the skip operand is recomputed in fp32 from the L2-hot input (LN statistics by a quad shuffle), so the block output is
//     rounded to bf16 once and two CTAs fit one SM.
// mma.sync rather than tcgen05: N is 16 / 32 and K = 16 per tap -- a 128-row tcgen05 tile would need the im2col operand
// materialised; the warp-level MMA reads the shifted views in place and the kernel is bound by loads / issue, not tensor rate.
#pragma once

#include "epilogues.cuh"

namespace mfac {

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr));
}
// D(16x8, fp32) += A(16x16, bf16, row) B(16x8, bf16, col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int S>
constexpr size_t convnext_mma_smem() {
  return (size_t)(S + 2) * (S + 2) * 32 + (size_t)S * S * 32 * 2 + 9 * 32 * 16 + 64 * 4;
}

// B-fragment registers of a [K, N] fp32 weight matrix (row-major, ld = N) for k-step ks and n-tile nt:
//   b0 = (k = 16 ks + 2t, 2t + 1; n = 8 nt + g), b1 = the same 8 rows further down
__device__ __forceinline__ void load_b_frag(const float* __restrict__ W, int ld, int ks, int nt, int g, int t, uint32_t& b0, uint32_t& b1) {
  const float* p = W + (size_t)(16 * ks + 2 * t) * ld + 8 * nt + g;
  b0 = pack_bf16(p[0], p[ld]);
  b1 = pack_bf16(p[8 * ld], p[9 * ld]);
}

template <int S>
__global__ void __launch_bounds__(256, 2) convnext_block_mma_kernel(const float* __restrict__ xs, const float* __restrict__ film,
                                                                    MfacConvBlockW w, __nv_bfloat16* __restrict__ out) {
  constexpr int CH = 16, P = S + 2, NPIX = S * S, MT = NPIX / 16, MT_W = MT / 8, TPR = S / 16;   // m-tiles, per warp, per image row
  static_assert(S % 16 == 0 && MT % 8 == 0 && NPIX % 256 == 0, "16-pixel row segments, eight warps, whole pixels per thread");
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* sIn = smem_raw;                                                   // [P*P][32 B] bf16, halves swizzled, zero halo
  uint32_t* sH = reinterpret_cast<uint32_t*>(smem_raw + P * P * 32);         // [MT][8 fragment words][32 lanes] GELU'd hidden, bf16 pairs
  uint4* sB3 = reinterpret_cast<uint4*>(sH + MT * 8 * 32);                   // [9 taps][32 lanes]: {nt0 b0, nt0 b1, nt1 b0, nt1 b1}
  float* sG = reinterpret_cast<float*>(sB3 + 9 * 32);                        // [32] GRN sums -> scale; [32..63] beta
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int64_t b = blockIdx.x;
  const float* fb = film + b * 2 * CH;

  // this thread's pixels of the input: all loads in flight before anything else (one HBM round trip, under the set-up below)
  constexpr int PPT = NPIX / 256;
  float4 raw[PPT][4];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const float4* src = reinterpret_cast<const float4*>(xs + (b * NPIX + tid + 256 * i) * CH);
#pragma unroll
    for (int c = 0; c < 4; ++c) raw[i][c] = __ldg(src + c);
  }
  for (int i = tid; i < P * P * 2; i += 256) reinterpret_cast<uint4*>(sIn)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 9 * 32; i += 256) {
    const int tap = i >> 5, l = i & 31, gg = l >> 2, tt = l & 3;
    uint4 f;
    load_b_frag(w.conv3_w + tap * CH * CH, CH, 0, 0, gg, tt, f.x, f.y);
    load_b_frag(w.conv3_w + tap * CH * CH, CH, 0, 1, gg, tt, f.z, f.w);
    sB3[i] = f;
  }
  if (tid < 32) { sG[tid] = 0.f; sG[32 + tid] = w.grn_beta ? w.grn_beta[tid] : 0.f; }
  __syncthreads();

  // LN over channels + FiLM -> bf16 into the padded plane (the conv operand)                          (conv_flow.py:176-187)
  {
    float f1[CH], f0[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { f1[c] = 1.0f + fb[c]; f0[c] = fb[CH + c]; }
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const int pix = tid + 256 * i;
      float v[CH];
#pragma unroll
      for (int c = 0; c < 4; ++c) { v[4 * c] = raw[i][c].x; v[4 * c + 1] = raw[i][c].y; v[4 * c + 2] = raw[i][c].z; v[4 * c + 3] = raw[i][c].w; }
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) { sum += v[c]; sq += v[c] * v[c]; }
      const float mu = sum * (1.0f / CH);
      const float rstd = rsqrtf(fmaxf(0.f, sq * (1.0f / CH) - mu * mu) + LN_EPS);
#pragma unroll
      for (int c = 0; c < CH; ++c) v[c] = f1[c] * ((v[c] - mu) * rstd) + f0[c];
      const int pp = ((pix / S) + 1) * P + (pix % S) + 1;
      const int sw = (pp >> 2) & 1;
      uint4* dst = reinterpret_cast<uint4*>(sIn + pp * 32);
      dst[0 ^ sw] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      dst[1 ^ sw] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
    }
  }
  __syncthreads();

  // ---- pass 1: 3x3 conv -> LN -> 1x1 expand -> GELU -> hidden fragments to shared memory; GRN sums of squares
  {
    uint32_t w1[4][2];
    float b3[2][2], b1[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      load_b_frag(w.pw1_w, 2 * CH, 0, nt, g, t, w1[nt][0], w1[nt][1]);
      b1[nt][0] = w.pw1_b[8 * nt + 2 * t];
      b1[nt][1] = w.pw1_b[8 * nt + 2 * t + 1];
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) { b3[nt][0] = w.conv3_b[8 * nt + 2 * t]; b3[nt][1] = w.conv3_b[8 * nt + 2 * t + 1]; }
    float gsq[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) gsq[nt][0] = gsq[nt][1] = 0.f;
    // ldmatrix row of this lane: matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15)
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhalf = lane >> 4;
    const uint32_t sIn_addr = smem_u32(sIn);
#pragma unroll 1
    for (int i = 0; i < MT_W; ++i) {
      const int m = warp + 8 * i, y = m / TPR, x0 = (m % TPR) * 16;
      float acc[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { acc[nt][0] = acc[nt][2] = b3[nt][0]; acc[nt][1] = acc[nt][3] = b3[nt][1]; }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int pp = (y + tap / 3) * P + x0 + lrow + tap % 3;
        uint32_t a[4];
        ldmatrix_x4(a, sIn_addr + pp * 32 + ((lhalf ^ ((pp >> 2) & 1)) << 4));
        const uint4 f = sB3[tap * 32 + lane];
        mma_bf16_16816(acc[0], a, f.x, f.y);
        mma_bf16_16816(acc[1], a, f.z, f.w);
      }
      // LayerNorm over the 16 channels of pixel rows g and g + 8 (a row is spread over the four lanes of a quad)
      float s0 = acc[0][0] + acc[0][1] + acc[1][0] + acc[1][1], s1 = acc[0][2] + acc[0][3] + acc[1][2] + acc[1][3];
      float q0 = acc[0][0] * acc[0][0] + acc[0][1] * acc[0][1] + acc[1][0] * acc[1][0] + acc[1][1] * acc[1][1];
      float q1 = acc[0][2] * acc[0][2] + acc[0][3] * acc[0][3] + acc[1][2] * acc[1][2] + acc[1][3] * acc[1][3];
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
      }
      const float mu0 = s0 * (1.0f / CH), mu1 = s1 * (1.0f / CH);
      const float r0 = rsqrtf(fmaxf(0.f, q0 * (1.0f / CH) - mu0 * mu0) + LN_EPS), r1 = rsqrtf(fmaxf(0.f, q1 * (1.0f / CH) - mu1 * mu1) + LN_EPS);
      uint32_t a[4];
      a[0] = pack_bf16((acc[0][0] - mu0) * r0, (acc[0][1] - mu0) * r0);
      a[1] = pack_bf16((acc[0][2] - mu1) * r1, (acc[0][3] - mu1) * r1);
      a[2] = pack_bf16((acc[1][0] - mu0) * r0, (acc[1][1] - mu0) * r0);
      a[3] = pack_bf16((acc[1][2] - mu1) * r1, (acc[1][3] - mu1) * r1);
      uint32_t* hw = sH + m * 256 + lane;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float h[4] = {b1[nt][0], b1[nt][1], b1[nt][0], b1[nt][1]};
        mma_bf16_16816(h, a, w1[nt][0], w1[nt][1]);
        const float2 ga = gelu_fast2(make_float2(h[0], h[1])), gb = gelu_fast2(make_float2(h[2], h[3]));   // tanh.approx, as the GEMM epilogues
        gsq[nt][0] += ga.x * ga.x + gb.x * gb.x;
        gsq[nt][1] += ga.y * ga.y + gb.y * gb.y;
        hw[(2 * nt) * 32] = pack_bf16(ga.x, ga.y);        // row g,     channels 8 nt + 2t, + 1
        hw[(2 * nt + 1) * 32] = pack_bf16(gb.x, gb.y);    // row g + 8
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v = gsq[nt][j];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0) atomicAdd(&sG[8 * nt + 2 * t + j], v);
      }
  }
  __syncthreads();
  if (tid == 0) {   // GRN: gx = ||v||_2 over the image, divided by its channel mean; scale = gamma + gx    (conv_flow.py:33-45)
    float mean = 0.f;
    for (int c = 0; c < 2 * CH; ++c) { sG[c] = sqrtf(sG[c]); mean += sG[c]; }
    mean = mean / (float)(2 * CH) + 1e-6f;
    for (int c = 0; c < 2 * CH; ++c) sG[c] = (w.grn_gamma ? w.grn_gamma[c] : 0.f) + sG[c] / mean;
  }
  __syncthreads();

  // ---- pass 2: GRN -> 1x1 contract -> layer scale -> + skip -> bf16
  uint32_t w2[2][2][2];
  float sc[4][2], be[4][2], b2[2][2], ls[2][2], f1[2][2], f0[2][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) load_b_frag(w.pw2_w, CH, ks, nt, g, t, w2[ks][nt][0], w2[ks][nt][1]);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) { sc[nt][j] = sG[8 * nt + 2 * t + j]; be[nt][j] = sG[32 + 8 * nt + 2 * t + j]; }
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = 8 * nt + 2 * t + j;
      b2[nt][j] = w.pw2_b[c];
      ls[nt][j] = w.layer_scale ? w.layer_scale[c] : 1.0f;
      f1[nt][j] = 1.0f + fb[c];
      f0[nt][j] = fb[CH + c];
    }
#pragma unroll 1
  for (int i = 0; i < MT_W; ++i) {
    const int m = warp + 8 * i, pix0 = m * 16;
    // the skip operand: LN + FiLM of the raw input again, in fp32 (this lane's four channels of rows g and g + 8)
    float2 xr[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
        xr[r][nt] = *reinterpret_cast<const float2*>(xs + (b * NPIX + pix0 + g + 8 * r) * CH + 8 * nt + 2 * t);
    const uint32_t* hw = sH + m * 256 + lane;
    float o[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) { o[nt][0] = o[nt][2] = b2[nt][0]; o[nt][1] = o[nt][3] = b2[nt][1]; }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t a[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {           // a0: (row g, n-tile 2 ks), a1: (row g + 8, same), a2 / a3: n-tile 2 ks + 1
        const int nt = 2 * ks + (q >> 1);
        const float2 v = unpack_bf16(hw[(2 * nt + (q & 1)) * 32]);
        a[q] = pack_bf16(v.x * sc[nt][0] + be[nt][0], v.y * sc[nt][1] + be[nt][1]);
      }
      mma_bf16_16816(o[0], a, w2[ks][0][0], w2[ks][0][1]);
      mma_bf16_16816(o[1], a, w2[ks][1][0], w2[ks][1][1]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float sum = xr[r][0].x + xr[r][0].y + xr[r][1].x + xr[r][1].y;
      float sq = xr[r][0].x * xr[r][0].x + xr[r][0].y * xr[r][0].y + xr[r][1].x * xr[r][1].x + xr[r][1].y * xr[r][1].y;
#pragma unroll
      for (int sh = 1; sh <= 2; sh <<= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, sh); sq += __shfl_xor_sync(0xffffffffu, sq, sh); }
      const float mu = sum * (1.0f / CH);
      const float rstd = rsqrtf(fmaxf(0.f, sq * (1.0f / CH) - mu * mu) + LN_EPS);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float sk0 = f1[nt][0] * ((xr[r][nt].x - mu) * rstd) + f0[nt][0], sk1 = f1[nt][1] * ((xr[r][nt].y - mu) * rstd) + f0[nt][1];
        *reinterpret_cast<uint32_t*>(out + (b * NPIX + pix0 + g + 8 * r) * CH + 8 * nt + 2 * t) =
            pack_bf16(o[nt][2 * r] * ls[nt][0] + sk0, o[nt][2 * r + 1] * ls[nt][1] + sk1);
      }
    }
  }
}

}  // namespace mfac
