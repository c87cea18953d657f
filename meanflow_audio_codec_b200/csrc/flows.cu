// Forward passes of the two other velocity networks behind the reference's call signature
//   model.apply({"params": p}, x[B,D], time[B,2], latents[B,T_lat,L_lat] | None) -> [B,D]
//
// ref (paths inside /root/reference/meanflow_audio_codec/):
//   ConditionalMLPMixerFlow   models/mlp_mixer.py:171-235  (block :102-163, MLPMixerBlock :14-94)
//   ConditionalConvFlow       models/conv_flow.py:213-271  (block :123-205, ConvNeXtBlock :53-115, GRN :14-45)
//
// Every Dense layer is a launch of the tcgen05 GEMM (gemm.cuh) with a fused epilogue; LayerNorm / AdaLN / FiLM, the
// token<->channel transposes and the whole ConvNeXt block (3x3 conv, LN, 1x1 expand, GELU, GRN, 1x1 contract,
// layer scale, skip) are row / tile kernels here.  Weights arrive as device pointers to bf16 kernels ([in,out]
// row-major, exactly Flax's layout) and fp32 biases; the Python mirror keeps the bf16 copies.
// Forward only: SURVEY.md R5 -- the reference never trains these two models (no encode method).
#include "gemm.cuh"
#include "epilogues.cuh"
#include "imf_layout.cuh"
#include "convnext_mma.cuh"
#include "mixer_fused.cuh"

namespace mfac {
namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)ceil_div<int64_t>(n, t); }

// y[M,N] = A[M,K] (bf16, K contiguous, ld = lda) @ W[K,N] (bf16, Flax [in,out]) with epilogue
template <class Epi>
int dense(const __nv_bfloat16* A, int lda, const MfacDense& w, int M, int N, int K, const Epi& epi, cudaStream_t s) {
  return launch_gemm<false, true>(GemmOperandDesc{A, lda, false}, GemmOperandDesc{w.w, N, true}, M, N, K, epi, s);
}

// cond[b, :] = sincos(t) + sincos(h) (+ latent_cond[b, :])   (mlp_mixer.py:220-229, conv_flow.py:256-265, utils.py:5-13)
__global__ void flow_cond_kernel(const float* __restrict__ time, const float* __restrict__ latent_cond,
                                 __nv_bfloat16* __restrict__ cond, int C) {
  const int64_t b = blockIdx.x;
  const float t = time[2 * b], h = time[2 * b + 1];
  const int half = C / 2;
  for (int j = threadIdx.x; j < C; j += blockDim.x) {
    const int q = j < half ? j : j - half;
    const float f = expf(-9.210340371976184f * (float)q / (float)half);
    float st, ct, sh, ch;
    sincosf(t * f, &st, &ct);
    sincosf(h * f, &sh, &ch);
    float v = j < half ? ct + ch : st + sh;
    if (latent_cond) v += latent_cond[b * C + j];
    cond[b * C + j] = __float2bfloat16(v);
  }
}

__global__ void to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(src[i]);
}
// x[b, j] += bias[j] * alpha   (what a split-K output projection's atomics then add the product to)
__global__ void add_scaled_bias_kernel(float* __restrict__ x, const float* __restrict__ bias, float alpha, int N, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] += bias[i % N] * alpha;
}
// dst[i] = bf16(gelu(src[i] + bias[i % N]))   (second half of a split-K projection whose activation needs the complete sum)
__global__ void bias_gelu_bf16_kernel(const float* __restrict__ src, const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst,
                                      int N, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(gelu_tanh(src[i] + bias[i % N]));
}
__global__ void copy_pair_kernel(const float* __restrict__ src, float* __restrict__ dst_f, __nv_bfloat16* __restrict__ dst_b,
                                 int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = src[i];
    dst_f[i] = v;
    dst_b[i] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------------------------------------ MLP-Mixer
// AdaLN over the channel axis of u[b, token, :] (LayerNorm eps 1e-6, no affine; (1+scale) n + shift with
// [scale|shift] = ss[b, :]).  TRANSPOSED: writes A[b, c, token] (token mixing operates on the transposed tensor,
// mlp_mixer.py:81-86); otherwise first adds the token-mix result T2[b, c, token] to u in place (skip, :87) and writes
// A[b, token, c] for the channel MLP (:90-92).  One thread per (b, token).
template <int CH, bool TRANSPOSED>
__global__ void __launch_bounds__(256) mixer_adaln_kernel(float* __restrict__ u, const float* __restrict__ T2,
                                                          const float* __restrict__ ss, __nv_bfloat16* __restrict__ A,
                                                          int tokens, int64_t B) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * tokens) return;
  const int64_t b = i / tokens;
  const int tok = (int)(i % tokens);
  float v[CH];
  float* up = u + i * CH;
#pragma unroll
  for (int c = 0; c < CH; c += 4) {
    const float4 q = *reinterpret_cast<const float4*>(up + c);
    v[c] = q.x; v[c + 1] = q.y; v[c + 2] = q.z; v[c + 3] = q.w;
  }
  if (!TRANSPOSED) {
#pragma unroll
    for (int c = 0; c < CH; ++c) v[c] += T2[(b * CH + c) * tokens + tok];
#pragma unroll
    for (int c = 0; c < CH; c += 4) *reinterpret_cast<float4*>(up + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
  }
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) { sum += v[c]; sq += v[c] * v[c]; }
  const float mu = sum * (1.0f / CH);
  const float rstd = rsqrtf(fmaxf(0.f, sq * (1.0f / CH) - mu * mu) + LN_EPS);
  const float* sc = ss + b * 2 * CH;
#pragma unroll
  for (int c = 0; c < CH; ++c) v[c] = (1.0f + sc[c]) * ((v[c] - mu) * rstd) + sc[CH + c];
  if (TRANSPOSED) {
#pragma unroll
    for (int c = 0; c < CH; ++c) A[(b * CH + c) * tokens + tok] = __float2bfloat16(v[c]);
  } else {
    uint32_t* ap = reinterpret_cast<uint32_t*>(A + i * CH);
#pragma unroll
    for (int c = 0; c < CH; c += 2) ap[c / 2] = pack_bf16(v[c], v[c + 1]);
  }
}

struct MixerPlan {
  float *x, *latc, *ss1, *ss2, *u, *t2;
  __nv_bfloat16 *xb, *cond, *latb, *a1, *h1, *a2, *h2, *u3, *w2t;
  void plan(Arena& ar, const MfacMixerDims& d, int64_t B) {
    const int64_t TC = (int64_t)d.tokens * d.channels;
    x = ar.take<float>(B * d.D);
    xb = ar.take<__nv_bfloat16>(B * d.D);
    cond = ar.take<__nv_bfloat16>(B * d.C);
    latc = ar.take<float>(B * d.C);
    latb = ar.take<__nv_bfloat16>(B * (int64_t)(d.latent_flat > 0 ? d.latent_flat : 8));
    ss1 = ar.take<float>(B * 2 * d.channels);
    ss2 = ar.take<float>(B * 2 * d.channels);
    u = ar.take<float>(B * TC);
    a1 = ar.take<__nv_bfloat16>(B * TC);
    h1 = ar.take<__nv_bfloat16>(B * (int64_t)d.channels * d.token_mix);
    t2 = ar.take<float>(B * TC);
    a2 = ar.take<__nv_bfloat16>(B * TC);
    // the channel-mix hidden tensor (4 MB per row at the reference widths) exists only where the fused kernel does not apply
    h2 = ar.take<__nv_bfloat16>(channel_mix_fused_ok(d.channels, d.channel_mix) && B * (int64_t)d.tokens <= 0x7fffffff
                                    ? 8 : B * (int64_t)d.tokens * d.channel_mix);
    u3 = ar.take<__nv_bfloat16>(B * TC);
    w2t = ar.take<__nv_bfloat16>((int64_t)d.channel_mix * d.channels);   // transposed second channel-mix kernel (fused path)
  }
};

bool mixer_dims_ok(const MfacMixerDims& d) {
  return d.D > 0 && d.C > 0 && d.C % 8 == 0 && d.nb > 0 && d.nb <= 64 && d.tokens > 0 && d.tokens % 8 == 0 &&
         (d.channels == 8 || d.channels == 16 || d.channels == 32) && d.token_mix % 8 == 0 && d.channel_mix % 8 == 0 &&
         d.D % 8 == 0 && d.latent_flat >= 0 && d.latent_flat % 8 == 0;
}

template <int CH>
int mixer_forward_impl(const MfacMixerDims& d, const MfacMixerWeights& w, const float* x, const float* time, const float* latents,
                       float* out, int64_t B, const MixerPlan& p, cudaStream_t s) {
  const int M = (int)B;
  const int TC = d.tokens * d.channels;
  copy_pair_kernel<<<nblk(B * d.D, 256), 256, 0, s>>>(x, p.x, p.xb, B * d.D);
  count_launch();
  if (latents) {
    to_bf16_kernel<<<nblk(B * d.latent_flat, 256), 256, 0, s>>>(latents, p.latb, B * d.latent_flat);
    count_launch();
    MFAC_OK(dense(p.latb, d.latent_flat, w.latent_proj, M, d.C, d.latent_flat, EpiLinearF32{w.latent_proj.b, p.latc, d.C}, s));
  }
  flow_cond_kernel<<<(unsigned)B, 128, 0, s>>>(time, latents ? p.latc : nullptr, p.cond, d.C);
  count_launch();
  const float inv_nb = 1.0f / (float)d.nb;
  for (int k = 0; k < d.nb; ++k) {
    const MfacMixerBlockW& bw = w.blocks[k];
    // input_proj: [B, D] -> u[B, tokens, CH]                                     (mlp_mixer.py:147-149)
    MFAC_OK(dense(p.xb, d.D, bw.input_proj, M, TC, d.D, EpiLinearF32{bw.input_proj.b, p.u, TC}, s));
    // token mixing                                                               (:77-87)
    MFAC_OK(dense(p.cond, d.C, bw.adaln1, M, 2 * CH, d.C, EpiLinearF32{bw.adaln1.b, p.ss1, 2 * CH}, s));
    mixer_adaln_kernel<CH, true><<<nblk(B * d.tokens, 256), 256, 0, s>>>(p.u, nullptr, p.ss1, p.a1, d.tokens, B);
    count_launch();
    MFAC_OK(dense(p.a1, d.tokens, bw.tok1, M * CH, d.token_mix, d.tokens,
                  EpiBiasGelu{bw.tok1.b, p.h1, nullptr, d.token_mix}, s));
    MFAC_OK(dense(p.h1, d.token_mix, bw.tok2, M * CH, d.tokens, d.token_mix, EpiLinearF32{bw.tok2.b, p.t2, d.tokens}, s));
    // channel mixing                                                             (:89-93)
    MFAC_OK(dense(p.cond, d.C, bw.adaln2, M, 2 * CH, d.C, EpiLinearF32{bw.adaln2.b, p.ss2, 2 * CH}, s));
    mixer_adaln_kernel<CH, false><<<nblk(B * d.tokens, 256), 256, 0, s>>>(p.u, p.t2, p.ss2, p.a2, d.tokens, B);
    count_launch();
    // both channel-mix layers in ONE kernel (the [B * tokens, channel_mix] hidden tensor stays on the SM); other geometries
    // run the two GEMMs through HBM
    const int fused = channel_mix_fused(p.a2, bw.ch1, bw.ch2, p.w2t, p.u, p.u3, (int64_t)M * d.tokens, CH, d.channel_mix, s);
    if (fused == MFAC_ERR_UNSUPPORTED) {
      MFAC_OK(dense(p.a2, CH, bw.ch1, M * d.tokens, d.channel_mix, CH, EpiBiasGelu{bw.ch1.b, p.h2, nullptr, d.channel_mix}, s));
      MFAC_OK(dense(p.h2, d.channel_mix, bw.ch2, M * d.tokens, CH, d.channel_mix,
                    EpiAffineResidual{bw.ch2.b, p.u, nullptr, p.u3, CH, 1.0f}, s));
    } else {
      MFAC_OK(fused);
    }
    // output_proj, x / num_blocks + residual                                     (:157-163)
    if (ceil_div(M, GEMM_BM) * ceil_div(d.D, 128) * 4 <= num_sms() && TC >= 2048 && d.D % 4 == 0) {
      // few rows: K = tokens * CH sliced over the machine, partial products reduced into x by fp32 atomics (79 -> 15 us at B = 256)
      add_scaled_bias_kernel<<<nblk(B * d.D, 256), 256, 0, s>>>(p.x, bw.output_proj.b, inv_nb, d.D, B * d.D);
      count_launch();
      MFAC_OK((launch_gemm<false, true>(GemmOperandDesc{p.u3, TC, false}, GemmOperandDesc{bw.output_proj.w, d.D, true}, M, d.D, TC,
                                        EpiScaledAtomicAdd{p.x, d.D, inv_nb}, s, 0, /*split_k=*/true)));
      to_bf16_kernel<<<nblk(B * d.D, 256), 256, 0, s>>>(p.x, p.xb, B * d.D);
      count_launch();
    } else {
      MFAC_OK(dense(p.u3, TC, bw.output_proj, M, d.D, TC, EpiAffineResidual{bw.output_proj.b, p.x, p.x, p.xb, d.D, inv_nb}, s));
    }
  }
  MFAC_CUDA_OK(cudaMemcpyAsync(out, p.x, (size_t)B * d.D * 4, cudaMemcpyDeviceToDevice, s));
  return launch_status();
}

// ------------------------------------------------------------------------------------------------ ConvNeXt
// One CTA per sample: xs[S,S,CH] (input_proj2 output) -> LN(CH) -> FiLM -> ConvNeXt block -> bf16 [S*S*CH].
//   ConvNeXt: 3x3 SAME conv + bias -> LN(CH, eps 1e-6) -> 1x1 (CH -> 2CH) + bias -> GELU(tanh) -> GRN -> 1x1 (2CH -> CH)
//             + bias -> * layer_scale_gamma -> + skip                                (conv_flow.py:69-115)
//   GRN: gx[c] = sqrt(sum_{h,w} v^2); gx /= mean_c(gx) + 1e-6; v * (gamma + gx) + beta   (conv_flow.py:33-45)
// The image never leaves shared memory: padded input plane, the 2CH-wide hidden activations, all weights.
template <int CH>
__global__ void __launch_bounds__(256) convnext_block_kernel(const float* __restrict__ xs, const float* __restrict__ film,
                                                             MfacConvBlockW w, __nv_bfloat16* __restrict__ out, int S) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int P = S + 2;
  float* sIn = reinterpret_cast<float*>(smem_raw);             // [P][P][CH], zero halo
  float* sH = sIn + P * P * CH;                                // [S*S][2CH]
  float* sW3 = sH + S * S * 2 * CH;                            // [9][CH][CH]  (kh, kw, in, out)
  float* sW1 = sW3 + 9 * CH * CH;                              // [CH][2CH]
  float* sW2 = sW1 + CH * 2 * CH;                              // [2CH][CH]
  float* sB3 = sW2 + 2 * CH * CH;                              // [CH]
  float* sB1 = sB3 + CH;                                       // [2CH]
  float* sB2 = sB1 + 2 * CH;                                   // [CH]
  float* sLs = sB2 + CH;                                       // [CH]
  float* sG = sLs + CH;                                        // [2CH] GRN sums -> (gamma + gx)
  float* sBeta = sG + 2 * CH;                                  // [2CH]
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  const int npix = S * S;

  for (int i = tid; i < P * P * CH; i += blockDim.x) sIn[i] = 0.f;
  for (int i = tid; i < 9 * CH * CH; i += blockDim.x) sW3[i] = w.conv3_w[i];
  for (int i = tid; i < 2 * CH * CH; i += blockDim.x) { sW1[i] = w.pw1_w[i]; sW2[i] = w.pw2_w[i]; }
  if (tid < CH) { sB3[tid] = w.conv3_b[tid]; sB2[tid] = w.pw2_b[tid]; sLs[tid] = w.layer_scale ? w.layer_scale[tid] : 1.0f; }
  if (tid < 2 * CH) { sB1[tid] = w.pw1_b[tid]; sG[tid] = 0.f; sBeta[tid] = w.grn_beta ? w.grn_beta[tid] : 0.f; }
  __syncthreads();

  // LN over channels + FiLM into the padded plane                                  (conv_flow.py:176-187)
  const float* fb = film + b * 2 * CH;
  for (int pix = tid; pix < npix; pix += blockDim.x) {
    const float* src = xs + (b * npix + pix) * CH;
    float v[CH];
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int c = 0; c < CH; c += 4) {
      const float4 q = *reinterpret_cast<const float4*>(src + c);
      v[c] = q.x; v[c + 1] = q.y; v[c + 2] = q.z; v[c + 3] = q.w;
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) { sum += v[c]; sq += v[c] * v[c]; }
    const float mu = sum * (1.0f / CH);
    const float rstd = rsqrtf(fmaxf(0.f, sq * (1.0f / CH) - mu * mu) + LN_EPS);
    float* dst = sIn + (((pix / S) + 1) * P + (pix % S) + 1) * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) dst[c] = (1.0f + fb[c]) * ((v[c] - mu) * rstd) + fb[CH + c];
  }
  __syncthreads();

  // pass 1: 3x3 conv -> LN -> 1x1 expand -> GELU, hidden activations to sH, per-channel sum of squares for GRN
  float gsq[2 * CH];
#pragma unroll
  for (int c = 0; c < 2 * CH; ++c) gsq[c] = 0.f;
  for (int pix = tid; pix < npix; pix += blockDim.x) {
    const int py = pix / S, px = pix % S;
    float acc[CH];
#pragma unroll
    for (int o = 0; o < CH; ++o) acc[o] = sB3[o];
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const float* in = sIn + ((py + tap / 3) * P + px + tap % 3) * CH;
      const float* wt = sW3 + tap * CH * CH;
#pragma unroll 4
      for (int ci = 0; ci < CH; ++ci) {
        const float xv = in[ci];
#pragma unroll
        for (int o = 0; o < CH; o += 4) {
          const float4 q = *reinterpret_cast<const float4*>(wt + ci * CH + o);
          acc[o] = fmaf(xv, q.x, acc[o]); acc[o + 1] = fmaf(xv, q.y, acc[o + 1]);
          acc[o + 2] = fmaf(xv, q.z, acc[o + 2]); acc[o + 3] = fmaf(xv, q.w, acc[o + 3]);
        }
      }
    }
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) { sum += acc[c]; sq += acc[c] * acc[c]; }
    const float mu = sum * (1.0f / CH);
    const float rstd = rsqrtf(fmaxf(0.f, sq * (1.0f / CH) - mu * mu) + LN_EPS);
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = (acc[c] - mu) * rstd;
    float* hp = sH + pix * 2 * CH;
#pragma unroll
    for (int o = 0; o < 2 * CH; o += 4) {
      float4 hv = *reinterpret_cast<const float4*>(sB1 + o);
#pragma unroll
      for (int ci = 0; ci < CH; ++ci) {
        const float4 q = *reinterpret_cast<const float4*>(sW1 + ci * 2 * CH + o);
        hv.x = fmaf(acc[ci], q.x, hv.x); hv.y = fmaf(acc[ci], q.y, hv.y);
        hv.z = fmaf(acc[ci], q.z, hv.z); hv.w = fmaf(acc[ci], q.w, hv.w);
      }
      hv.x = gelu_tanh(hv.x); hv.y = gelu_tanh(hv.y); hv.z = gelu_tanh(hv.z); hv.w = gelu_tanh(hv.w);
      *reinterpret_cast<float4*>(hp + o) = hv;
      gsq[o] += hv.x * hv.x; gsq[o + 1] += hv.y * hv.y; gsq[o + 2] += hv.z * hv.z; gsq[o + 3] += hv.w * hv.w;
    }
  }
#pragma unroll
  for (int c = 0; c < 2 * CH; ++c) {
    const float t = warp_sum(gsq[c]);
    if ((tid & 31) == 0) atomicAdd(&sG[c], t);
  }
  __syncthreads();
  if (tid == 0) {
    float mean = 0.f;
    for (int c = 0; c < 2 * CH; ++c) { sG[c] = sqrtf(sG[c]); mean += sG[c]; }
    mean = mean / (float)(2 * CH) + 1e-6f;
    for (int c = 0; c < 2 * CH; ++c) sG[c] = (w.grn_gamma ? w.grn_gamma[c] : 0.f) + sG[c] / mean;
  }
  __syncthreads();

  // pass 2: GRN -> 1x1 contract -> layer scale -> + skip -> bf16
  for (int pix = tid; pix < npix; pix += blockDim.x) {
    const float* hp = sH + pix * 2 * CH;
    float acc[CH];
#pragma unroll
    for (int o = 0; o < CH; ++o) acc[o] = sB2[o];
#pragma unroll 4
    for (int ci = 0; ci < 2 * CH; ++ci) {
      const float hv = hp[ci] * sG[ci] + sBeta[ci];
#pragma unroll
      for (int o = 0; o < CH; o += 4) {
        const float4 q = *reinterpret_cast<const float4*>(sW2 + ci * CH + o);
        acc[o] = fmaf(hv, q.x, acc[o]); acc[o + 1] = fmaf(hv, q.y, acc[o + 1]);
        acc[o + 2] = fmaf(hv, q.z, acc[o + 2]); acc[o + 3] = fmaf(hv, q.w, acc[o + 3]);
      }
    }
    const float* skip = sIn + (((pix / S) + 1) * P + (pix % S) + 1) * CH;
    uint32_t* op = reinterpret_cast<uint32_t*>(out + (b * npix + pix) * CH);
#pragma unroll
    for (int o = 0; o < CH; o += 2)
      op[o / 2] = pack_bf16(acc[o] * sLs[o] + skip[o], acc[o + 1] * sLs[o + 1] + skip[o + 1]);
  }
}

template <int CH>
size_t convnext_smem(int S) {
  const int P = S + 2;
  return sizeof(float) * ((size_t)P * P * CH + (size_t)S * S * 2 * CH + 9 * CH * CH + 4 * CH * CH + 3 * CH + 3 * 2 * CH);
}

struct ConvPlan {
  float *x, *latc, *film, *xs, *q1f;
  __nv_bfloat16 *xb, *cond, *latb, *p1, *xf, *q1;
  void plan(Arena& ar, const MfacConvDims& d, int64_t B) {
    const int64_t SC = (int64_t)d.S * d.S * d.channels;
    x = ar.take<float>(B * d.D);
    xb = ar.take<__nv_bfloat16>(B * d.D);
    cond = ar.take<__nv_bfloat16>(B * d.C);
    latc = ar.take<float>(B * d.C);
    latb = ar.take<__nv_bfloat16>(B * (int64_t)(d.latent_flat > 0 ? d.latent_flat : 8));
    film = ar.take<float>(B * 2 * d.channels);
    p1 = ar.take<__nv_bfloat16>(B * d.bottleneck);
    xs = ar.take<float>(B * SC);
    xf = ar.take<__nv_bfloat16>(B * SC);
    q1 = ar.take<__nv_bfloat16>(B * d.bottleneck);
    q1f = ar.take<float>(B * d.bottleneck);   // split-K partial sums of output_proj1
  }
};

bool conv_dims_ok(const MfacConvDims& d) {
  return d.D > 0 && d.D % 8 == 0 && d.C > 0 && d.C % 8 == 0 && d.nb > 0 && d.nb <= 64 && d.S > 0 && d.S <= 40 &&
         (d.channels == 4 || d.channels == 8 || d.channels == 16) && d.bottleneck > 0 && d.bottleneck % 8 == 0 &&
         d.latent_flat >= 0 && d.latent_flat % 8 == 0 && ((int64_t)d.S * d.S * d.channels) % 8 == 0;
}

// tensor-core block kernel for the reference geometry (16 channels, S = 16 or 32); MFAC_ERR_UNSUPPORTED for the other
// geometries, which keep the fp32 kernel
template <int CH>
int conv_block_mma(int S, int64_t B, const float* xs, const float* film, const MfacConvBlockW& bw, __nv_bfloat16* xf, cudaStream_t s) {
  if constexpr (CH == 16) {
    static const bool off = getenv("MFAC_NO_CONV_MMA") != nullptr;
    if (off || (S != 32 && S != 16)) return MFAC_ERR_UNSUPPORTED;
    static PerDeviceOnce configured;
    if (configured.need()) {
      MFAC_CUDA_OK(cudaFuncSetAttribute(convnext_block_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)convnext_mma_smem<32>()));
      MFAC_CUDA_OK(cudaFuncSetAttribute(convnext_block_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)convnext_mma_smem<16>()));
      configured.done();
    }
    if (S == 32) convnext_block_mma_kernel<32><<<(unsigned)B, 256, convnext_mma_smem<32>(), s>>>(xs, film, bw, xf);
    else convnext_block_mma_kernel<16><<<(unsigned)B, 256, convnext_mma_smem<16>(), s>>>(xs, film, bw, xf);
    return launch_status();
  } else {
    return MFAC_ERR_UNSUPPORTED;
  }
}

template <int CH>
int conv_forward_impl(const MfacConvDims& d, const MfacConvWeights& w, const float* x, const float* time, const float* latents,
                      float* out, int64_t B, const ConvPlan& p, cudaStream_t s) {
  const int M = (int)B;
  const int SC = d.S * d.S * CH;
  const size_t smem = convnext_smem<CH>(d.S);
  if (smem > 227 * 1024) return MFAC_ERR_UNSUPPORTED;
  static PerDeviceOnce configured;
  if (configured.need()) {
    MFAC_CUDA_OK(cudaFuncSetAttribute(convnext_block_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.done();
  }
  copy_pair_kernel<<<nblk(B * d.D, 256), 256, 0, s>>>(x, p.x, p.xb, B * d.D);
  count_launch();
  if (latents) {
    to_bf16_kernel<<<nblk(B * d.latent_flat, 256), 256, 0, s>>>(latents, p.latb, B * d.latent_flat);
    count_launch();
    MFAC_OK(dense(p.latb, d.latent_flat, w.latent_proj, M, d.C, d.latent_flat, EpiLinearF32{w.latent_proj.b, p.latc, d.C}, s));
  }
  flow_cond_kernel<<<(unsigned)B, 128, 0, s>>>(time, latents ? p.latc : nullptr, p.cond, d.C);
  count_launch();
  const float inv_nb = 1.0f / (float)d.nb;
  for (int k = 0; k < d.nb; ++k) {
    const MfacConvBlockW& bw = w.blocks[k];
    // input projection through the bottleneck                                     (conv_flow.py:166-174)
    MFAC_OK(dense(p.xb, d.D, bw.input_proj1, M, d.bottleneck, d.D, EpiBiasGelu{bw.input_proj1.b, p.p1, nullptr, d.bottleneck}, s));
    MFAC_OK(dense(p.p1, d.bottleneck, bw.input_proj2, M, SC, d.bottleneck, EpiLinearF32{bw.input_proj2.b, p.xs, SC}, s));
    // FiLM parameters                                                             (:180-181)
    MFAC_OK(dense(p.cond, d.C, bw.conditioning, M, 2 * CH, d.C, EpiLinearF32{bw.conditioning.b, p.film, 2 * CH}, s));
    const int mma = conv_block_mma<CH>(d.S, B, p.xs, p.film, bw, p.xf, s);
    if (mma == MFAC_ERR_UNSUPPORTED) convnext_block_kernel<CH><<<(unsigned)B, 256, smem, s>>>(p.xs, p.film, bw, p.xf, d.S);
    else MFAC_OK(mma);
    count_launch();
    // output projection, x / num_blocks + residual                                (:195-205)
    if (ceil_div(M, GEMM_BM) * ceil_div(d.bottleneck, 128) * 4 <= num_sms() && SC >= 2048 && d.bottleneck % 4 == 0) {
      // [B, S*S*CH] x [S*S*CH, 128]: one column of tiles unless K is sliced; the GELU needs the complete sum, so the slices reduce
      // into an fp32 scratch and a row kernel applies bias + GELU
      MFAC_CUDA_OK(cudaMemsetAsync(p.q1f, 0, (size_t)B * d.bottleneck * 4, s));
      MFAC_OK((launch_gemm<false, true>(GemmOperandDesc{p.xf, SC, false}, GemmOperandDesc{bw.output_proj1.w, d.bottleneck, true}, M,
                                        d.bottleneck, SC, EpiScaledAtomicAdd{p.q1f, d.bottleneck, 1.0f}, s, 0, /*split_k=*/true)));
      bias_gelu_bf16_kernel<<<nblk(B * d.bottleneck, 256), 256, 0, s>>>(p.q1f, bw.output_proj1.b, p.q1, d.bottleneck, B * d.bottleneck);
      count_launch();
    } else {
      MFAC_OK(dense(p.xf, SC, bw.output_proj1, M, d.bottleneck, SC, EpiBiasGelu{bw.output_proj1.b, p.q1, nullptr, d.bottleneck}, s));
    }
    MFAC_OK(dense(p.q1, d.bottleneck, bw.output_proj2, M, d.D, d.bottleneck,
                  EpiAffineResidual{bw.output_proj2.b, p.x, p.x, p.xb, d.D, inv_nb}, s));
  }
  MFAC_CUDA_OK(cudaMemcpyAsync(out, p.x, (size_t)B * d.D * 4, cudaMemcpyDeviceToDevice, s));
  return launch_status();
}

}  // namespace
}  // namespace mfac

using namespace mfac;

extern "C" {

size_t mfac_mixer_workspace_bytes(const MfacMixerDims* d, int64_t B) {
  if (!d || !mixer_dims_ok(*d) || B <= 0) return 0;
  Arena ar(nullptr, 0);
  MixerPlan p;
  p.plan(ar, *d, B);
  return ar.off + 256;
}

int mfac_mixer_forward(const MfacMixerDims* d, const MfacMixerWeights* w, const float* x, const float* time, const float* latents,
                       float* out, int64_t B, void* ws, size_t ws_bytes, void* stream) {
  if (!d || !w || !w->blocks || !x || !time || !out) return MFAC_ERR_NULL;
  if (!mixer_dims_ok(*d)) return MFAC_ERR_UNSUPPORTED;
  if (B <= 0 || B * (int64_t)d->tokens > 0x7fffffff) return MFAC_ERR_BAD_SHAPE;
  if (latents && d->latent_flat <= 0) return MFAC_ERR_BAD_SHAPE;
  if (!ws) return MFAC_ERR_WORKSPACE;
  Arena ar(ws, ws_bytes);
  MixerPlan p;
  p.plan(ar, *d, B);
  if (ar.overflow) return MFAC_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  sweep_reset();
  switch (d->channels) {
    case 8: return mixer_forward_impl<8>(*d, *w, x, time, latents, out, B, p, s);
    case 16: return mixer_forward_impl<16>(*d, *w, x, time, latents, out, B, p, s);
    case 32: return mixer_forward_impl<32>(*d, *w, x, time, latents, out, B, p, s);
  }
  return MFAC_ERR_UNSUPPORTED;
}

size_t mfac_conv_workspace_bytes(const MfacConvDims* d, int64_t B) {
  if (!d || !conv_dims_ok(*d) || B <= 0) return 0;
  Arena ar(nullptr, 0);
  ConvPlan p;
  p.plan(ar, *d, B);
  return ar.off + 256;
}

int mfac_conv_forward(const MfacConvDims* d, const MfacConvWeights* w, const float* x, const float* time, const float* latents,
                      float* out, int64_t B, void* ws, size_t ws_bytes, void* stream) {
  if (!d || !w || !w->blocks || !x || !time || !out) return MFAC_ERR_NULL;
  if (!conv_dims_ok(*d)) return MFAC_ERR_UNSUPPORTED;
  if (B <= 0 || B > 0x7fffffff) return MFAC_ERR_BAD_SHAPE;
  if (latents && d->latent_flat <= 0) return MFAC_ERR_BAD_SHAPE;
  if (!ws) return MFAC_ERR_WORKSPACE;
  Arena ar(ws, ws_bytes);
  ConvPlan p;
  p.plan(ar, *d, B);
  if (ar.overflow) return MFAC_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  sweep_reset();
  switch (d->channels) {
    case 4: return conv_forward_impl<4>(*d, *w, x, time, latents, out, B, p, s);
    case 8: return conv_forward_impl<8>(*d, *w, x, time, latents, out, B, p, s);
    case 16: return conv_forward_impl<16>(*d, *w, x, time, latents, out, B, p, s);
  }
  return MFAC_ERR_UNSUPPORTED;
}

}  // extern "C"
