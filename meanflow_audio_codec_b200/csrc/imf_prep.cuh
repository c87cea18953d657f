// Device-side pieces of the iMF prologue shared by imf_prep_kernel (imf_kernels.cuh) and the fused tokenise + prologue kernel
// (mdct.cu): Philox draws, the (t, r) rule, the sinusoidal conditioning rows.  Inline device functions and plain structs only
// (this header is included by more than one translation unit).
//
// Reference semantics (paths inside /root/reference/meanflow_audio_codec/): time embedding utils.py:5-13, (t, r) sampling
// utils.py:32-45, time_sampling.py:44-135.
#pragma once

#include "imf_layout.cuh"

namespace mfac {

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (own stream; the reference's threefry stream is not imitated,
// parity runs pass e/t/r explicitly -- SURVEY.md R6)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f); }
// 4 x N(0,1) for counter (idx, stream, step)
__device__ __forceinline__ float4 philox_normal4(uint64_t idx, uint32_t stream, uint64_t seed, uint64_t step) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), stream ^ (uint32_t)(step << 8), (uint32_t)(step >> 24)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float r0 = sqrtf(-2.0f * logf(u01(r.x))), r1 = sqrtf(-2.0f * logf(u01(r.z)));
  float s0, c0, s1, c1;
  sincosf(6.283185307179586f * u01(r.y), &s0, &c0);
  sincosf(6.283185307179586f * u01(r.w), &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// cond[j] = cos(t f_j) + cos(h f_j), cond[half+j] = sin(t f_j) + sin(h f_j); optional d/ds along (tdot=1,hdot=1)
__device__ __forceinline__ void write_cond_row(float t, float h, int C, int Cp, __nv_bfloat16* cond, __nv_bfloat16* dcond) {
  const int half = C / 2;
  for (int j = threadIdx.x; j < Cp; j += blockDim.x) {
    float v = 0.f, dv = 0.f;
    if (j < C) {
      const int q = j < half ? j : j - half;
      const float f = expf(-9.210340371976184f * (float)q / (float)half);
      float st, ct, sh, ch;
      sincosf(t * f, &st, &ct);
      sincosf(h * f, &sh, &ch);
      if (j < half) { v = ct + ch; dv = -f * (st + sh); }
      else { v = st + sh; dv = f * (ct + ch); }
    }
    cond[j] = __float2bfloat16(v);
    if (dcond) dcond[j] = __float2bfloat16(dv);
  }
}

// z_t = (1 - t) x + (noise_min + noise_max t) e  and  target = noise_max e - x  (noise_schedules.py:79-88) with the rounding
// spelled out, so that every kernel that forms them (imf_prep_kernel, the fused tokenise + prologue kernel) gives the same bits
__device__ __forceinline__ float zt_of(float omt, float x, float nscale, float e) { return __fmaf_rn(omt, x, __fmul_rn(nscale, e)); }
__device__ __forceinline__ float target_of(float nmax, float e, float x) { return __fmaf_rn(nmax, e, -x); }

struct PrepArgs {
  const float* x;      // [B, D]
  const float* e_in;   // [B, D] or null
  const float* t_in;   // [B] or null
  const float* r_in;   // [B] or null
  float* e;            // [B, Dp] or null  the noise actually used (parity / statistics tests ask for it)
  float* target;       // [B, Dp]  noise_max e - x: what the loss compares v_pred with (noise_schedules.py:88)
  float* z;            // [B, Dp] or null  z_t for the v pass (improved mean flow only), updated in place by it
  float* z2;           // [B, Dp]  z_t for the u pass
  float* seed;         // [B, Dp] or null  tangent seed noise_max e - x (mean flow: the JVP runs along the true velocity)
  __nv_bfloat16* xb;   // [B, Dp]
  float* t;            // [B]
  float* r;            // [B]
  __nv_bfloat16 *cond_v, *cond_u, *dcond_u;  // [B, Cp]
  MfacImfConfig cfg;
  int64_t B;
};

// (t, r) of row b: copied, or drawn like utils.py:36-45 (sample_tr) / time_sampling.py:44-75 from the Philox stream
__device__ __forceinline__ void draw_tr(const PrepArgs& a, int64_t b, uint64_t step, float& t, float& r) {
  if (a.t_in) {
    t = a.t_in[b];
    r = a.r_in[b];
  } else {
    const float4 n4 = philox_normal4(a.cfg.row_offset + b, 1u, a.cfg.seed, step);
    float lt = 1.0f / (1.0f + expf(-(n4.x * a.cfg.time_std + a.cfg.time_mean)));
    const float lr = 1.0f / (1.0f + expf(-(n4.y * a.cfg.time_std + a.cfg.time_mean)));
    if (a.cfg.uniform_time) lt = 0.5f * (1.0f + erff(n4.x * 0.70710678118654752f));  // Phi(N(0,1)) ~ U(0,1)
    t = fmaxf(lt, lr);
    r = fminf(lt, lr);
    if (b < (int64_t)((float)a.B * a.cfg.data_proportion)) r = t;  // utils.py:41-44, per local shard
    if (a.cfg.method == MFAC_LOSS_FLOW_MATCHING) t = lt;           // a single time (time_sampling.py:44-75)
  }
  if (a.cfg.method == MFAC_LOSS_FLOW_MATCHING) r = t;               // h = 0 (loss_strategies.py:88)
}

}  // namespace mfac
