// Row kernels and fused GEMM epilogues of the iMF step (everything that is not a tcgen05 MMA).
//
// Reference semantics (paths inside /root/reference/meanflow_audio_codec/):
//   time embedding        utils.py:5-13
//   (t, r) sampling       utils.py:32-45
//   z_t / target          trainers/noise_schedules.py:69-88
//   block (AdaLN + MLP)   models/mlp_flow.py:83-117
//   tangent recurrences   jax.jvp at trainers/loss_strategies.py:263-267
//   loss                  utils.py:16-25, loss_strategies.py:270-277
//   backward recurrences  jax.value_and_grad at loss_strategies.py:279
#pragma once

#include "imf_layout.cuh"
#include "imf_prep.cuh"
#include "epilogues.cuh"

namespace mfac {

constexpr int ROW_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red /*[32]*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// ---------------------------------------------------------------------------------------
// iMF prologue: one block per row.  e, t, r drawn or copied; z_t; bf16 copy of x; three cond rows.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS) imf_prep_kernel(PrepArgs a, Dims d) {
  MFAC_PDL_SYNC();
  const int64_t b = blockIdx.x;
  const uint64_t step = a.cfg.step_dev ? *a.cfg.step_dev : a.cfg.step;
  __shared__ float s_tr[2];
  if (threadIdx.x == 0) {
    float t, r;
    draw_tr(a, b, step, t, r);
    s_tr[0] = t; s_tr[1] = r;
    a.t[b] = t; a.r[b] = r;
  }
  __syncthreads();
  const float t = s_tr[0], r = s_tr[1];
  const float nscale = a.cfg.noise_min + a.cfg.noise_max * t;
  const bool vec = d.D == d.Dp && (d.D & 3) == 0;   // unpadded rows: 16-byte accesses (row bases are then 16-byte aligned)
  for (int j4 = threadIdx.x * 4; j4 < d.Dp; j4 += blockDim.x * 4) {
    float ev[4] = {0.f, 0.f, 0.f, 0.f};
    if (!a.e_in && j4 < d.D) {
      const uint64_t idx = ((a.cfg.row_offset + (uint64_t)b) * (uint64_t)d.Dp + (uint64_t)j4) >> 2;
      const float4 n4 = philox_normal4(idx, 0u, a.cfg.seed, step);
      ev[0] = n4.x; ev[1] = n4.y; ev[2] = n4.z; ev[3] = n4.w;
    }
    if (vec) {
      const int64_t at = b * d.Dp + j4;
      const float4 xv = *reinterpret_cast<const float4*>(a.x + at);
      const float4 e = a.e_in ? *reinterpret_cast<const float4*>(a.e_in + at) : make_float4(ev[0], ev[1], ev[2], ev[3]);
      const float omt = 1.0f - t;
      const float4 zt = make_float4(zt_of(omt, xv.x, nscale, e.x), zt_of(omt, xv.y, nscale, e.y), zt_of(omt, xv.z, nscale, e.z),
                                    zt_of(omt, xv.w, nscale, e.w));
      if (a.e) *reinterpret_cast<float4*>(a.e + at) = e;
      if (a.z) *reinterpret_cast<float4*>(a.z + at) = zt;
      *reinterpret_cast<float4*>(a.z2 + at) = zt;
      const float nmax = a.cfg.noise_max;
      const float4 tg = make_float4(target_of(nmax, e.x, xv.x), target_of(nmax, e.y, xv.y), target_of(nmax, e.z, xv.z),
                                    target_of(nmax, e.w, xv.w));
      *reinterpret_cast<float4*>(a.target + at) = tg;
      if (a.seed) *reinterpret_cast<float4*>(a.seed + at) = tg;
      *reinterpret_cast<uint2*>(a.xb + at) = make_uint2(pack_bf16(xv.x, xv.y), pack_bf16(xv.z, xv.w));
      continue;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j4 + q;
      float xv = 0.f, e = 0.f;
      if (j < d.D) {
        xv = a.x[b * d.D + j];
        e = a.e_in ? a.e_in[b * d.D + j] : ev[q];
      }
      if (a.e) a.e[b * d.Dp + j] = e;
      const float zt = zt_of(1.0f - t, xv, nscale, e);
      if (a.z) a.z[b * d.Dp + j] = zt;
      a.z2[b * d.Dp + j] = zt;
      a.target[b * d.Dp + j] = target_of(a.cfg.noise_max, e, xv);
      if (a.seed) a.seed[b * d.Dp + j] = target_of(a.cfg.noise_max, e, xv);
      a.xb[b * d.Dp + j] = __float2bfloat16(xv);
    }
  }
  if (a.cond_v) write_cond_row(t, 0.f, d.C, d.Cp, a.cond_v + b * d.Cp, nullptr);
  write_cond_row(t, t - r, d.C, d.Cp, a.cond_u + b * d.Cp, a.dcond_u + b * d.Cp);
}

// time[B,2] -> cond[B,Cp]   (mfac_mlp_forward / samplers)
__global__ void __launch_bounds__(128) cond_from_time_kernel(const float* time, __nv_bfloat16* cond, Dims d) {
  const int64_t b = blockIdx.x;
  write_cond_row(time[2 * b], time[2 * b + 1], d.C, d.Cp, cond + b * d.Cp, nullptr);
}
// constant (t, h) for every row (samplers)
__global__ void __launch_bounds__(128) cond_const_kernel(float t, float h, __nv_bfloat16* cond, Dims d) {
  const int64_t b = blockIdx.x;
  write_cond_row(t, h, d.C, d.Cp, cond + b * d.Cp, nullptr);
}

// fp32 [B, n] -> padded fp32 [B, np] and/or bf16 [B, np]
__global__ void pad_rows_kernel(const float* src, int n, float* dst_f, __nv_bfloat16* dst_b, int np, int64_t B) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * np) return;
  const int64_t b = i / np;
  const int j = (int)(i % np);
  const float v = j < n ? src[b * n + j] : 0.f;
  if (dst_f) dst_f[i] = v;
  if (dst_b) dst_b[i] = __float2bfloat16(v);
}
__global__ void unpad_rows_kernel(const float* src, int np, float* dst, int n, int64_t B) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n) return;
  const int64_t b = i / n;
  const int j = (int)(i % n);
  dst[i] = src[b * np + j];
}

// ---------------------------------------------------------------------------------------
// AdaLN: n = LN([lat|x]) (no affine, eps 1e-6, var = max(0, E[c^2]-mu^2)); hin = (1+s1) n + shift.
// TANGENT adds cdot = [0|xdot]: ndot = (cdot - mean(cdot) - n mean(n cdot)) rstd;
//                               hindot = s1dot n + (1+s1) ndot + shiftdot.
// One block per row; the row is staged in shared memory (any Ip).
// ---------------------------------------------------------------------------------------
struct LnModArgs {
  const float* lat;          // [B, Lp] or null (zeros)
  const float* x;            // [B, Dp]
  const __nv_bfloat16* m;    // [B, Mp]
  __nv_bfloat16* hin;        // [B, Ip]; may be null for TANGENT (tangent-only pass: hin / mu / rstd come from the primal pass)
  const float* xd;           // tangent
  const __nv_bfloat16* md;   // tangent
  __nv_bfloat16* hind;       // tangent
  float* mu;                 // [B] or null
  float* rstd;               // [B] or null
  int64_t m_stride;          // row stride of m in elements: Mp, or 0 when every row shares one modulation row
                             // (samplers: the modulation depends on (t, h) only, which is constant over the batch)
  int reverse;               // vectorised kernels: 1 = CTAs walk the rows from the last to the first (gemm.cuh: sweep_next)
};

template <bool TANGENT>
__global__ void __launch_bounds__(ROW_THREADS) lnmod_kernel(LnModArgs a, Dims d) {
  extern __shared__ float s_row[];  // c[Ip] (+ cd[Ip])
  __shared__ float red[32];
  float* s_c = s_row;
  float* s_cd = s_row + d.Ip;
  const int64_t b = blockIdx.x;
  float sum = 0.f, sq = 0.f, sumd = 0.f;
  for (int p = threadIdx.x; p < d.Ip; p += blockDim.x) {
    float c = 0.f, cd = 0.f;
    if (p < d.Lp) {
      if (a.lat && p < d.L) c = a.lat[b * d.Lp + p];
    } else if (p - d.Lp < d.D) {
      c = a.x[b * d.Dp + (p - d.Lp)];
      if (TANGENT) cd = a.xd[b * d.Dp + (p - d.Lp)];
    }
    s_c[p] = c;
    sum += c;
    sq += c * c;
    if (TANGENT) { s_cd[p] = cd; sumd += cd; }
  }
  const float inv_i = 1.0f / (float)d.I;
  const float mu = block_sum(sum, red) * inv_i;
  const float ex2 = block_sum(sq, red) * inv_i;
  const float rstd = rsqrtf(fmaxf(0.f, ex2 - mu * mu) + LN_EPS);
  float mean_cd = 0.f, mean_ncd = 0.f;
  if (TANGENT) {
    mean_cd = block_sum(sumd, red) * inv_i;
    float acc = 0.f;
    for (int p = threadIdx.x; p < d.Ip; p += blockDim.x) {
      const bool real = p < d.L || (p >= d.Lp && p - d.Lp < d.D);
      if (real) acc += (s_c[p] - mu) * rstd * s_cd[p];
    }
    mean_ncd = block_sum(acc, red) * inv_i;
  }
  if (threadIdx.x == 0 && a.mu) { a.mu[b] = mu; a.rstd[b] = rstd; }
  const __nv_bfloat16* mrow = a.m + b * a.m_stride;
  const __nv_bfloat16* mdrow = TANGENT ? a.md + b * d.Mp : nullptr;
  for (int p = threadIdx.x; p < d.Ip; p += blockDim.x) {
    const bool real = p < d.L || (p >= d.Lp && p - d.Lp < d.D);
    float h = 0.f, hd = 0.f;
    if (real) {
      const float n = (s_c[p] - mu) * rstd;
      const float s1 = __bfloat162float(mrow[p]), sh = __bfloat162float(mrow[d.Ip + p]);
      h = (1.0f + s1) * n + sh;
      if (TANGENT) {
        const float nd = (s_cd[p] - mean_cd - n * mean_ncd) * rstd;
        hd = __bfloat162float(mdrow[p]) * n + (1.0f + s1) * nd + __bfloat162float(mdrow[d.Ip + p]);
      }
    }
    if (a.hin) a.hin[b * d.Ip + p] = __float2bfloat16(h);   // null: tangent-only pass (the primal pass stored hin already)
    if (TANGENT) a.hind[b * d.Ip + p] = __float2bfloat16(hd);
  }
}

// ---------------------------------------------------------------------------------------
// loss: delta = u + (t-r) dudt - (nmax e - x);  s_b = sum delta^2;  w_b = 1/(s_b + c)
//       g_u = 2 w_b delta / B   (weighted)   or   2 delta / (B D)   (plain MSE)
// ---------------------------------------------------------------------------------------
struct LossArgs {
  const float *u, *dudt, *target;  // [B, Dp]; target = noise_max e - x, written by the prologue
  const float *t, *r;         // [B]
  float* g_x;                 // [B, Dp]
  float* row_loss;            // [B]
  float* per_example;         // [B] or null
  MfacImfConfig cfg;
  int64_t B;
};
__global__ void __launch_bounds__(ROW_THREADS) imf_loss_kernel(LossArgs a, Dims d) {
  MFAC_PDL_SYNC();
  extern __shared__ __align__(16) float s_delta[];  // [Dp]
  __shared__ float red[32];
  const int64_t b = blockIdx.x;
  // MeanFlowLoss clips (t - r) to [0, 1] (loss_strategies.py:178); ImprovedMeanFlowLoss does not (:270)
  float tr = a.t[b] - a.r[b];
  if (a.cfg.method == MFAC_LOSS_MEAN_FLOW) tr = fminf(fmaxf(tr, 0.f), 1.f);
  float sq = 0.f;
  const bool vec = d.D == d.Dp && (d.D & 3) == 0;   // unpadded rows: 16-byte accesses
  // (t - r) == 0 (every r == t row): v_pred = u + 0 * du/dt = u, and du/dt is not read -- the step does not even compute it
  // for the leading r == t rows (imf.cu)
  const bool use_dudt = a.dudt != nullptr && tr != 0.f;
  if (vec) {
    for (int j = threadIdx.x * 4; j < d.Dp; j += blockDim.x * 4) {
      const int64_t i = b * d.Dp + j;
      float4 vp = *reinterpret_cast<const float4*>(a.u + i);
      if (use_dudt) {
        const float4 dd = *reinterpret_cast<const float4*>(a.dudt + i);
        vp.x += tr * dd.x; vp.y += tr * dd.y; vp.z += tr * dd.z; vp.w += tr * dd.w;
      }
      const float4 tg = *reinterpret_cast<const float4*>(a.target + i);
      const float4 dl = make_float4(vp.x - tg.x, vp.y - tg.y, vp.z - tg.z, vp.w - tg.w);
      *reinterpret_cast<float4*>(s_delta + j) = dl;
      sq += dl.x * dl.x + dl.y * dl.y + dl.z * dl.z + dl.w * dl.w;
    }
  } else {
    for (int j = threadIdx.x; j < d.Dp; j += blockDim.x) {
      float dl = 0.f;
      if (j < d.D) {
        const int64_t i = b * d.Dp + j;
        const float vpred = use_dudt ? a.u[i] + tr * a.dudt[i] : a.u[i];
        dl = vpred - a.target[i];
      }
      s_delta[j] = dl;
      sq += dl * dl;
    }
  }
  const float s = block_sum(sq, red);
  float w, rl;
  if (a.cfg.method == MFAC_LOSS_MEAN_FLOW) {
    // adaptive reweighting: w = 1 / (mean_D delta^2 + c)^(1 - gamma), loss = mean_B(w mean_D delta^2)
    const float dsq = s / (float)d.D;
    w = powf(dsq + a.cfg.loss_c, -(1.0f - a.cfg.gamma));
    rl = w * dsq / (float)a.B;
    w = 2.0f * w / ((float)d.D * (float)a.B);
  } else if (a.cfg.use_weighted_loss) {
    w = 1.0f / (s + a.cfg.loss_c);
    rl = w * s / (float)a.B;
    w = 2.0f * w / (float)a.B;
  } else {
    rl = s / ((float)a.B * (float)d.D);
    w = 2.0f / ((float)a.B * (float)d.D);
  }
  if (threadIdx.x == 0) {
    a.row_loss[b] = rl;
    if (a.per_example) a.per_example[b] = s;
  }
  if (vec) {
    for (int j = threadIdx.x * 4; j < d.Dp; j += blockDim.x * 4) {
      const float4 dl = *reinterpret_cast<const float4*>(s_delta + j);
      *reinterpret_cast<float4*>(a.g_x + b * d.Dp + j) = make_float4(w * dl.x, w * dl.y, w * dl.z, w * dl.w);
    }
  } else {
    for (int j = threadIdx.x; j < d.Dp; j += blockDim.x) a.g_x[b * d.Dp + j] = w * s_delta[j];
  }
}
// deterministic sum of row_loss[B] -> loss
__global__ void __launch_bounds__(1024) sum_rows_kernel(const float* v, int64_t n, float* out) {
  MFAC_PDL_SYNC();
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = s;
}

// ---------------------------------------------------------------------------------------
// backward elementwise pieces
// ---------------------------------------------------------------------------------------
// LayerNorm / modulation backward, one block per row.
struct LnBwdArgs {
  // g_hin (= g_shift) arrives in bf16, already sitting in its final place g_m[:, Ip:2Ip] (written by the dX GEMM epilogue)
  const float* lat;         // [B, Lp]
  const float* x;           // [B, Dp]  block input
  const float *mu, *rstd;   // [B]
  const __nv_bfloat16* m;   // [B, Mp]
  __nv_bfloat16* g_m;       // [B, Mp]  reads [Ip, 2Ip), writes [0, Ip)
  float* g_lat;             // [B, Lp]  +=
  float* g_x;               // [B, Dp]  +=
  int reverse;              // vectorised kernel: 1 = CTAs walk the rows from the last to the first
};
__global__ void __launch_bounds__(ROW_THREADS) ln_bwd_kernel(LnBwdArgs a, Dims d) {
  extern __shared__ float s_row[];  // n[Ip], g_n[Ip]
  __shared__ float red[32];
  float* s_n = s_row;
  float* s_gn = s_row + d.Ip;
  const int64_t b = blockIdx.x;
  const float mu = a.mu[b], rstd = a.rstd[b];
  float s1sum = 0.f, s2sum = 0.f;
  for (int p = threadIdx.x; p < d.Ip; p += blockDim.x) {
    const bool real = p < d.L || (p >= d.Lp && p - d.Lp < d.D);
    float n = 0.f, gn = 0.f, gh = 0.f;
    if (real) {
      const float c = p < d.Lp ? a.lat[b * d.Lp + p] : a.x[b * d.Dp + (p - d.Lp)];
      n = (c - mu) * rstd;
      gh = __bfloat162float(a.g_m[b * d.Mp + d.Ip + p]);
      gn = gh * (1.0f + __bfloat162float(a.m[b * d.Mp + p]));
    }
    s_n[p] = n;
    s_gn[p] = gn;
    s1sum += gn;
    s2sum += gn * n;
    a.g_m[b * d.Mp + p] = __float2bfloat16(gh * n);          // g_s1  (g_shift = g_hin is already in place)
  }
  const float inv_i = 1.0f / (float)d.I;
  const float m1 = block_sum(s1sum, red) * inv_i;
  const float m2 = block_sum(s2sum, red) * inv_i;
  for (int p = threadIdx.x; p < d.Ip; p += blockDim.x) {
    const bool real = p < d.L || (p >= d.Lp && p - d.Lp < d.D);
    if (!real) continue;
    const float gc = (s_gn[p] - m1 - s_n[p] * m2) * rstd;
    if (p < d.Lp) a.g_lat[b * d.Lp + p] += gc;
    else a.g_x[b * d.Dp + (p - d.Lp)] += gc;
  }
}

__global__ void f32_to_bf16_kernel(const float* src, __nv_bfloat16* dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(src[i]);
}

// padded -> real column maps of the flat (Flax-order) gradient
constexpr int COLSUM_VROWS = 256;
// rows per CTA slab of the column-sum style kernels: 256 for large batches (few atomics), down to 8 for small ones so that a
// 128-row batch still spreads over dozens of CTAs instead of one per 256 columns
inline int colsum_vrows(int64_t B) {
  int v = COLSUM_VROWS;
  while (v > 8 && B / v < 64) v >>= 1;
  return v;
}
enum ColMap { MAP_ID = 0, MAP_CM = 1, MAP_MM = 2, MAP_C1ALL = 3 };
__device__ __forceinline__ int map_col(int kind, int p, int limit, const Dims& d) {
  if (kind == MAP_CM) return d.cm_inv(p);
  if (kind == MAP_MM) return d.mm_inv(p);
  return p < limit ? p : -1;
}
// ---------------------------------------------------------------------------------------
// sampler elementwise updates on padded [B, Dp] fp32 buffers
// ---------------------------------------------------------------------------------------
// out = a + alpha * (g1 * k1 + g2 * k2)     (a and k2 may be null)
__global__ void axpy2_kernel(const float* a, float alpha, float g1, const float* k1, float g2, const float* k2, float* out,
                             int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = g1 * k1[i];
  if (k2) v += g2 * k2[i];
  out[i] = (a ? a[i] : 0.f) + alpha * v;
}
__global__ void fill_normal_kernel(float* out, int Dp, int D, int64_t B, uint64_t seed) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= B * Dp) return;
  const float4 n4 = philox_normal4((uint64_t)(i4 >> 2), 7u, seed, 0);
  const float v[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) out[i4 + q] = ((int)((i4 + q) % Dp) < D) ? v[q] : 0.f;
}

// ---------------------------------------------------------------------------------------
// Vectorised fast paths of the row kernels: one warp per row, the whole row in registers, 16/32-byte
// accesses.  Preconditions (checked by the host): D == Dp, L == Lp (no padded columns) and
// Ip == 256 * NV with NV <= 8.  Lane l owns the 8-column vectors v = l + 32 i, i < NV.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void ld8_f32(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8_f32(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void ld8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ uint4 ld8_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void cvt8_bf16(const uint4 u, float (&v)[8]) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

template <int NV, bool TANGENT>
__global__ void __launch_bounds__(TANGENT ? 128 : 256, TANGENT ? 3 : 1) lnmod_vec_kernel(LnModArgs a, Dims d, int64_t B) {
  MFAC_PDL_SYNC();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)(a.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float c[NV][8], cd[TANGENT ? NV : 1][8];
  float sum = 0.f, sq = 0.f, sumd = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = 8 * (lane + 32 * i);
    if (col < d.Lp) {
      if (a.lat) ld8_f32(a.lat + b * d.Lp + col, c[i]);
      else {
#pragma unroll
        for (int q = 0; q < 8; ++q) c[i][q] = 0.f;
      }
      if (TANGENT) {
#pragma unroll
        for (int q = 0; q < 8; ++q) cd[i][q] = 0.f;
      }
    } else {
      ld8_f32(a.x + b * d.Dp + (col - d.Lp), c[i]);
      if (TANGENT) ld8_f32(a.xd + b * d.Dp + (col - d.Lp), cd[i]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      sum += c[i][q];
      sq += c[i][q] * c[i][q];
      if (TANGENT) sumd += cd[i][q];
    }
  }
  // The stores below may alias the loads as far as the compiler can tell, so it would keep every modulation load behind
  // the previous chunk's store (one exposed memory latency per chunk).  The tangent variant (half the occupancy of the
  // primal one) therefore requests all of its modulation vectors here, before the reductions, and keeps them packed.
  const __nv_bfloat16* mrow = a.m + b * a.m_stride;
  uint4 r_s1[TANGENT ? NV : 1], r_sh[TANGENT ? NV : 1], r_s1d[TANGENT ? NV : 1], r_shd[TANGENT ? NV : 1];
  if (TANGENT) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = 8 * (lane + 32 * i);
      r_s1[i] = ld8_raw(mrow + col);
      r_sh[i] = ld8_raw(mrow + d.Ip + col);
    }
  }
  const float inv_i = 1.0f / (float)d.I;
  const float mu = warp_sum(sum) * inv_i;
  const float rstd = rsqrtf(fmaxf(0.f, warp_sum(sq) * inv_i - mu * mu) + LN_EPS);
  float mean_cd = 0.f, mean_ncd = 0.f;
  if (TANGENT) {
    mean_cd = warp_sum(sumd) * inv_i;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc += (c[i][q] - mu) * rstd * cd[i][q];
    mean_ncd = warp_sum(acc) * inv_i;
  }
  if (lane == 0 && a.mu) { a.mu[b] = mu; a.rstd[b] = rstd; }
  if (TANGENT) {   // second (and last) batch of loads: the tangent modulation vectors
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = 8 * (lane + 32 * i);
      r_s1d[i] = ld8_raw(a.md + b * d.Mp + col);
      r_shd[i] = ld8_raw(a.md + b * d.Mp + d.Ip + col);
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = 8 * (lane + 32 * i);
    float s1[8], sh[8], h[8], n[8];
    if (TANGENT) {
      cvt8_bf16(r_s1[i], s1);
      cvt8_bf16(r_sh[i], sh);
    } else {
      ld8_bf16(mrow + col, s1);
      ld8_bf16(mrow + d.Ip + col, sh);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      n[q] = (c[i][q] - mu) * rstd;
      h[q] = (1.0f + s1[q]) * n[q] + sh[q];
    }
    if (!TANGENT || a.hin) st8_bf16(a.hin + b * d.Ip + col, h);   // hin == null: tangent-only pass
    if (TANGENT) {
      float s1d[8], shd[8];
      cvt8_bf16(r_s1d[i], s1d);
      cvt8_bf16(r_shd[i], shd);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float nd = (cd[i][q] - mean_cd - n[q] * mean_ncd) * rstd;
        h[q] = s1d[q] * n[q] + (1.0f + s1[q]) * nd + shd[q];
      }
      st8_bf16(a.hind + b * d.Ip + col, h);
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_vec_kernel(LnBwdArgs a, Dims d, int64_t B) {
  MFAC_PDL_SYNC();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)(a.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  // every load of the row is requested before the first store (the stores alias the loads as far as the compiler knows,
  // which would otherwise serialise one memory latency per chunk)
  float n[NV][8], acc[NV][8];
  uint4 r_gh[NV], r_s1[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = 8 * (lane + 32 * i);
    if (col < d.Lp) {
      ld8_f32(a.lat + b * d.Lp + col, n[i]);
      ld8_f32(a.g_lat + b * d.Lp + col, acc[i]);
    } else {
      ld8_f32(a.x + b * d.Dp + (col - d.Lp), n[i]);
      ld8_f32(a.g_x + b * d.Dp + (col - d.Lp), acc[i]);
    }
    r_gh[i] = ld8_raw(a.g_m + b * d.Mp + d.Ip + col);
    r_s1[i] = ld8_raw(a.m + b * d.Mp + col);
  }
  const float mu = a.mu[b], rstd = a.rstd[b];
  float gn[NV][8];
  float s1sum = 0.f, s2sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = 8 * (lane + 32 * i);
    float gh[8], s1[8], gs1[8];
    cvt8_bf16(r_gh[i], gh);
    cvt8_bf16(r_s1[i], s1);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      n[i][q] = (n[i][q] - mu) * rstd;
      gn[i][q] = gh[q] * (1.0f + s1[q]);
      gs1[q] = gh[q] * n[i][q];
      s1sum += gn[i][q];
      s2sum += gn[i][q] * n[i][q];
    }
    st8_bf16(a.g_m + b * d.Mp + col, gs1);
  }
  const float inv_i = 1.0f / (float)d.I;
  const float m1 = warp_sum(s1sum) * inv_i, m2 = warp_sum(s2sum) * inv_i;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = 8 * (lane + 32 * i);
    float* dst = col < d.Lp ? a.g_lat + b * d.Lp + col : a.g_x + b * d.Dp + (col - d.Lp);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[i][q] += (gn[i][q] - m1 - n[i][q] * m2) * rstd;
    st8_f32(dst, acc[i]);
  }
}


// ---------------------------------------------------------------------------------------
// Wide rows (Ip > 2048: the noise-dimension sweep, D = 3584 / 7680): one CTA of 256 threads per row, the row in registers
// as up to WIDE_MAXV 8-column vectors per thread (v = tid + 256 i), 16/32-byte accesses, block-wide reductions.  Same
// arithmetic as the warp-per-row kernels above.  Preconditions: D == Dp, L == Lp, Ip <= 8 * 256 * WIDE_MAXV.
// ---------------------------------------------------------------------------------------
constexpr int WIDE_MAXV = 4;
__device__ __forceinline__ float2 block_sum2(float2 v, float2* red /*[32]*/) {
  v.x = warp_sum(v.x); v.y = warp_sum(v.y);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float2 t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : make_float2(0.f, 0.f);
  if (w == 0) {
    t.x = warp_sum(t.x); t.y = warp_sum(t.y);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

template <bool TANGENT>
__global__ void __launch_bounds__(256) lnmod_wide_kernel(LnModArgs a, Dims d) {
  MFAC_PDL_SYNC();
  __shared__ float2 red[32];
  const int64_t b = a.reverse ? (int64_t)gridDim.x - 1 - blockIdx.x : blockIdx.x;
  float c[WIDE_MAXV][8], cd[TANGENT ? WIDE_MAXV : 1][8];
  // every load of the row (data and modulation, primal and tangent) is requested before the first reduction: one memory round
  // trip per row instead of one per phase (the CTA is the only thing hiding latency here)
  uint4 r_s1[WIDE_MAXV], r_sh[WIDE_MAXV], r_s1d[TANGENT ? WIDE_MAXV : 1], r_shd[TANGENT ? WIDE_MAXV : 1];
  const __nv_bfloat16* mrow = a.m + b * a.m_stride;
  float2 s12 = make_float2(0.f, 0.f);
  float sumd = 0.f;
#pragma unroll
  for (int i = 0; i < WIDE_MAXV; ++i) {
    const int col = 8 * ((int)threadIdx.x + 256 * i);
#pragma unroll
    for (int q = 0; q < 8; ++q) { c[i][q] = 0.f; if (TANGENT) cd[i][q] = 0.f; }
    r_s1[i] = make_uint4(0u, 0u, 0u, 0u);
    r_sh[i] = make_uint4(0u, 0u, 0u, 0u);
    if (TANGENT) { r_s1d[i] = make_uint4(0u, 0u, 0u, 0u); r_shd[i] = make_uint4(0u, 0u, 0u, 0u); }
    if (col < d.Ip) {
      if (col < d.Lp) {
        if (a.lat) ld8_f32(a.lat + b * d.Lp + col, c[i]);
      } else {
        ld8_f32(a.x + b * d.Dp + (col - d.Lp), c[i]);
        if (TANGENT) ld8_f32(a.xd + b * d.Dp + (col - d.Lp), cd[i]);
      }
      r_s1[i] = ld8_raw(mrow + col);
      r_sh[i] = ld8_raw(mrow + d.Ip + col);
      if (TANGENT) {
        r_s1d[i] = ld8_raw(a.md + b * d.Mp + col);
        r_shd[i] = ld8_raw(a.md + b * d.Mp + d.Ip + col);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < WIDE_MAXV; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      s12.x += c[i][q];
      s12.y += c[i][q] * c[i][q];
      if (TANGENT) sumd += cd[i][q];
    }
  const float inv_i = 1.0f / (float)d.I;
  const float2 tot = block_sum2(s12, red);
  const float mu = tot.x * inv_i;
  const float rstd = rsqrtf(fmaxf(0.f, tot.y * inv_i - mu * mu) + LN_EPS);
  float mean_cd = 0.f, mean_ncd = 0.f;
  if (TANGENT) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < WIDE_MAXV; ++i) {
      const bool live = 8 * ((int)threadIdx.x + 256 * i) < d.Ip;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc += live ? (c[i][q] - mu) * rstd * cd[i][q] : 0.f;
    }
    const float2 t2 = block_sum2(make_float2(sumd, acc), red);
    mean_cd = t2.x * inv_i;
    mean_ncd = t2.y * inv_i;
  }
  if (threadIdx.x == 0 && a.mu) { a.mu[b] = mu; a.rstd[b] = rstd; }
#pragma unroll
  for (int i = 0; i < WIDE_MAXV; ++i) {
    const int col = 8 * ((int)threadIdx.x + 256 * i);
    if (col >= d.Ip) continue;
    float s1[8], sh[8], h[8], n[8];
    cvt8_bf16(r_s1[i], s1);
    cvt8_bf16(r_sh[i], sh);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      n[q] = (c[i][q] - mu) * rstd;
      h[q] = (1.0f + s1[q]) * n[q] + sh[q];
    }
    if (!TANGENT || a.hin) st8_bf16(a.hin + b * d.Ip + col, h);
    if (TANGENT) {
      float s1d[8], shd[8];
      cvt8_bf16(r_s1d[i], s1d);
      cvt8_bf16(r_shd[i], shd);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float nd = (cd[i][q] - mean_cd - n[q] * mean_ncd) * rstd;
        h[q] = s1d[q] * n[q] + (1.0f + s1[q]) * nd + shd[q];
      }
      st8_bf16(a.hind + b * d.Ip + col, h);
    }
  }
}

__global__ void __launch_bounds__(256) ln_bwd_wide_kernel(LnBwdArgs a, Dims d) {
  MFAC_PDL_SYNC();
  __shared__ float2 red[32];
  const int64_t b = a.reverse ? (int64_t)gridDim.x - 1 - blockIdx.x : blockIdx.x;
  float n[WIDE_MAXV][8], acc[WIDE_MAXV][8], gn[WIDE_MAXV][8];
  uint4 r_gh[WIDE_MAXV], r_s1[WIDE_MAXV];
#pragma unroll
  for (int i = 0; i < WIDE_MAXV; ++i) {
    const int col = 8 * ((int)threadIdx.x + 256 * i);
    r_gh[i] = make_uint4(0u, 0u, 0u, 0u);
    r_s1[i] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int q = 0; q < 8; ++q) { n[i][q] = 0.f; acc[i][q] = 0.f; }
    if (col < d.Ip) {
      if (col < d.Lp) {
        ld8_f32(a.lat + b * d.Lp + col, n[i]);
        ld8_f32(a.g_lat + b * d.Lp + col, acc[i]);
      } else {
        ld8_f32(a.x + b * d.Dp + (col - d.Lp), n[i]);
        ld8_f32(a.g_x + b * d.Dp + (col - d.Lp), acc[i]);
      }
      r_gh[i] = ld8_raw(a.g_m + b * d.Mp + d.Ip + col);
      r_s1[i] = ld8_raw(a.m + b * d.Mp + col);
    }
  }
  const float mu = a.mu[b], rstd = a.rstd[b];
  float2 s12 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < WIDE_MAXV; ++i) {
    const int col = 8 * ((int)threadIdx.x + 256 * i);
    float gh[8], s1[8], gs1[8];
    cvt8_bf16(r_gh[i], gh);
    cvt8_bf16(r_s1[i], s1);
    const bool live = col < d.Ip;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      n[i][q] = live ? (n[i][q] - mu) * rstd : 0.f;
      gn[i][q] = gh[q] * (1.0f + s1[q]);
      gs1[q] = gh[q] * n[i][q];
      s12.x += gn[i][q];
      s12.y += gn[i][q] * n[i][q];
    }
    if (live) st8_bf16(a.g_m + b * d.Mp + col, gs1);
  }
  const float inv_i = 1.0f / (float)d.I;
  const float2 tot = block_sum2(s12, red);
  const float m1 = tot.x * inv_i, m2 = tot.y * inv_i;
#pragma unroll
  for (int i = 0; i < WIDE_MAXV; ++i) {
    const int col = 8 * ((int)threadIdx.x + 256 * i);
    if (col >= d.Ip) continue;
    float* dst = col < d.Lp ? a.g_lat + b * d.Lp + col : a.g_x + b * d.Dp + (col - d.Lp);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[i][q] += (gn[i][q] - m1 - n[i][q] * m2) * rstd;
    st8_f32(dst, acc[i]);
  }
}

// Backward of the block output, fused with the two bias-gradient column sums that read its results:
//   g_o = g_x (1+s2)/nb -> bf16 (GEMM operand), db2 += colsum(g_o);  g_s2 = g_x o / nb -> g_m[:, 2Ip:], dbc2[s2 part] += colsum.
// grid (Dp/256, ceil(B/256)); thread = 8 columns x every 8th row of a 256-row slab (same shape as the column-sum kernel).
__global__ void __launch_bounds__(256) bwd_block_out_vec_kernel(const float* g_x, const __nv_bfloat16* m, const __nv_bfloat16* o,
                                                                __nv_bfloat16* g_o, __nv_bfloat16* g_m, float* db_o, float* db_m,
                                                                Dims d, int64_t B, int reverse, int vrows) {
  MFAC_PDL_SYNC();
  __shared__ float s_red[2][8][256];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + cg * 8;
  const int64_t r0 = (int64_t)(reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y) * vrows;
  const int64_t r1 = min(B, r0 + vrows);
  const float inv_nb = 1.0f / (float)d.nb;
  float so[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ss[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < d.Dp) {
#pragma unroll 2
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      float g[8], s2[8], ov[8], go[8], gs2[8];
      ld8_f32(g_x + r * d.Dp + col, g);
      ld8_bf16(m + r * d.Mp + 2 * d.Ip + col, s2);
      ld8_bf16(o + r * d.Dp + col, ov);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        go[q] = g[q] * (1.0f + s2[q]) * inv_nb;
        gs2[q] = g[q] * ov[q] * inv_nb;
        so[q] += go[q];
        ss[q] += gs2[q];
      }
      st8_bf16(g_o + r * d.Dp + col, go);
      st8_bf16(g_m + r * d.Mp + 2 * d.Ip + col, gs2);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) { s_red[0][rl][cg * 8 + q] = so[q]; s_red[1][rl][cg * 8 + q] = ss[q]; }
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < d.Dp) {
    float a = 0.f, b2 = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) { a += s_red[0][q][threadIdx.x]; b2 += s_red[1][q][threadIdx.x]; }
    const int co = map_col(MAP_ID, c, d.D, d);
    if (co >= 0) atomicAdd(db_o + co, a);
    const int cm = map_col(MAP_MM, 2 * d.Ip + c, 0, d);
    if (cm >= 0) atomicAdd(db_m + cm, b2);
  }
}

// Column sums (bias gradients), one pass: grid (ceil(ld/256), R); thread = 8 columns x every 8th row of a 256-row slab;
// the slab's sums are added to out[map(col)] with red.global.add (out is zeroed with the rest of the gradient, like the
// split-K weight gradients it sits next to).
__global__ void __launch_bounds__(256) colsum_atomic_vec_kernel(const __nv_bfloat16* G, int ld, int ncols, int64_t B, float* out,
                                                                int kind, int limit, Dims d, int vrows) {
  MFAC_PDL_SYNC();
  __shared__ float s_red[8][256];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + cg * 8;
  const int64_t r0 = (int64_t)blockIdx.y * vrows;
  const int64_t r1 = min(B, r0 + vrows);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < ncols) {
#pragma unroll 2
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      float v[8];
      ld8_bf16(G + r * ld + col, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += v[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) s_red[rl][cg * 8 + q] = acc[q];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < ncols) {
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) sum += s_red[q][threadIdx.x];
    if (kind == MAP_C1ALL) {   // batched cond1 bias: column k*Cp + cc of [B, nb*Cp] -> block k's cond1.bias[cc]
      const int k = c / d.Cp, cc = c - k * d.Cp;
      if (cc < d.C) atomicAdd(out + (int64_t)k * d.blk_stride + d.o_c1b + cc, sum);
    } else {
      const int oc = map_col(kind, c, limit, d);
      if (oc >= 0) atomicAdd(out + oc, sum);
    }
  }
}

// weight gradient: padded (row, col) -> flat fp32 [rows_real, cols_real] with inverse maps.
// atomic = 1 under split-K: partial products are accumulated with red.global.add (output pre-zeroed).
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
struct EpiGradStore {
  static constexpr const char* name = "grad_store";
  static constexpr int kPrefetchDepth = 0;  // store-only, rolled epilogue
  static constexpr bool kTmaStore = false;
  float* G;       // leaf base in the flat gradient
  int ld;         // real number of columns
  int row_kind, row_limit, col_kind, col_limit;
  int atomic;
  Dims d;
  using Regs = NoRegs;
  using ColRegs = NoRegs;
  __device__ __forceinline__ void prefetch(int, int, int, int, int, int) const {}
  __device__ __forceinline__ void load_col(int, ColRegs&) const {}
  __device__ __forceinline__ void load(int, int, Regs&) const {}
  __device__ __forceinline__ void frag(int row, int col, float4 acc, const Regs&, const ColRegs&) const {
    const int r = map_col(row_kind, row, row_limit, d);
    if (r < 0) return;
    float* dst = G + (int64_t)r * ld;
    const int c0 = map_col(col_kind, col, col_limit, d);
    const int c3 = map_col(col_kind, col + 3, col_limit, d);
    const float v[4] = {acc.x, acc.y, acc.z, acc.w};
    if (c0 >= 0 && c3 == c0 + 3 && (reinterpret_cast<uintptr_t>(dst + c0) & 15) == 0) {
      if (atomic) red_add_v4(dst + c0, acc);
      else st_f4(dst + c0, acc);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = map_col(col_kind, col + j, col_limit, d);
        if (c >= 0) {
          if (atomic) atomicAdd(dst + c, v[j]);
          else dst[c] = v[j];
        }
      }
    }
  }
};


// Weight gradient of the batched first modulation layer: (row r, column k*Cp + c) of cond^T @ g_ac_all -> block k's
// cond1.kernel[r, c] in the flat gradient.  Split-K partials accumulate with red.global.add.
struct EpiGradStoreC1 {
  static constexpr const char* name = "grad_store";
  static constexpr int kPrefetchDepth = 0;
  static constexpr bool kTmaStore = false;
  float* grads;   // flat gradient base
  Dims d;
  using Regs = NoRegs;
  using ColRegs = NoRegs;
  __device__ __forceinline__ void prefetch(int, int, int, int, int, int) const {}
  __device__ __forceinline__ void load_col(int, ColRegs&) const {}
  __device__ __forceinline__ void load(int, int, Regs&) const {}
  __device__ __forceinline__ void frag(int row, int col, float4 acc, const Regs&, const ColRegs&) const {
    if (row >= d.C) return;
    const int k = col / d.Cp, c = col - k * d.Cp;
    float* dst = grads + (int64_t)k * d.blk_stride + d.o_c1w + (int64_t)row * d.C + c;
    const float v[4] = {acc.x, acc.y, acc.z, acc.w};
    if (c + 3 < d.C && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      red_add_v4(dst, acc);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < d.C) atomicAdd(dst + j, v[j]);
    }
  }
};

}  // namespace mfac
