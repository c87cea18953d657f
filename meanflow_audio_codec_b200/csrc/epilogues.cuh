// Fused GEMM epilogue functors shared by the MLP flow (imf.cu) and the mixer / convnet forwards (flows.cu).
// Contract: see gemm.cuh ("Epilogue contract").
#pragma once

#include "mfac_common.cuh"

namespace mfac {

constexpr float LN_EPS = 1e-6f;  // flax.linen.LayerNorm(epsilon=1e-6)

// ---------------------------------------------------------------------------------------
// fused GEMM epilogues: frag(row, col, acc) handles 4 consecutive columns of one output row; 8 adjacent
// lanes cover 32 columns of the same row, so every access below is sector/line coalesced (see gemm.cuh).
// Read-only operands go through ld.global.nc so the 8 unrolled fragments' loads are issued back to back.
// GELU uses tanh.approx.f32 (one MUFU op): its 2^-11 relative error is far inside the bf16 rounding of
// the values it feeds.
// ---------------------------------------------------------------------------------------
// gelu(a) = a (1/2 + 1/2 tanh(a (k0 + k0 k1 a^2)));  three packed multiplies, two packed FMAs and two MUFU per pair
__device__ __forceinline__ float2 gelu_fast2(float2 a) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  const float2 t = f2mul(a, a);
  const float2 arg = f2mul(a, f2fma(t, f2splat(k0k1), f2splat(k0)));
  const float2 th = make_float2(tanh_fast(arg.x), tanh_fast(arg.y));
  return f2mul(a, f2fma(th, f2splat(0.5f), f2splat(0.5f)));
}
// gelu'(a) = 1/2 (1 + th) + 1/2 a (1 - th^2) (k0 + 3 k0 k1 a^2)
__device__ __forceinline__ float2 dgelu_fast2(float2 a) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  const float2 t = f2mul(a, a);
  const float2 arg = f2mul(a, f2fma(t, f2splat(k0k1), f2splat(k0)));
  const float2 th = make_float2(tanh_fast(arg.x), tanh_fast(arg.y));
  const float2 q = f2mul(a, f2fma(t, f2splat(1.5f * k0k1), f2splat(0.5f * k0)));      // 1/2 a (k0 + 3 k0 k1 a^2)
  const float2 s = f2fma(make_float2(-th.x, -th.y), th, f2splat(1.0f));                // 1 - th^2
  return f2fma(q, s, f2fma(th, f2splat(0.5f), f2splat(0.5f)));
}
__device__ __forceinline__ float4 gelu_fast4(float4 a) {
  const float2 lo = gelu_fast2(make_float2(a.x, a.y)), hi = gelu_fast2(make_float2(a.z, a.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// acc * gelu'(a)
__device__ __forceinline__ float4 mul_dgelu_fast4(float4 acc, float4 a) {
  const float2 lo = f2mul(make_float2(acc.x, acc.y), dgelu_fast2(make_float2(a.x, a.y)));
  const float2 hi = f2mul(make_float2(acc.z, acc.w), dgelu_fast2(make_float2(a.z, a.w)));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldg_bf4(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 ldg_bf4_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ float4 bf4_to_f4(uint2 u) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st_bf4(__nv_bfloat16* p, float4 v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
}
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  const float2 lo = f2add(make_float2(a.x, a.y), make_float2(b.x, b.y)), hi = f2add(make_float2(a.z, a.w), make_float2(b.z, b.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

constexpr int EPI_THREADS = 256;  // epilogue threads of the GEMM kernel (prefetch work split)
struct BiasCol { float4 b; };
struct NoRegs {};

// g = gelu(acc + bias); optionally keeps the pre-activation a (bf16) for the tangent/backward.
struct EpiBiasGelu {
  static constexpr bool kNarrowTiles = true;
  static constexpr const char* name = "bias_gelu";
  static constexpr int kPrefetchDepth = 1;
  static constexpr bool kTmaStore = false;
  const float* bias;        // padded fp32 [N]
  __nv_bfloat16* g;         // [M, ld]
  __nv_bfloat16* a_out;     // [M, ld] or null
  int64_t ld;
  int keep_rows = 0x7fffffff;   // a_out is only wanted for rows < keep_rows (the shared r == t pass: later rows get theirs from the u pass)
  using Regs = NoRegs;
  using ColRegs = BiasCol;
  __device__ __forceinline__ void prefetch(int, int, int, int, int, int) const {}
  __device__ __forceinline__ void load_col(int col, ColRegs& c) const { c.b = ldg_f4(bias + col); }
  __device__ __forceinline__ void load(int, int, Regs&) const {}
  __device__ __forceinline__ void frag(int row, int col, float4 acc, const Regs&, const ColRegs& c) const {
    acc = add4(acc, c.b);
    const int64_t at = (int64_t)row * ld + col;
    if (a_out && row < keep_rows) st_bf4(a_out + at, acc);
    st_bf4(g + at, gelu_fast4(acc));
  }
};
// out = acc * gelu'(a)   (tangent through GELU, and the backward of GELU)
struct EpiMulDgelu {
  static constexpr bool kNarrowTiles = true;
  static constexpr const char* name = "mul_dgelu";
  static constexpr int kPrefetchDepth = 2;
  static constexpr bool kTmaStore = false;
  const __nv_bfloat16* a;   // [M, ld]
  __nv_bfloat16* out;       // [M, ld]
  int64_t ld;
  struct Regs { uint2 av; };
  using ColRegs = NoRegs;
  __device__ __forceinline__ void prefetch(int m0, int n0, int bn, int M, int N, int etid) const {
    bn = min(bn, N - n0);   // never reach past the last column (partial last tile)
    l2_prefetch_tile<2>(a, ld, m0, n0, bn, M, etid, EPI_THREADS);
  }
  __device__ __forceinline__ void load_col(int, ColRegs&) const {}
  __device__ __forceinline__ void load(int row, int col, Regs& r) const { r.av = ldg_bf4_raw(a + (int64_t)row * ld + col); }
  __device__ __forceinline__ void frag(int row, int col, float4 acc, const Regs& r, const ColRegs&) const {
    const int64_t at = (int64_t)row * ld + col;
    const float4 av = bf4_to_f4(r.av);
    st_bf4(out + at, mul_dgelu_fast4(acc, av));
  }
};
// Column-sum variants (backward: the bias gradient is the column sum of the tensor the epilogue writes).  frag() returns
// the fp32 values it stored; the epilogue driver (gemm.cuh, epi_colsum) adds them up over the warp's 32 rows and hands
// one float4 per 4 columns to col_add(), which accumulates it into the fp32 bias gradient with red.global.add.
__device__ __forceinline__ void red_add_f4(float* p, float4 v) {
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  } else {
    atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
  }
}
struct EpiMulDgeluColsum : EpiMulDgelu {
  static constexpr bool kColSum = true;
  float* colsum;   // [N] fp32, accumulated
  __device__ __forceinline__ float4 frag(int row, int col, float4 acc, const Regs& r, const ColRegs&) const {
    const float4 v = mul_dgelu_fast4(acc, bf4_to_f4(r.av));
    st_bf4(out + (int64_t)row * ld + col, v);
    return v;
  }
  __device__ __forceinline__ void col_add(int col, float4 v) const { red_add_f4(colsum + col, v); }
};
// out = acc (+ bias)
struct EpiLinearBf16 {
  static constexpr bool kNarrowTiles = true;
  static constexpr const char* name = "linear_bf16";
  static constexpr int kPrefetchDepth = 1;
  static constexpr bool kTmaStore = false;
  const float* bias;  // or null
  __nv_bfloat16* out;
  int64_t ld;
  using Regs = NoRegs;
  using ColRegs = BiasCol;
  __device__ __forceinline__ void prefetch(int, int, int, int, int, int) const {}
  __device__ __forceinline__ void load_col(int col, ColRegs& c) const { c.b = bias ? ldg_f4(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(int, int, Regs&) const {}
  __device__ __forceinline__ void frag(int row, int col, float4 acc, const Regs&, const ColRegs& c) const {
    st_bf4(out + (int64_t)row * ld + col, add4(acc, c.b));
  }
};
struct EpiLinearBf16Colsum : EpiLinearBf16 {
  static constexpr bool kColSum = true;
  float* colsum;   // [N] fp32, accumulated
  __device__ __forceinline__ float4 frag(int row, int col, float4 acc, const Regs&, const ColRegs& c) const {
    const float4 v = add4(acc, c.b);
    st_bf4(out + (int64_t)row * ld + col, v);
    return v;
  }
  __device__ __forceinline__ void col_add(int col, float4 v) const { red_add_f4(colsum + col, v); }
};
struct EpiLinearF32 {
  static constexpr const char* name = "linear_f32";
  static constexpr int kPrefetchDepth = 1;
  static constexpr bool kTmaStore = false;
  const float* bias;  // or null
  float* out;
  int64_t ld;
  using Regs = NoRegs;
  using ColRegs = BiasCol;
  __device__ __forceinline__ void prefetch(int, int, int, int, int, int) const {}
  __device__ __forceinline__ void load_col(int col, ColRegs& c) const { c.b = bias ? ldg_f4(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(int, int, Regs&) const {}
  __device__ __forceinline__ void frag(int row, int col, float4 acc, const Regs&, const ColRegs& c) const {
    st_f4(out + (int64_t)row * ld + col, add4(acc, c.b));
  }
};
// out = (acc + bias) * alpha + res   (fp32 and/or bf16 copy); res may be null.  Residual adds of the mixer / convnet
// blocks: x / num_blocks + residual  (mlp_mixer.py:163, conv_flow.py:205) and the mixer's inner skip connections.
struct EpiAffineResidual {
  static constexpr const char* name = "affine_residual";
  static constexpr int kPrefetchDepth = 2;
  static constexpr bool kTmaStore = false;
  const float* bias;       // [N] or null
  const float* res;        // [M, ld] or null
  float* out_f;            // [M, ld] or null (may alias res)
  __nv_bfloat16* out_b;    // [M, ld] or null
  int64_t ld;
  float alpha;
  struct Regs { float4 r; };
  using ColRegs = BiasCol;
  __device__ __forceinline__ void prefetch(int m0, int n0, int bn, int M, int N, int etid) const {
    bn = min(bn, N - n0);   // never reach past the last column (partial last tile)
    if (res && (ld & 31) == 0) l2_prefetch_tile<4>(res, ld, m0, n0, bn, M, etid, EPI_THREADS);
  }
  __device__ __forceinline__ void load_col(int col, ColRegs& c) const { c.b = bias ? ldg_f4(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(int row, int col, Regs& r) const {
    r.r = res ? *reinterpret_cast<const float4*>(res + (int64_t)row * ld + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void frag(int row, int col, float4 acc, const Regs& r, const ColRegs& c) const {
    const int64_t at = (int64_t)row * ld + col;
    const float4 v = make_float4((acc.x + c.b.x) * alpha + r.r.x, (acc.y + c.b.y) * alpha + r.r.y,
                                 (acc.z + c.b.z) * alpha + r.r.z, (acc.w + c.b.w) * alpha + r.r.w);
    if (out_f) st_f4(out_f + at, v);
    if (out_b) st_bf4(out_b + at, v);
  }
};
// out += acc * alpha with fp32 reductions at L2: the epilogue of a split-K GEMM whose output already holds everything that is
// not the product (residual + scaled bias).  The mixer's output projection at small batch is [B, 16384] x [16384, 1024]:
// 16 output tiles for 148 SMs unless K is sliced.
struct EpiScaledAtomicAdd {
  static constexpr const char* name = "scaled_atomic_add";
  static constexpr int kPrefetchDepth = 0;
  static constexpr bool kTmaStore = false;
  float* out;
  int64_t ld;
  float alpha;
  using Regs = NoRegs;
  using ColRegs = NoRegs;
  __device__ __forceinline__ void prefetch(int, int, int, int, int, int) const {}
  __device__ __forceinline__ void load_col(int, ColRegs&) const {}
  __device__ __forceinline__ void load(int, int, Regs&) const {}
  __device__ __forceinline__ void frag(int row, int col, float4 acc, const Regs&, const ColRegs&) const {
    red_add_f4(out + (int64_t)row * ld + col, make_float4(acc.x * alpha, acc.y * alpha, acc.z * alpha, acc.w * alpha));
  }
};
// block output: o = acc + b2;  x_new = o (1 + s2) / nb + x_old      (mlp_flow.py:112-117)
struct EpiBlockOut {
  static constexpr bool kNarrowTiles = true;
  static constexpr const char* name = "block_out";
  static constexpr int kPrefetchDepth = 2;
  static constexpr bool kTmaStore = false;
  const float* bias;          // padded [Dp]
  const __nv_bfloat16* m;     // [M, Mp]; s2 at column offset s2_off
  const float* x_old;         // [M, Dp]
  float* x_new;               // [M, Dp] (may alias x_old: each element is read and written by the same thread)
  __nv_bfloat16* o_out;       // [M, Dp] or null
  int64_t ldm, ldx;
  int s2_off;
  float inv_nb;
  int keep_rows = 0x7fffffff;   // o_out is only wanted for rows < keep_rows
  struct Regs { uint2 s2; float4 xo; };
  using ColRegs = BiasCol;
  __device__ __forceinline__ void prefetch(int m0, int n0, int bn, int M, int N, int etid) const {
    bn = min(bn, N - n0);   // never reach past the last column (partial last tile)
    l2_prefetch_tile<2>(m, ldm, m0, s2_off + n0, bn, M, etid, EPI_THREADS);
    l2_prefetch_tile<4>(x_old, ldx, m0, n0, bn, M, etid, EPI_THREADS);
  }
  __device__ __forceinline__ void load_col(int col, ColRegs& c) const { c.b = ldg_f4(bias + col); }
  __device__ __forceinline__ void load(int row, int col, Regs& r) const {
    r.s2 = ldg_bf4_raw(m + (int64_t)row * ldm + s2_off + col);
    r.xo = ldg_f4(x_old + (int64_t)row * ldx + col);
  }
  __device__ __forceinline__ void frag(int row, int col, float4 acc, const Regs& r, const ColRegs& c) const {
    const int64_t at = (int64_t)row * ldx + col;
    const float4 s2 = bf4_to_f4(r.s2), xo = r.xo;
    acc = add4(acc, c.b);
    if (o_out && row < keep_rows) st_bf4(o_out + at, acc);
    st_f4(x_new + at, make_float4(acc.x * ((1.0f + s2.x) * inv_nb) + xo.x, acc.y * ((1.0f + s2.y) * inv_nb) + xo.y,
                                  acc.z * ((1.0f + s2.z) * inv_nb) + xo.z, acc.w * ((1.0f + s2.w) * inv_nb) + xo.w));
  }
};
// tangent of the block output: xd_new = (od (1+s2) + o s2d) / nb + xd_old
struct EpiBlockOutTangent {
  static constexpr bool kNarrowTiles = true;
  static constexpr const char* name = "block_out_tangent";
  static constexpr int kPrefetchDepth = 1;
  static constexpr bool kTmaStore = false;
  const __nv_bfloat16* m;     // primal modulation  [M, Mp]
  const __nv_bfloat16* md;    // tangent modulation [M, Mp]
  const __nv_bfloat16* o;     // primal o [M, Dp]
  const float* xd_old;
  float* xd_new;
  int64_t ldm, ldx;
  int s2_off;
  float inv_nb;
  struct Regs { uint2 s2, s2d, ov; float4 xd; };
  using ColRegs = NoRegs;
  __device__ __forceinline__ void prefetch(int m0, int n0, int bn, int M, int N, int etid) const {
    bn = min(bn, N - n0);   // never reach past the last column (partial last tile)
    l2_prefetch_tile<2>(m, ldm, m0, s2_off + n0, bn, M, etid, EPI_THREADS);
    l2_prefetch_tile<2>(md, ldm, m0, s2_off + n0, bn, M, etid, EPI_THREADS);
    l2_prefetch_tile<2>(o, ldx, m0, n0, bn, M, etid, EPI_THREADS);
    l2_prefetch_tile<4>(xd_old, ldx, m0, n0, bn, M, etid, EPI_THREADS);
  }
  __device__ __forceinline__ void load_col(int, ColRegs&) const {}
  __device__ __forceinline__ void load(int row, int col, Regs& r) const {
    const int64_t at = (int64_t)row * ldx + col, am = (int64_t)row * ldm + s2_off + col;
    r.s2 = ldg_bf4_raw(m + am); r.s2d = ldg_bf4_raw(md + am); r.ov = ldg_bf4_raw(o + at); r.xd = ldg_f4(xd_old + at);
  }
  __device__ __forceinline__ void frag(int row, int col, float4 acc, const Regs& r, const ColRegs&) const {
    const int64_t at = (int64_t)row * ldx + col;
    const float4 s2 = bf4_to_f4(r.s2), s2d = bf4_to_f4(r.s2d), ov = bf4_to_f4(r.ov), xd = r.xd;
    st_f4(xd_new + at, make_float4((acc.x * (1.0f + s2.x) + ov.x * s2d.x) * inv_nb + xd.x,
                                   (acc.y * (1.0f + s2.y) + ov.y * s2d.y) * inv_nb + xd.y,
                                   (acc.z * (1.0f + s2.z) + ov.z * s2d.z) * inv_nb + xd.z,
                                   (acc.w * (1.0f + s2.w) + ov.w * s2d.w) * inv_nb + xd.w));
  }
};

// ---------------------------------------------------------------------------------------
// TMA-store functors (gemm.cuh: epilogue_tile_tma).  compute() sees 32 consecutive columns of ONE row (the accumulator's
// native layout) and returns them as 16 packed bf16 pairs; N must be a multiple of 64.
// ---------------------------------------------------------------------------------------
struct EpiTmaps;
int make_tmap_bf16_sw(CUtensorMap* out, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                      int box_outer, int swizzle_bytes);

// out = acc (+ bias) -> bf16
struct EpiLinearBf16Tma {
  static constexpr const char* name = "linear_bf16";
  static constexpr int kPrefetchDepth = 1;
  static constexpr bool kTmaStore = true;
  static constexpr int kNumOut = 1;
  const float* bias;  // or null
  __nv_bfloat16* out;
  int64_t ld;
  __device__ __forceinline__ void prefetch(int, int, int, int, int, int) const {}
  template <class Maps>
  int make_maps(Maps& m, int M, int N) const {
    const int rc = make_tmap_bf16_sw(&m.m[0], out, N, M, ld, 64, 32, 128);
    m.m[1] = m.m[0];
    return rc;
  }
  __device__ __forceinline__ void compute(int col0, const float (&acc)[32], uint32_t (&o)[16]) const {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = bias ? ldg_f4(bias + col0 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 v = add4(make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]), b);
      o[2 * q] = pack_bf16(v.x, v.y);
      o[2 * q + 1] = pack_bf16(v.z, v.w);
    }
  }
  __device__ __forceinline__ void store_row(int row, int col0, const float (&acc)[32]) const {
    for (int j = 0; j < 32; ++j) out[(int64_t)row * ld + col0 + j] = __float2bfloat16(acc[j] + (bias ? bias[col0 + j] : 0.f));
  }
};

// g = gelu(acc + bias) -> bf16  (forward passes that keep no pre-activation: v pass, samplers, mixer / convnet)
struct EpiBiasGeluTma {
  static constexpr const char* name = "bias_gelu";
  static constexpr int kPrefetchDepth = 1;
  static constexpr bool kTmaStore = true;
  static constexpr int kNumOut = 1;
  const float* bias;
  __nv_bfloat16* g;
  int64_t ld;
  __device__ __forceinline__ void prefetch(int, int, int, int, int, int) const {}
  template <class Maps>
  int make_maps(Maps& m, int M, int N) const {
    const int rc = make_tmap_bf16_sw(&m.m[0], g, N, M, ld, 64, 32, 128);
    m.m[1] = m.m[0];
    return rc;
  }
  __device__ __forceinline__ void compute(int col0, const float (&acc)[32], uint32_t (&o)[16]) const {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = add4(make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]), ldg_f4(bias + col0 + 4 * q));
      const float4 gv = gelu_fast4(v);
      o[2 * q] = pack_bf16(gv.x, gv.y);
      o[2 * q + 1] = pack_bf16(gv.z, gv.w);
    }
  }
  __device__ __forceinline__ void store_row(int row, int col0, const float (&acc)[32]) const {
    for (int j = 0; j < 32; ++j) {
      const float v = acc[j] + bias[col0 + j];
      g[(int64_t)row * ld + col0 + j] = __float2bfloat16(gelu_fast2(make_float2(v, v)).x);
    }
  }
};

}  // namespace mfac
