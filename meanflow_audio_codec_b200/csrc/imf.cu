// Orchestration of the MLP velocity network, the iMF loss/gradient step and the samplers.
//
// ref (paths inside /root/reference/meanflow_audio_codec/):
//   ConditionalFlow.__call__ / encode     models/mlp_flow.py:125-230
//   ImprovedMeanFlowLoss.compute_loss     trainers/loss_strategies.py:227-280
//   sample (Heun, h = 0)                  evaluators/sampling.py:5-95
//
// Every matrix product is a launch of the tcgen05 GEMM in gemm.cuh with a fused epilogue from
// imf_kernels.cuh; the remaining work is the row kernels there.  The three network evaluations
// of the iMF loss run as:  v-pass (B rows; with the r == t rule known it is the SAVED primal pass and
// already the u pass of those rows)  ->  u-pass primal interleaved with the tangent pass block by block
// (the JVP shares W1c/W2c/W1/W2 with the primal; only the rows with r != t)  ->  loss  ->  backward
// through the primal u rows and the encoder only (v and du/dt carry no gradient).
#include "gemm.cuh"
#include "imf_kernels.cuh"

namespace mfac {

namespace {

struct Shadow {
  const __nv_bfloat16* w;
  const float* b;
  Shadow(const void* p, const Dims& d)
      : w(reinterpret_cast<const __nv_bfloat16*>(p)),
        b(reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p) + d.bias_section_bytes_offset)) {}
};

inline unsigned blocks_for(int64_t n, int t) { return (unsigned)ceil_div<int64_t>(n, t); }

// ---- GEMM flavours ---------------------------------------------------------------------
// y[M,N] = A[M,K] @ W[K,N]           A K-major (ld lda), W = shadow kernel [in=K, out=N] (MN-major B)
template <class Epi>
int gemm_fwd(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int M, int N, int K, const Epi& epi, cudaStream_t s) {
  return launch_gemm<false, true>(GemmOperandDesc{A, lda, false}, GemmOperandDesc{W, N, true}, M, N, K, epi, s);
}
// y[M,N] = G[M,K] @ W^T              W = shadow kernel [in=N, out=K] read as K-major B
template <class Epi>
int gemm_dx(const __nv_bfloat16* G, int ldg, const __nv_bfloat16* W, int M, int N, int K, const Epi& epi, cudaStream_t s) {
  return launch_gemm<false, false>(GemmOperandDesc{G, ldg, false}, GemmOperandDesc{W, K, false}, M, N, K, epi, s);
}
// dW[M,N] = Act[Bk,M]^T @ G[Bk,N]    both operands MN-major (batch is the contraction)
template <class Epi>
int gemm_dw(const __nv_bfloat16* Act, int lda, const __nv_bfloat16* G, int ldg, int M, int N, int Bk, const Epi& epi,
            cudaStream_t s) {
  return launch_gemm<true, true>(GemmOperandDesc{Act, lda, true}, GemmOperandDesc{G, ldg, true}, M, N, Bk, epi, s, 0,
                                 /*split_k=*/true);
}

// bias + GELU (+ keep pre-activation) and plain linear -> bf16.  Store-bound shapes (small K: the modulation MLP) go
// through the TMA-store epilogue; compute-bound ones keep the transposed epilogue and its deeper operand ring (measured:
// 35 vs 48 us for M=18944, N=3584, K=128, but 58 vs 53 us for N=K=1280).
int gemm_bias_gelu(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int M, int N, int K, const float* bias,
                   __nv_bfloat16* g, __nv_bfloat16* a, int64_t ld, cudaStream_t s) {
  if (a || N % 64 != 0 || K > 256) return gemm_fwd(A, lda, W, M, N, K, EpiBiasGelu{bias, g, a, ld}, s);
  return gemm_fwd(A, lda, W, M, N, K, EpiBiasGeluTma{bias, g, ld}, s);
}
int gemm_linear_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int M, int N, int K, const float* bias,
                     __nv_bfloat16* out, int64_t ld, cudaStream_t s) {
  if (N % 64 != 0 || K > 256) return gemm_fwd(A, lda, W, M, N, K, EpiLinearBf16{bias, out, ld}, s);
  return gemm_fwd(A, lda, W, M, N, K, EpiLinearBf16Tma{bias, out, ld}, s);
}

// NV of the vectorised row kernels, or 0 when the geometry needs the generic (padded) kernels.
inline int vec_rows(const Dims& d) {
  if (d.D != d.Dp || d.L != d.Lp || d.Ip % 256 != 0 || d.Ip > 2048) return 0;
  return d.Ip / 256;
}

// wide rows without padded columns: the CTA-per-row vectorised kernels (imf_kernels.cuh)
inline bool wide_rows(const Dims& d) {
  return vec_rows(d) == 0 && d.D == d.Dp && d.L == d.Lp && d.Ip <= 8 * 256 * WIDE_MAXV;
}

int lnmod(bool tangent, const LnModArgs& a_in, const Dims& d, int64_t B, cudaStream_t s) {
  LnModArgs a = a_in;
  a.reverse = (vec_rows(d) || wide_rows(d)) ? sweep_next() : 0;
  if (wide_rows(d)) {
    if (tangent) launch_pdl(lnmod_wide_kernel<true>, dim3((unsigned)B), dim3(256), 0, s, a, d);
    else launch_pdl(lnmod_wide_kernel<false>, dim3((unsigned)B), dim3(256), 0, s, a, d);
    count_launch();
    return launch_status();
  }
  const size_t smem = (size_t)d.Ip * 4 * (tangent ? 2 : 1);
  if (smem > 200 * 1024) return MFAC_ERR_UNSUPPORTED;
  static PerDeviceOnce configured;
  if (configured.need()) {
    MFAC_CUDA_OK(cudaFuncSetAttribute(lnmod_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MFAC_CUDA_OK(cudaFuncSetAttribute(lnmod_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MFAC_CUDA_OK(cudaFuncSetAttribute(ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MFAC_CUDA_OK(cudaFuncSetAttribute(imf_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured.done();
  }
  const int nv = vec_rows(d);
  // the tangent variant keeps a whole row's operands in flight (~250 registers): 4-row CTAs so that two fit an SM
  const unsigned vgrid = (unsigned)ceil_div<int64_t>(B, 8), tgrid = (unsigned)ceil_div<int64_t>(B, 4);
#define MFAC_LNMOD_CASE(NVV)                                                              \
  case NVV:                                                                               \
    if (tangent) launch_pdl(lnmod_vec_kernel<NVV, true>, dim3(tgrid), dim3(128), 0, s, a, d, B);   \
    else launch_pdl(lnmod_vec_kernel<NVV, false>, dim3(vgrid), dim3(256), 0, s, a, d, B);        \
    break;
  switch (nv) {
    MFAC_LNMOD_CASE(1) MFAC_LNMOD_CASE(2) MFAC_LNMOD_CASE(3) MFAC_LNMOD_CASE(4)
    MFAC_LNMOD_CASE(5) MFAC_LNMOD_CASE(6) MFAC_LNMOD_CASE(7) MFAC_LNMOD_CASE(8)
    default:
      if (tangent) lnmod_kernel<true><<<(unsigned)B, ROW_THREADS, smem, s>>>(a, d);
      else lnmod_kernel<false><<<(unsigned)B, ROW_THREADS, smem, s>>>(a, d);
  }
#undef MFAC_LNMOD_CASE
  count_launch();
  return launch_status();
}

int ln_bwd(const LnBwdArgs& a_in, const Dims& d, int64_t B, cudaStream_t s) {
  LnBwdArgs a = a_in;
  a.reverse = (vec_rows(d) || wide_rows(d)) ? sweep_next() : 0;
  if (wide_rows(d)) {
    launch_pdl(ln_bwd_wide_kernel, dim3((unsigned)B), dim3(256), 0, s, a, d);
    count_launch();
    return launch_status();
  }
  const unsigned vgrid = (unsigned)ceil_div<int64_t>(B, 4);   // ~200 registers per thread: 4-row CTAs, two per SM
#define MFAC_LNBWD_CASE(NVV) case NVV: launch_pdl(ln_bwd_vec_kernel<NVV>, dim3(vgrid), dim3(128), 0, s, a, d, B); break;
  switch (vec_rows(d)) {
    MFAC_LNBWD_CASE(1) MFAC_LNBWD_CASE(2) MFAC_LNBWD_CASE(3) MFAC_LNBWD_CASE(4)
    MFAC_LNBWD_CASE(5) MFAC_LNBWD_CASE(6) MFAC_LNBWD_CASE(7) MFAC_LNBWD_CASE(8)
    default: ln_bwd_kernel<<<(unsigned)B, ROW_THREADS, (size_t)d.Ip * 8, s>>>(a, d);
  }
#undef MFAC_LNBWD_CASE
  count_launch();
  return launch_status();
}

// Batches this small run the concurrent schedules (independent kernel chains on the library's side streams, gemm.cuh: ForkCtx)
// (a batch counts with its width: 4096 rows of D = 1024 are "small", 4096 rows of D = 7680 fill the machine on their own)
inline bool concurrent_rows(int64_t B, const Dims& d) {
  return concurrency_max_rows() > 0 && B <= concurrency_max_rows() && B * (int64_t)d.Dp <= (int64_t)concurrency_max_rows() * 1024;
}

// Scratch of one (unsaved) forward evaluation.
struct FwdScratch {
  __nv_bfloat16 *gc, *m, *hin, *g;
  int64_t m_blk_stride = 0;   // > 0: one modulation buffer per block (m + k * m_blk_stride) -- the hoisted schedule
  void plan(Arena& ar, const Dims& d, int64_t B, bool per_block_m = false) {
    gc = ar.take<__nv_bfloat16>(B * d.Ca);   // first modulation layer of all blocks: [B, nb*Cp]
    m_blk_stride = per_block_m ? B * d.Mp : 0;
    m = ar.take<__nv_bfloat16>(B * d.Mp * (per_block_m ? d.nb : 1));
    hin = ar.take<__nv_bfloat16>(B * d.Ip);
    g = ar.take<__nv_bfloat16>(B * d.Ip);
  }
};

// The second modulation layer of a block depends on the conditioning only, not on the residual stream: in the concurrent
// schedule all blocks' modulation GEMMs of a pass leave the block chain for a side stream (`sm`) right after the batched
// first layer, and block k waits for event k.
struct Hoist {
  ForkCtx* fc = nullptr;
  cudaStream_t sm = nullptr;
  cudaEvent_t ev[64];
  bool on() const { return fc != nullptr; }
  template <class Fn>
  int run(cudaStream_t st, int nb, Fn&& fn) {
    MFAC_OK(stream_after(fc, st, sm));
    for (int k = 0; k < nb; ++k) {
      MFAC_OK(fn(k, sm));
      ev[k] = fc->next_event();
      MFAC_CUDA_OK(cudaEventRecord(ev[k], sm));
    }
    return MFAC_SUCCESS;
  }
  int wait(cudaStream_t st, int k) const { return cuda_status(cudaStreamWaitEvent(st, ev[k], 0)); }
};

// x <- f(x, cond, lat) in place over all blocks (no activations kept).
// uniform_cond: `cond` is ONE row shared by the whole batch (samplers evaluate every row at the same (t, h)); the
// modulation MLP then runs on a single row and its output is broadcast (row stride 0) instead of being
// materialised as a [B, 2I+D] tensor -- 7 KB per row per block less HBM traffic and one large GEMM less.
// x_src (optional): the input lives there and is left untouched -- block 0 reads it (LayerNorm input and residual) and writes x,
// the later blocks run in place on x; saves the samplers one copy of the state per network evaluation.
int forward_pass(const Dims& d, const Shadow& sh, const __nv_bfloat16* cond, const float* lat, float* x, int64_t B,
                 const FwdScratch& sc, cudaStream_t s, bool uniform_cond = false, Hoist* hz = nullptr, const float* x_src = nullptr) {
  const int M = (int)B;
  const int Mc = uniform_cond ? 1 : M;
  const int64_t m_stride = uniform_cond ? 0 : d.Mp;
  const float inv_nb = 1.0f / (float)d.nb;
  const bool hoist = hz && hz->on() && sc.m_blk_stride > 0;
  // first modulation layer of ALL blocks in one GEMM (they share the input row `cond`)
  MFAC_OK(gemm_bias_gelu(cond, d.Cp, sh.w + d.s_c1all, Mc, d.Ca, d.Cp, sh.b + d.b_c1all, sc.gc, nullptr, d.Ca, s));
  auto mod = [&](int k, cudaStream_t st) -> int {
    return gemm_linear_bf16(sc.gc + k * d.Cp, d.Ca, sh.w + k * d.s_blk_stride + d.s_c2w, Mc, d.Mp, d.Cp,
                            sh.b + k * d.b_blk_stride + d.b_c2, sc.m + (hoist ? k * sc.m_blk_stride : 0), d.Mp, st);
  };
  if (hoist) MFAC_OK(hz->run(s, d.nb, mod));
  for (int k = 0; k < d.nb; ++k) {
    const __nv_bfloat16* w = sh.w + k * d.s_blk_stride;
    const float* bias = sh.b + k * d.b_blk_stride;
    const __nv_bfloat16* m = sc.m + (hoist ? k * sc.m_blk_stride : 0);
    if (hoist) MFAC_OK(hz->wait(s, k));
    else MFAC_OK(mod(k, s));
    const float* x_in = (k == 0 && x_src) ? x_src : x;
    LnModArgs la{lat, x_in, m, sc.hin, nullptr, nullptr, nullptr, nullptr, nullptr, m_stride};
    MFAC_OK(lnmod(false, la, d, B, s));
    MFAC_OK(gemm_bias_gelu(sc.hin, d.Ip, w + d.s_m1w, M, d.Ip, d.Ip, bias + d.b_m1, sc.g, nullptr, d.Ip, s));
    MFAC_OK(gemm_fwd(sc.g, d.Ip, w + d.s_m2w, M, d.Dp, d.Ip,
                     EpiBlockOut{bias + d.b_m2, m, x_in, x, nullptr, m_stride, d.Dp, 2 * d.Ip, inv_nb}, s));
  }
  return MFAC_SUCCESS;
}

// latents = enc2(gelu(enc1(x)));  keeps a_e / g_e when asked (backward)
int encoder_pass(const Dims& d, const Shadow& sh, const __nv_bfloat16* xb, __nv_bfloat16* a_e, __nv_bfloat16* g_e, float* lat,
                 int64_t B, cudaStream_t s) {
  MFAC_OK(gemm_bias_gelu(xb, d.Dp, sh.w + d.s_e1w, (int)B, d.Hep, d.Dp, sh.b + d.b_e1, g_e, a_e, d.Hep, s));
  MFAC_OK(gemm_fwd(g_e, d.Hep, sh.w + d.s_e2w, (int)B, d.Lp, d.Hep, EpiLinearF32{sh.b + d.b_e2, lat, d.Lp}, s));
  return MFAC_SUCCESS;
}

// out[map(col)] += sum_rows G[:, col]  (bias gradients; out is zeroed with the rest of the gradient before the backward)
int colsum(const __nv_bfloat16* G, int ld, int64_t B, float* out, int kind, int limit, const Dims& d, cudaStream_t s,
           int ncols = 0) {
  if (ncols == 0) ncols = ld;
  const int vrows = colsum_vrows(B);
  const int R = (int)ceil_div<int64_t>(B, vrows);
  launch_pdl(colsum_atomic_vec_kernel, dim3(ceil_div(ncols, 256), R), dim3(256), 0, s, G, ld, ncols, B, out, kind, limit, d, vrows);
  count_launch();
  return launch_status();
}

// ---- workspaces ------------------------------------------------------------------------
struct SavedBlock {
  __nv_bfloat16 *ac, *gc, *m, *hin, *a, *g, *o;
  float *mu, *rstd;
};

struct LossGradPlan {
  // prologue
  float *e, *t, *r, *target;
  __nv_bfloat16 *xb, *cond_v, *cond_u, *dcond_u;
  // encoder
  __nv_bfloat16 *a_e, *g_e;
  float* lat;
  // v pass
  float* v;
  FwdScratch fs;
  // u pass
  float* xs;  // [(nb+1), B, Dp]
  SavedBlock blk[64];
  // tangent transients
  __nv_bfloat16 *gcd, *md, *hind, *gd;   // gcd: [B, nb*Cp]
  int64_t md_blk_stride = 0;             // > 0: one tangent-modulation buffer per block (hoisted schedule)
  __nv_bfloat16 *ac_all, *gc_all;        // [B, nb*Cp]: pre-activation / activation of the batched first modulation layer
  float* xd;
  // loss / backward
  float *row_loss, *g_x, *g_lat;
  __nv_bfloat16 *g_o, *g_a, *g_m, *g_ac, *g_latb, *g_ae;
  __nv_bfloat16 *g_o2, *g_a2, *g_m2;   // second set (concurrent schedule: block k's weight gradients read set k & 1 on a side stream)

  void plan(Arena& ar, const Dims& d, int64_t B) {
    e = ar.take<float>(B * d.Dp);
    target = ar.take<float>(B * d.Dp);
    t = ar.take<float>(B);
    r = ar.take<float>(B);
    xb = ar.take<__nv_bfloat16>(B * d.Dp);
    cond_v = ar.take<__nv_bfloat16>(B * d.Cp);
    cond_u = ar.take<__nv_bfloat16>(B * d.Cp);
    dcond_u = ar.take<__nv_bfloat16>(B * d.Cp);
    a_e = ar.take<__nv_bfloat16>(B * d.Hep);
    g_e = ar.take<__nv_bfloat16>(B * d.Hep);
    lat = ar.take<float>(B * d.Lp);
    v = ar.take<float>(B * d.Dp);
    fs.plan(ar, d, B, concurrent_rows(B, d));
    xs = ar.take<float>((int64_t)(d.nb + 1) * B * d.Dp);
    ac_all = ar.take<__nv_bfloat16>(B * d.Ca);
    gc_all = ar.take<__nv_bfloat16>(B * d.Ca);
    for (int k = 0; k < d.nb; ++k) {
      SavedBlock& sb = blk[k];
      sb.ac = ac_all ? ac_all + k * d.Cp : nullptr;   // column slices of the batched [B, nb*Cp] tensors (ld = Ca)
      sb.gc = gc_all ? gc_all + k * d.Cp : nullptr;
      sb.m = ar.take<__nv_bfloat16>(B * d.Mp);
      sb.hin = ar.take<__nv_bfloat16>(B * d.Ip);
      sb.a = ar.take<__nv_bfloat16>(B * d.Ip);
      sb.g = ar.take<__nv_bfloat16>(B * d.Ip);
      sb.o = ar.take<__nv_bfloat16>(B * d.Dp);
      sb.mu = ar.take<float>(B);
      sb.rstd = ar.take<float>(B);
    }
    gcd = ar.take<__nv_bfloat16>(B * d.Ca);
    md_blk_stride = concurrent_rows(B, d) ? B * d.Mp : 0;
    md = ar.take<__nv_bfloat16>(B * d.Mp * (md_blk_stride ? d.nb : 1));
    hind = ar.take<__nv_bfloat16>(B * d.Ip);
    gd = ar.take<__nv_bfloat16>(B * d.Ip);
    xd = ar.take<float>(B * d.Dp);
    row_loss = ar.take<float>(B);
    g_x = ar.take<float>(B * d.Dp);
    g_lat = ar.take<float>(B * d.Lp);
    g_o = ar.take<__nv_bfloat16>(B * d.Dp);
    g_a = ar.take<__nv_bfloat16>(B * d.Ip);
    g_m = ar.take<__nv_bfloat16>(B * d.Mp);
    g_ac = ar.take<__nv_bfloat16>(B * d.Ca);
    g_latb = ar.take<__nv_bfloat16>(B * d.Lp);
    g_ae = ar.take<__nv_bfloat16>(B * d.Hep);
    g_o2 = g_o; g_a2 = g_a; g_m2 = g_m;
    if (concurrent_rows(B, d)) {
      g_o2 = ar.take<__nv_bfloat16>(B * d.Dp);
      g_a2 = ar.take<__nv_bfloat16>(B * d.Ip);
      g_m2 = ar.take<__nv_bfloat16>(B * d.Mp);
    }
  }
};

struct ForwardPlan {
  float *x, *lat;
  __nv_bfloat16 *xb, *cond, *a_e, *g_e;
  FwdScratch fs;
  void plan(Arena& ar, const Dims& d, int64_t B) {
    x = ar.take<float>(B * d.Dp);
    lat = ar.take<float>(B * d.Lp);
    xb = ar.take<__nv_bfloat16>(B * d.Dp);
    cond = ar.take<__nv_bfloat16>(B * d.Cp);
    a_e = nullptr;
    g_e = ar.take<__nv_bfloat16>(B * d.Hep);
    fs.plan(ar, d, B);
  }
};

struct SamplePlan {
  float *x, *x2, *k1, *k2, *tmp, *lat;
  __nv_bfloat16* cond;
  FwdScratch fs;
  void plan(Arena& ar, const Dims& d, int64_t B) {
    x = ar.take<float>(B * d.Dp);
    x2 = ar.take<float>(B * d.Dp);
    k1 = ar.take<float>(B * d.Dp);
    k2 = ar.take<float>(B * d.Dp);
    tmp = ar.take<float>(B * d.Dp);
    lat = ar.take<float>(B * d.Lp);
    cond = ar.take<__nv_bfloat16>(B * d.Cp);
    fs.plan(ar, d, B);
  }
};

}  // namespace
}  // namespace mfac

using namespace mfac;

extern "C" {

size_t mfac_workspace_bytes(int32_t kind, const MfacMlpDims* dims, int64_t B) {
  Dims d;
  if (make_dims(dims, &d) != MFAC_SUCCESS || B <= 0 || d.nb > 64) return 0;
  if (d.nb < 1 && kind != MFAC_WS_FORWARD) return 0;
  Arena ar(nullptr, 0);
  if (kind == MFAC_WS_FORWARD) { ForwardPlan p; p.plan(ar, d, B); }
  else if (kind == MFAC_WS_LOSS_GRAD) { LossGradPlan p; p.plan(ar, d, B); }
  else if (kind == MFAC_WS_SAMPLE) { SamplePlan p; p.plan(ar, d, B); }
  else return 0;
  return ar.off + 256;
}

int mfac_mlp_encode(const MfacMlpDims* dims, const float* params, const void* shadow, const float* x, float* latents,
                    int64_t B, void* ws, size_t ws_bytes, void* stream) {
  (void)params;
  sweep_reset();
  Dims d;
  MFAC_OK(make_dims(dims, &d));
  if (!shadow || !x || !latents) return MFAC_ERR_NULL;
  if (B <= 0 || B > 0x7fffffff) return MFAC_ERR_BAD_SHAPE;
  if (!ws) return MFAC_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  ForwardPlan p;
  p.plan(ar, d, B);
  if (ar.overflow) return MFAC_ERR_WORKSPACE;
  Shadow sh(shadow, d);
  pad_rows_kernel<<<blocks_for(B * d.Dp, 256), 256, 0, s>>>(x, d.D, nullptr, p.xb, d.Dp, B);
  count_launch();
  MFAC_OK(encoder_pass(d, sh, p.xb, nullptr, p.g_e, p.lat, B, s));
  unpad_rows_kernel<<<blocks_for(B * d.L, 256), 256, 0, s>>>(p.lat, d.Lp, latents, d.L, B);
  count_launch();
  return launch_status();
}

int mfac_mlp_forward(const MfacMlpDims* dims, const float* params, const void* shadow, const float* x, const float* time,
                     const float* latents, float* out, int64_t B, void* ws, size_t ws_bytes, void* stream) {
  (void)params;
  sweep_reset();
  Dims d;
  MFAC_OK(make_dims(dims, &d));
  if (!shadow || !x || !time || !out) return MFAC_ERR_NULL;
  if (B <= 0 || B > 0x7fffffff || d.nb > 64 || d.nb < 1) return MFAC_ERR_BAD_SHAPE;
  if (!ws) return MFAC_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  ForwardPlan p;
  p.plan(ar, d, B);
  if (ar.overflow) return MFAC_ERR_WORKSPACE;
  Shadow sh(shadow, d);
  pad_rows_kernel<<<blocks_for(B * d.Dp, 256), 256, 0, s>>>(x, d.D, p.x, nullptr, d.Dp, B);
  count_launch();
  if (latents) {
    pad_rows_kernel<<<blocks_for(B * d.Lp, 256), 256, 0, s>>>(latents, d.L, p.lat, nullptr, d.Lp, B);
    count_launch();
  }
  cond_from_time_kernel<<<(unsigned)B, 128, 0, s>>>(time, p.cond, d);
  count_launch();
  MFAC_OK(forward_pass(d, sh, p.cond, latents ? p.lat : nullptr, p.x, B, p.fs, s));
  unpad_rows_kernel<<<blocks_for(B * d.D, 256), 256, 0, s>>>(p.x, d.Dp, out, d.D, B);
  count_launch();
  return launch_status();
}

}  // extern "C"

namespace mfac {
// Internal notification of the fused training step: grads[off, off + cnt) (nseg strided repetitions) will receive no further
// contribution once everything enqueued on `st` so far has run.  Block k's slice minus its first modulation layer is final
// right after block k's backward; the first-modulation-layer slices of all blocks (batched into one GEMM) and the encoder
// follow at the end.
struct SliceHook {
  virtual int final(int64_t off, int64_t cnt, int64_t stride, int nseg, cudaStream_t st) = 0;
  virtual ~SliceHook() = default;
};
bool comm_ready();
// raw audio instead of tokens: the step tokenises inside its prologue
struct AudioInput {
  const float* audio;   // [B, T]
  int64_t T;
  int N, hop;
};
int tokenize_prep_launch(const float* audio, int64_t T, int N, int hop, const PrepArgs& pa, const Dims& d, cudaStream_t stream);  // mdct.cu
int comm_allreduce_segments_f32(float* buf, int64_t count, int64_t stride, int nseg, cudaStream_t stream);

static int loss_grad_impl(const MfacMlpDims* dims, const MfacImfConfig* cfg, const void* shadow, const float* x, const float* e,
                          const float* t, const float* r, float* loss, float* grads, const MfacImfAux* aux, int64_t B, void* ws,
                          size_t ws_bytes, void* stream, SliceHook* hook, const AudioInput* audio = nullptr) {
  sweep_reset();
  Dims d;
  MFAC_OK(make_dims(dims, &d));
  if (!cfg || !shadow || (!x && !audio) || !loss || !grads) return MFAC_ERR_NULL;
  if (d.nb < 1) return MFAC_ERR_BAD_SHAPE;
  if (audio) {
    if (!audio->audio) return MFAC_ERR_NULL;
    if (audio->T <= 0 || audio->N <= 0 || audio->hop <= 0) return MFAC_ERR_BAD_SHAPE;
    const int64_t nf = audio->T < audio->N ? 1 : (audio->T - audio->N) / audio->hop + 1;
    if (nf * audio->N != d.D) return MFAC_ERR_BAD_SHAPE;   // the tokens of a clip are one model row
  }
  if (B <= 0 || B > 0x7fffffff || d.nb > 64 || d.nb < 1) return MFAC_ERR_BAD_SHAPE;
  if ((t == nullptr) != (r == nullptr)) return MFAC_ERR_NULL;
  if (!ws) return MFAC_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  LossGradPlan p;
  p.plan(ar, d, B);
  if (ar.overflow) return MFAC_ERR_WORKSPACE;
  Shadow sh(shadow, d);
  const int M = (int)B;
  const float inv_nb = 1.0f / (float)d.nb;

  // ---- prologue: (e, t, r), z_t, cond rows
  if (cfg->method < MFAC_LOSS_IMPROVED_MEAN_FLOW || cfg->method > MFAC_LOSS_FLOW_MATCHING) return MFAC_ERR_UNSUPPORTED;
  const bool need_v = cfg->method == MFAC_LOSS_IMPROVED_MEAN_FLOW;   // tangent seed = the network's own velocity
  const bool tangent = cfg->method != MFAC_LOSS_FLOW_MATCHING;       // flow matching has no JVP
  // Rows [0, h) have r == t (sample_tr's rule, utils.py:41-44): there the u pass sees exactly the input and the (t, 0)
  // conditioning of the v pass, so ONE saved primal pass serves as both.  h comes from the library's own draw, or from
  // the caller's promise for explicit (t, r).
  int64_t h = 0;
  if (need_v && cfg->rows_r_equals_t >= 0)
    h = t ? (cfg->rows_r_equals_t < B ? cfg->rows_r_equals_t : B) : (int64_t)((float)B * cfg->data_proportion);
  if (h < 0) h = 0;
  if (h > B) h = B;
  const bool share = need_v && h * 4 >= B;
  if (!share) h = 0;
  const int Mu = (int)(B - h);   // rows whose u pass differs from their v pass
  // Concurrent schedule for small batches (a 128-row GEMM fills 10 of 148 SMs): the three forward evaluations are
  // independent chains -- v on the rows with r != t (unsaved), the saved primal pass with the (t, 0) conditioning on the rows with
  // r == t, the saved primal u pass on the rows with r != t -- and run side by side on the library's side streams; the tangent
  // pass follows on its own (it needs v and the u pass's activations); in the backward the weight gradients, bias column sums
  // and the modulation-MLP input gradient leave the critical chain for a side stream.  Same kernels, same results.
  ForkCtx* fc = concurrent_rows(B, d) ? fork_ctx() : nullptr;
  const bool conc = fc != nullptr;
  const bool conc_fwd = conc && need_v;
  // z_t -> xs[0] (u pass) and, for improved mean flow without sharing, v (v pass, in place); mean flow seeds the tangent
  // with e - x
  PrepArgs pa{x, e, t, r, (aux && aux->e) ? p.e : nullptr, p.target, need_v && (!share || conc_fwd) ? p.v : nullptr, p.xs,
              cfg->method == MFAC_LOSS_MEAN_FLOW ? p.v : nullptr,
              p.xb, p.t, p.r, need_v ? p.cond_v : nullptr, p.cond_u, p.dcond_u, *cfg, B};
  phase_mark(0, s);
  int fused_tok = MFAC_ERR_UNSUPPORTED;
  if (audio) {
    // tokenise inside the prologue: the MDCT kernel's store stage IS the prologue, no token tensor goes to HBM
    fused_tok = tokenize_prep_launch(audio->audio, audio->T, audio->N, audio->hop, pa, d, s);
    if (fused_tok != MFAC_SUCCESS && fused_tok != MFAC_ERR_UNSUPPORTED) return fused_tok;
    if (fused_tok == MFAC_ERR_UNSUPPORTED) {
      // geometry outside the fused kernel (long clips, other windows): tokens go through scratch (g_x is free until the loss)
      MFAC_OK(mfac_mdct_f32(audio->audio, p.g_x, B, audio->T, audio->N, audio->hop, stream));
      pa.x = p.g_x;
    }
  }
  if (fused_tok != MFAC_SUCCESS) {
    launch_pdl(imf_prep_kernel, dim3((unsigned)B), dim3(ROW_THREADS), 0, s, pa, d);
    count_launch();
  }
  // ---- latents = encode(x)
  MFAC_OK(encoder_pass(d, sh, p.xb, p.a_e, p.g_e, p.lat, B, s));
  phase_mark(1, s);
  // one block of the saved primal pass over rows [r0, r0 + rows): activations go to the per-block buffers the tangent
  // pass and the backward read
  auto primal_mod = [&](int k, int64_t r0, int rows, cudaStream_t st) -> int {
    const __nv_bfloat16* w = sh.w + k * d.s_blk_stride;
    const float* bias = sh.b + k * d.b_blk_stride;
    SavedBlock& sb = p.blk[k];
    return gemm_linear_bf16(sb.gc + r0 * d.Ca, d.Ca, w + d.s_c2w, rows, d.Mp, d.Cp, bias + d.b_c2, sb.m + r0 * d.Mp, d.Mp, st);
  };
  // keep: rows (relative to r0) whose pre-activation a and block output o are stored for the tangent pass / backward
  auto primal_mlp = [&](int k, int64_t r0, int rows, int keep, cudaStream_t st) -> int {
    const __nv_bfloat16* w = sh.w + k * d.s_blk_stride;
    const float* bias = sh.b + k * d.b_blk_stride;
    SavedBlock& sb = p.blk[k];
    float* x_in = p.xs + (int64_t)k * B * d.Dp + r0 * d.Dp;
    float* x_out = p.xs + (int64_t)(k + 1) * B * d.Dp + r0 * d.Dp;
    MFAC_OK(gemm_fwd(sb.hin + r0 * d.Ip, d.Ip, w + d.s_m1w, rows, d.Ip, d.Ip,
                     EpiBiasGelu{bias + d.b_m1, sb.g + r0 * d.Ip, sb.a + r0 * d.Ip, d.Ip, keep}, st));
    return gemm_fwd(sb.g + r0 * d.Ip, d.Ip, w + d.s_m2w, rows, d.Dp, d.Ip,
                    EpiBlockOut{bias + d.b_m2, sb.m + r0 * d.Mp, x_in, x_out, sb.o + r0 * d.Dp, d.Mp, d.Dp, 2 * d.Ip, inv_nb, keep}, st);
  };
  // the saved primal pass (all blocks) over rows [r0, r0 + rows) with conditioning rows `cond`
  auto saved_pass = [&](const __nv_bfloat16* cond, int64_t r0, int rows, int keep, cudaStream_t st, Hoist* hz) -> int {
    MFAC_OK(gemm_bias_gelu(cond + r0 * d.Cp, d.Cp, sh.w + d.s_c1all, rows, d.Ca, d.Cp, sh.b + d.b_c1all, p.gc_all + r0 * d.Ca,
                           p.ac_all + r0 * d.Ca, d.Ca, st));
    if (hz) MFAC_OK(hz->run(st, d.nb, [&](int k, cudaStream_t sm) { return primal_mod(k, r0, rows, sm); }));
    for (int k = 0; k < d.nb; ++k) {
      SavedBlock& sb = p.blk[k];
      if (hz) MFAC_OK(hz->wait(st, k));
      else MFAC_OK(primal_mod(k, r0, rows, st));
      LnModArgs la{p.lat + r0 * d.Lp, p.xs + (int64_t)k * B * d.Dp + r0 * d.Dp, sb.m + r0 * d.Mp, sb.hin + r0 * d.Ip, nullptr, nullptr,
                   nullptr, sb.mu + r0, sb.rstd + r0, d.Mp};
      MFAC_OK(lnmod(false, la, d, rows, st));
      MFAC_OK(primal_mlp(k, r0, rows, keep, st));
    }
    return MFAC_SUCCESS;
  };
  // one block of the tangent pass over rows [h, B); with_primal: the fused AdaLN kernel also produces the primal's hin / statistics
  // tangent of block k's modulation (no bias: d/dt of a constant)
  auto tangent_mod = [&](int k, cudaStream_t st) -> int {
    return gemm_linear_bf16(p.gcd + h * d.Ca + k * d.Cp, d.Ca, sh.w + k * d.s_blk_stride + d.s_c2w, Mu, d.Mp, d.Cp, nullptr,
                            p.md + k * p.md_blk_stride + h * d.Mp, d.Mp, st);
  };
  auto tangent_block = [&](int k, bool with_primal, cudaStream_t st, Hoist* hz) -> int {
    const __nv_bfloat16* w = sh.w + k * d.s_blk_stride;
    SavedBlock& sb = p.blk[k];
    float* x_in = p.xs + (int64_t)k * B * d.Dp + h * d.Dp;
    const float* xd_in = (k == 0 ? p.v : p.xd) + h * d.Dp;
    const __nv_bfloat16* md = p.md + k * p.md_blk_stride + h * d.Mp;
    if (hz) MFAC_OK(hz->wait(st, k));
    else MFAC_OK(tangent_mod(k, st));
    LnModArgs la{p.lat + h * d.Lp, x_in, sb.m + h * d.Mp, with_primal ? sb.hin + h * d.Ip : nullptr, xd_in, md,
                 p.hind + h * d.Ip, with_primal ? sb.mu + h : nullptr, with_primal ? sb.rstd + h : nullptr, d.Mp};
    MFAC_OK(lnmod(true, la, d, Mu, st));
    // the tangent GEMMs read the primal pre-activation a and block output o: primal first
    if (with_primal) MFAC_OK(primal_mlp(k, h, Mu, Mu, st));
    MFAC_OK(gemm_fwd(p.hind + h * d.Ip, d.Ip, w + d.s_m1w, Mu, d.Ip, d.Ip, EpiMulDgelu{sb.a + h * d.Ip, p.gd + h * d.Ip, d.Ip}, st));
    return gemm_fwd(p.gd + h * d.Ip, d.Ip, w + d.s_m2w, Mu, d.Dp, d.Ip,
                    EpiBlockOutTangent{sb.m + h * d.Mp, md, sb.o + h * d.Dp, xd_in, p.xd + h * d.Dp, d.Mp, d.Dp,
                                       2 * d.Ip, inv_nb}, st);
  };
  if (conc_fwd) {
    // ---- two independent forward chains side by side, then the tangent chain.  The saved pass of the rows with r == t (their u
    // IS v) and the u pass of the other rows are ONE chain over all rows: cond_u = embed(t, t - r) equals cond_v = embed(t, 0) bit
    // for bit where t - r == 0, so the rows [0, h) of cond_u already hold the (t, 0) conditioning.
    cudaStream_t sV = fc->side[0];
    Hoist hzU, hzV, hzT;   // every chain's modulation GEMMs run ahead of it on a side stream of their own
    hzU.fc = hzV.fc = hzT.fc = fc;
    hzU.sm = hzT.sm = fc->side[3]; hzV.sm = fc->side[4];
    MFAC_OK(stream_after(fc, s, sV));
    if (Mu > 0)   // v = f(z, [t, 0], lat) on the rows with r != t (in place on p.v, nothing kept)
      MFAC_OK(forward_pass(d, sh, p.cond_v + h * d.Cp, p.lat + h * d.Lp, p.v + h * d.Dp, Mu, p.fs, sV, false, &hzV));
    MFAC_OK(saved_pass(p.cond_u, 0, M, M, s, &hzU));
    if (Mu > 0 && tangent) {
      // tangent of the first modulation layer (needs the primal's pre-activation ac_all, which side[3] already waits for) and
      // the tangent modulation GEMMs of all blocks, behind the primal's modulation GEMMs on the same side stream
      MFAC_OK(gemm_fwd(p.dcond_u + h * d.Cp, d.Cp, sh.w + d.s_c1all, Mu, d.Ca, d.Cp,
                       EpiMulDgelu{p.ac_all + h * d.Ca, p.gcd + h * d.Ca, d.Ca}, hzT.sm));
      MFAC_OK(hzT.run(hzT.sm, d.nb, tangent_mod));
    }
    MFAC_OK(stream_after(fc, sV, s));
    if (h > 0 && aux && aux->v)
      MFAC_CUDA_OK(cudaMemcpyAsync(p.v, p.xs + (int64_t)d.nb * B * d.Dp, (size_t)h * d.Dp * 4, cudaMemcpyDeviceToDevice, s));
    phase_mark(2, s);
    for (int k = 0; k < d.nb && Mu > 0 && tangent; ++k) MFAC_OK(tangent_block(k, false, s, &hzT));
  } else {
  // ---- v = f(z, [t, 0], lat)
  if (share) {
    // saved primal pass over ALL rows with the (t, 0) conditioning: v for every row, and already u (with every
    // activation the tangent pass and the backward need) for rows [0, h)
    MFAC_OK(saved_pass(p.cond_v, 0, M, (int)h, s, nullptr));   // rows [h, B) get their a / o from the u pass below
    // v is the tangent seed of the rows that still get a u pass (and an optional test output for all rows)
    const int64_t v0 = (aux && aux->v) ? 0 : h;
    if (v0 < B)
      MFAC_CUDA_OK(cudaMemcpyAsync(p.v + v0 * d.Dp, p.xs + (int64_t)d.nb * B * d.Dp + v0 * d.Dp, (size_t)(B - v0) * d.Dp * 4,
                                   cudaMemcpyDeviceToDevice, s));
  } else if (need_v) {
    MFAC_OK(forward_pass(d, sh, p.cond_v, p.lat, p.v, B, p.fs, s));
  }
  // ---- (u, du/dt) = jvp(f, (z, [t, t-r]), (v, [1, 1]))  on rows [h, B)
  // Rows [0, h) are finished: their u is the shared pass's output, and their du/dt is never used -- the loss multiplies it
  // by (t - r), which is exactly 0 there (v_pred = u + 0 * du/dt = u bit for bit).  Everything below runs on the Mu rows
  // whose u evaluation differs from their v evaluation.
  // first modulation layer of all blocks, primal and tangent, in two GEMMs
  if (Mu > 0) {
    MFAC_OK(gemm_bias_gelu(p.cond_u + h * d.Cp, d.Cp, sh.w + d.s_c1all, Mu, d.Ca, d.Cp, sh.b + d.b_c1all, p.gc_all + h * d.Ca,
                           p.ac_all + h * d.Ca, d.Ca, s));
    if (tangent)
      MFAC_OK(gemm_fwd(p.dcond_u + h * d.Cp, d.Cp, sh.w + d.s_c1all, Mu, d.Ca, d.Cp,
                       EpiMulDgelu{p.ac_all + h * d.Ca, p.gcd + h * d.Ca, d.Ca}, s));
  }
  for (int k = 0; k < d.nb && Mu > 0; ++k) {
    if (tangent) {
      MFAC_OK(primal_mod(k, h, Mu, s));
      MFAC_OK(tangent_block(k, true, s, nullptr));
    } else {
      SavedBlock& sb = p.blk[k];
      MFAC_OK(primal_mod(k, h, Mu, s));
      LnModArgs la{p.lat + h * d.Lp, p.xs + (int64_t)k * B * d.Dp + h * d.Dp, sb.m + h * d.Mp, sb.hin + h * d.Ip, nullptr, nullptr,
                   nullptr, sb.mu + h, sb.rstd + h, d.Mp};
      MFAC_OK(lnmod(false, la, d, Mu, s));
      MFAC_OK(primal_mlp(k, h, Mu, Mu, s));
    }
  }
  }
  if (h > 0 && tangent && aux && aux->dudt) MFAC_CUDA_OK(cudaMemsetAsync(p.xd, 0, (size_t)h * d.Dp * 4, s));  // reported as 0
  const float* u = p.xs + (int64_t)d.nb * B * d.Dp;
  phase_mark(3, s);
  // ---- loss and its seed gradient
  LossArgs lo{u, tangent ? p.xd : nullptr, p.target, p.t, p.r, p.g_x, p.row_loss, aux ? aux->per_example : nullptr, *cfg, B};
  launch_pdl(imf_loss_kernel, dim3((unsigned)B), dim3(ROW_THREADS), (size_t)d.Dp * 4, s, lo, d);
  count_launch();
  launch_pdl(sum_rows_kernel, dim3(1), dim3(1024), 0, s, (const float*)p.row_loss, B, loss);
  count_launch();
  // ---- backward through the primal u rows (weight gradients accumulate split-K partials atomically)
  MFAC_CUDA_OK(cudaMemsetAsync(grads, 0, (size_t)d.total * 4, s));
  MFAC_CUDA_OK(cudaMemsetAsync(p.g_lat, 0, (size_t)B * d.Lp * 4, s));
  phase_mark(4, s);
  // Concurrent schedule: the critical chain (block-output backward -> two dX GEMMs -> LayerNorm backward) stays on `s`; the
  // weight gradients, the remaining bias column sums and the modulation-MLP input gradient of block k run on a side stream from
  // transient set k & 1, which the critical chain only overwrites again two blocks later (after that block's side work).
  // The side work itself is three independent chains (they read g_o / g_a / g_m and write disjoint gradient slices): at small
  // batch one side stream running five ~15 us kernels per block was the backward's critical path (80 us per block against 29 us
  // for the chain on `s`), so the two big weight gradients get streams of their own and join sW before the block's slice is final.
  cudaStream_t sW = conc ? fc->side[2] : s, sW1 = conc ? fc->side[3] : s, sW2 = conc ? fc->side[4] : s;
  cudaEvent_t side_done[2] = {nullptr, nullptr}, g_lat_done = nullptr;
  for (int k = d.nb - 1; k >= 0; --k) {
    const __nv_bfloat16* w = sh.w + k * d.s_blk_stride;
    SavedBlock& sb = p.blk[k];
    float* gk = grads + (int64_t)k * d.blk_stride;
    const float* x_in = p.xs + (int64_t)k * B * d.Dp;
    const int par = conc ? (k & 1) : 0;
    __nv_bfloat16* g_o = par ? p.g_o2 : p.g_o;
    __nv_bfloat16* g_a = par ? p.g_a2 : p.g_a;
    __nv_bfloat16* g_m = par ? p.g_m2 : p.g_m;
    if (conc && side_done[par]) MFAC_CUDA_OK(cudaStreamWaitEvent(s, side_done[par], 0));   // set `par` is free again
    // g_o, g_s2 and both of their bias-gradient column sums (db2, the s2 third of dbc2) in one pass
    const int vrows = colsum_vrows(B);
    launch_pdl(bwd_block_out_vec_kernel, dim3(ceil_div(d.Dp, 256), (unsigned)ceil_div<int64_t>(B, vrows)), dim3(256), 0, s,
               (const float*)p.g_x, (const __nv_bfloat16*)sb.m, (const __nv_bfloat16*)sb.o, g_o, g_m, gk + d.o_m2b, gk + d.o_c2b, d, B,
               sweep_next(), vrows);
    count_launch();
    if (conc) MFAC_OK(stream_after(fc, s, sW));
    MFAC_OK(gemm_dw(sb.g, d.Ip, g_o, d.Dp, d.Ip, d.Dp, M, EpiGradStore{gk + d.o_m2w, d.D, MAP_CM, 0, MAP_ID, d.D, 1, d}, sW));
    // Unpadded geometries: the dX epilogues that write g_a and g_shift also accumulate their column sums (db1 and the shift
    // third of dbc2) -- no separate pass over those tensors.
    const bool fused_colsum = d.I == d.Ip;
    if (fused_colsum) {
      EpiMulDgeluColsum ep;
      ep.a = sb.a; ep.out = g_a; ep.ld = d.Ip; ep.colsum = gk + d.o_m1b;
      MFAC_OK(gemm_dx(g_o, d.Dp, w + d.s_m2w, M, d.Ip, d.Dp, ep, s));
    } else {
      MFAC_OK(gemm_dx(g_o, d.Dp, w + d.s_m2w, M, d.Ip, d.Dp, EpiMulDgelu{sb.a, g_a, d.Ip}, s));
    }
    if (conc) MFAC_OK(stream_after(fc, s, sW1));
    MFAC_OK(gemm_dw(sb.hin, d.Ip, g_a, d.Ip, d.Ip, d.Ip, M, EpiGradStore{gk + d.o_m1w, d.I, MAP_CM, 0, MAP_CM, 0, 1, d}, sW1));
    if (!fused_colsum) MFAC_OK(colsum(g_a, d.Ip, B, gk + d.o_m1b, MAP_CM, 0, d, sW1));
    // g_hin = g_a W1^T goes out in bf16 straight into g_m[:, Ip:2Ip]: it IS the shift gradient (hin = (1+s1) n + shift)
    if (fused_colsum) {
      EpiLinearBf16Colsum ep;
      ep.bias = nullptr; ep.out = g_m + d.Ip; ep.ld = d.Mp; ep.colsum = gk + d.o_c2b + d.I;
      MFAC_OK(gemm_dx(g_a, d.Ip, w + d.s_m1w, M, d.Ip, d.Ip, ep, s));
    } else {
      MFAC_OK(gemm_dx(g_a, d.Ip, w + d.s_m1w, M, d.Ip, d.Ip, EpiLinearBf16{nullptr, g_m + d.Ip, d.Mp}, s));
    }
    LnBwdArgs lb{p.lat, x_in, sb.mu, sb.rstd, sb.m, g_m, p.g_lat, p.g_x};
    MFAC_OK(ln_bwd(lb, d, B, s));
    if (conc && k == 0) {   // g_lat is complete: the encoder backward may start (below, on streams of its own)
      g_lat_done = fc->next_event();
      MFAC_CUDA_OK(cudaEventRecord(g_lat_done, s));
    }
    phase_mark(5 + (d.nb - 1 - k), s);
    if (conc) {
      MFAC_OK(stream_after(fc, s, sW));
      MFAC_OK(stream_after(fc, s, sW2));
    }
    MFAC_OK(gemm_dw(sb.gc, d.Ca, g_m, d.Mp, d.Cp, d.Mp, M,
                    EpiGradStore{gk + d.o_c2w, 2 * d.I + d.D, MAP_ID, d.C, MAP_MM, 0, 1, d}, sW2));
    // s1 third (and the shift third where it was not fused above)
    MFAC_OK(colsum(g_m, d.Mp, B, gk + d.o_c2b, MAP_MM, 0, d, sW, fused_colsum ? d.Ip : 2 * d.Ip));
    MFAC_OK(gemm_dx(g_m, d.Mp, w + d.s_c2w, M, d.Cp, d.Mp, EpiMulDgelu{sb.ac, p.g_ac + k * d.Cp, d.Ca}, sW));
    if (conc) {
      MFAC_OK(stream_after(fc, sW1, sW));
      MFAC_OK(stream_after(fc, sW2, sW));
      side_done[par] = fc->next_event();
      MFAC_CUDA_OK(cudaEventRecord(side_done[par], sW));
    }
    // block k's slice (all of it but the first modulation layer) is final once the side work above has run
    if (hook) MFAC_OK(hook->final((int64_t)k * d.blk_stride + d.o_c2b, d.blk_stride - d.o_c2b, 0, 1, sW));
    if (k == 0) {
      if (conc) MFAC_OK(stream_after(fc, sW, s));   // every block's side work (in order on sW) is done
      // first modulation layer, all blocks at once: dW = cond^T @ g_ac_all, db = column sums
      MFAC_OK(gemm_dw(p.cond_u, d.Cp, p.g_ac, d.Ca, d.Cp, d.Ca, M, EpiGradStoreC1{grads, d}, s));
      MFAC_OK(colsum(p.g_ac, d.Ca, B, grads, MAP_C1ALL, 0, d, s));
    }
    // with the batched cond1 gradients every block's slice is only final after block 0; buckets are announced then
    if (aux && aux->grad_ready && k == 0)
      for (int kk = d.nb - 1; kk >= 0; --kk) aux->grad_ready(aux->grad_ready_user, (int64_t)kk * d.blk_stride, d.blk_stride);
  }
  // ---- encoder backward (concurrent schedule: on two side streams beside the first-modulation-layer gradients above; it only
  // needs g_lat, which block 0's LayerNorm backward completed)
  cudaStream_t sE = conc ? fc->side[5] : s, sE2 = conc ? fc->side[0] : s;
  if (conc) MFAC_CUDA_OK(cudaStreamWaitEvent(sE, g_lat_done, 0));
  f32_to_bf16_kernel<<<blocks_for(B * d.Lp, 256), 256, 0, sE>>>(p.g_lat, p.g_latb, B * d.Lp);
  count_launch();
  if (conc) MFAC_OK(stream_after(fc, sE, sE2));
  MFAC_OK(gemm_dw(p.g_e, d.Hep, p.g_latb, d.Lp, d.Hep, d.Lp, M, EpiGradStore{grads + d.o_e2w, d.L, MAP_ID, d.He, MAP_ID, d.L, 1, d}, sE2));
  MFAC_OK(colsum(p.g_latb, d.Lp, B, grads + d.o_e2b, MAP_ID, d.L, d, sE2));
  MFAC_OK(gemm_dx(p.g_latb, d.Lp, sh.w + d.s_e2w, M, d.Hep, d.Lp, EpiMulDgelu{p.a_e, p.g_ae, d.Hep}, sE));
  MFAC_OK(gemm_dw(p.xb, d.Dp, p.g_ae, d.Hep, d.Dp, d.Hep, M, EpiGradStore{grads + d.o_e1w, d.He, MAP_ID, d.D, MAP_ID, d.He, 1, d}, sE));
  MFAC_OK(colsum(p.g_ae, d.Hep, B, grads + d.o_e1b, MAP_ID, d.He, d, sE));
  if (conc) {
    MFAC_OK(stream_after(fc, sE2, s));
    MFAC_OK(stream_after(fc, sE, s));
  }
  if (aux && aux->grad_ready)
    aux->grad_ready(aux->grad_ready_user, (int64_t)d.nb * d.blk_stride, d.total - (int64_t)d.nb * d.blk_stride);
  phase_mark(90, s);
  if (hook) {   // the nb first-modulation-layer slices [k * stride, + o_c2b) and the encoder
    MFAC_OK(hook->final(0, d.o_c2b, d.blk_stride, d.nb, s));
    MFAC_OK(hook->final((int64_t)d.nb * d.blk_stride, d.total - (int64_t)d.nb * d.blk_stride, 0, 1, s));
  }
  // ---- optional intermediates for parity tests
  if (aux) {
    const unsigned nbk = blocks_for(B * d.D, 256);
    if (aux->v && need_v) { unpad_rows_kernel<<<nbk, 256, 0, s>>>(p.v, d.Dp, aux->v, d.D, B); count_launch(); }
    if (aux->u) { unpad_rows_kernel<<<nbk, 256, 0, s>>>(u, d.Dp, aux->u, d.D, B); count_launch(); }
    if (aux->dudt && tangent) { unpad_rows_kernel<<<nbk, 256, 0, s>>>(p.xd, d.Dp, aux->dudt, d.D, B); count_launch(); }
    if (aux->e) { unpad_rows_kernel<<<nbk, 256, 0, s>>>(p.e, d.Dp, aux->e, d.D, B); count_launch(); }
    if (aux->t) MFAC_CUDA_OK(cudaMemcpyAsync(aux->t, p.t, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
    if (aux->r) MFAC_CUDA_OK(cudaMemcpyAsync(aux->r, p.r, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
  }
  return launch_status();
}
}  // namespace mfac

extern "C" {

int mfac_imf_loss_grad(const MfacMlpDims* dims, const MfacImfConfig* cfg, const float* params, const void* shadow,
                       const float* x, const float* e, const float* t, const float* r, float* loss, float* grads,
                       const MfacImfAux* aux, int64_t B, void* ws, size_t ws_bytes, void* stream) {
  (void)params;
  return loss_grad_impl(dims, cfg, shadow, x, e, t, r, loss, grads, aux, B, ws, ws_bytes, stream, nullptr);
}

int mfac_uses_concurrent_schedule(const MfacMlpDims* dims, int64_t B) {
  Dims d;
  if (make_dims(dims, &d) != MFAC_SUCCESS || B <= 0) return 0;
  return concurrent_rows(B, d) ? 1 : 0;
}

int mfac_imf_loss_grad_audio(const MfacMlpDims* dims, const MfacImfConfig* cfg, const float* params, const void* shadow,
                             const float* audio, int64_t T, int32_t window_size, int32_t hop_size, const float* e, const float* t,
                             const float* r, float* loss, float* grads, const MfacImfAux* aux, int64_t B, void* ws,
                             size_t ws_bytes, void* stream) {
  (void)params;
  AudioInput in{audio, T, window_size, hop_size};
  return loss_grad_impl(dims, cfg, shadow, nullptr, e, t, r, loss, grads, aux, B, ws, ws_bytes, stream, nullptr, &in);
}

namespace {
// AdamW (and, for world > 1, the gradient all-reduce before it) applied slice by slice as the backward finalises them
struct TrainHook : SliceHook {
  Dims d;
  AdamWLaunch h;
  float* grads;
  int world;
  // Large batches keep the machine busy on their own (single-stream schedule): there the exchange and the update stay ONE
  // all-reduce and ONE pass over the parameters after the backward (nine smaller collectives on the compute stream only add
  // latency, and NCCL kernels beside the persistent one-CTA-per-SM GEMMs cost more than the exchange).
  bool deferred;
  // world > 1, concurrent schedule: `bucket_blocks` consecutive blocks are exchanged as ONE NCCL group (measured on 2 x B200 at
  // 128 rows: one 14 MB collective per block does not reach NVLink bandwidth: 1.74 ms / step against 1.44 ms for a single
  // 113 MB bucket after the backward; two-block buckets on a stream of their own: 1.40 ms).
  int bucket_blocks = 2;
  int pending = 0;
  // Concurrent schedule: the exchange and the update of a finalised slice run on a stream of their own (side[1], idle during the
  // backward) behind the stream the slice became final on -- not ON it: that stream carries the weight-gradient GEMMs, and a
  // collective per bucket in front of them made it the critical path (2 x B200, 128 rows: 1.61 ms against 1.44 ms).
  ForkCtx* fc = nullptr;
  cudaStream_t on(cudaStream_t st) {
    if (!fc) return st;
    if (stream_after(fc, st, fc->side[1]) != MFAC_SUCCESS) return st;
    return fc->side[1];
  }
  int flush(int k_lo, int nblk, cudaStream_t st) {
    if (nblk <= 0) return MFAC_SUCCESS;
    // one NCCL group = one fused launch over the blocks' slices; their first-modulation-layer regions stay out of it (the batched
    // GEMM that fills them runs later on the main stream and must not race with a collective over the same words)
    const int64_t off = (int64_t)k_lo * d.blk_stride + d.o_c2b, cnt = d.blk_stride - d.o_c2b;
    if (world > 1) MFAC_OK(comm_allreduce_segments_f32(grads + off, cnt, d.blk_stride, nblk, st));
    return adamw_segments(d, h, off, cnt, d.blk_stride, nblk, st);
  }
  int final(int64_t off, int64_t cnt, int64_t stride, int nseg, cudaStream_t st) override {
    if (deferred) {
      if (off != (int64_t)d.nb * d.blk_stride) return MFAC_SUCCESS;   // the encoder slice is announced last
      if (world > 1) MFAC_OK(comm_allreduce_segments_f32(grads, d.total, 0, 1, st));
      return adamw_segments(d, h, 0, d.total, 0, 1, st);
    }
    if (nseg == 1 && off < (int64_t)d.nb * d.blk_stride) {   // block k (announced nb-1 .. 0)
      const int k = (int)(off / d.blk_stride);
      ++pending;
      if (pending >= (world > 1 ? bucket_blocks : 1) || k == 0) {
        MFAC_OK(flush(k, pending, on(st)));
        pending = 0;
      }
      return MFAC_SUCCESS;
    }
    st = on(st);
    if (world > 1) MFAC_OK(comm_allreduce_segments_f32(grads + off, cnt, stride, nseg, st));
    return adamw_segments(d, h, off, cnt, stride, nseg, st);
  }
};
}  // namespace

namespace {
int train_step_impl(const MfacMlpDims* dims, const MfacImfConfig* cfg, const MfacAdamWConfig* opt, float* params, void* shadow,
                    float* mu, float* nu, int64_t count, uint64_t* count_dev, float* scratch_dev, const float* x,
                    const AudioInput* audio, const float* e, const float* t, const float* r, float* loss, float* grads,
                    const MfacImfAux* aux, int64_t B, int32_t world, void* ws, size_t ws_bytes, void* stream) {
  if (!opt || !params || !shadow || !mu || !nu || !grads) return MFAC_ERR_NULL;
  if (count < 0 || world < 1) return MFAC_ERR_BAD_SHAPE;
  if (world > 1 && !comm_ready()) return MFAC_ERR_NCCL;
  TrainHook hook;
  MFAC_OK(make_dims(dims, &hook.d));
  hook.h = AdamWLaunch{params, grads, mu, nu, shadow, opt->lr, opt->b1, opt->b2, opt->eps, opt->weight_decay, 1.0f / (float)world,
                       0.f, 0.f, nullptr};
  hook.grads = grads;
  hook.world = world;
  hook.deferred = !(concurrent_rows(B, hook.d) && fork_ctx() != nullptr);
  hook.fc = hook.deferred ? nullptr : fork_ctx();
  if (const char* ev = getenv("MFAC_DP_BUCKET_BLOCKS")) hook.bucket_blocks = atoi(ev) > 0 ? atoi(ev) : 2;
  // bias correction first (host value, or the device counter read and advanced once per step for graph replay)
  MFAC_OK(adamw_prepare(hook.h, count, count_dev, scratch_dev, (cudaStream_t)stream, /*advance=*/false));
  MFAC_OK(loss_grad_impl(dims, cfg, shadow, x, e, t, r, loss, grads, aux, B, ws, ws_bytes, stream, &hook, audio));
  if (hook.fc) MFAC_OK(stream_after(hook.fc, hook.fc->side[1], (cudaStream_t)stream));   // exchange / update stream rejoins
  // the prologue read the counter as the RNG step (MfacImfConfig.step_dev may be the same word): advance it last
  if (count_dev) MFAC_OK(adamw_advance(count_dev, (cudaStream_t)stream));
  phase_mark(99, (cudaStream_t)stream);
  return MFAC_SUCCESS;
}
}  // namespace

int mfac_imf_train_step(const MfacMlpDims* dims, const MfacImfConfig* cfg, const MfacAdamWConfig* opt, float* params, void* shadow,
                        float* mu, float* nu, int64_t count, uint64_t* count_dev, float* scratch_dev, const float* x,
                        const float* e, const float* t, const float* r, float* loss, float* grads, const MfacImfAux* aux,
                        int64_t B, int32_t world, void* ws, size_t ws_bytes, void* stream) {
  return train_step_impl(dims, cfg, opt, params, shadow, mu, nu, count, count_dev, scratch_dev, x, nullptr, e, t, r, loss, grads, aux,
                         B, world, ws, ws_bytes, stream);
}

int mfac_imf_train_step_audio(const MfacMlpDims* dims, const MfacImfConfig* cfg, const MfacAdamWConfig* opt, float* params,
                              void* shadow, float* mu, float* nu, int64_t count, uint64_t* count_dev, float* scratch_dev,
                              const float* audio, int64_t T, int32_t window_size, int32_t hop_size, const float* e, const float* t,
                              const float* r, float* loss, float* grads, const MfacImfAux* aux, int64_t B, int32_t world, void* ws,
                              size_t ws_bytes, void* stream) {
  AudioInput in{audio, T, window_size, hop_size};
  return train_step_impl(dims, cfg, opt, params, shadow, mu, nu, count, count_dev, scratch_dev, nullptr, &in, e, t, r, loss, grads,
                         aux, B, world, ws, ws_bytes, stream);
}

int mfac_sample(const MfacMlpDims* dims, const float* params, const void* shadow, const float* latents, const float* noise,
                int32_t mode, int32_t n_steps, float guidance_scale, uint64_t seed, float* out, int64_t B, void* ws,
                size_t ws_bytes, void* stream) {
  (void)params;
  sweep_reset();
  Dims d;
  MFAC_OK(make_dims(dims, &d));
  if (!shadow || !latents || !out) return MFAC_ERR_NULL;
  if (B <= 0 || B > 0x7fffffff || n_steps <= 0 || d.nb > 64 || d.nb < 1) return MFAC_ERR_BAD_SHAPE;
  if (mode != MFAC_SAMPLE_HEUN && mode != MFAC_SAMPLE_MF) return MFAC_ERR_UNSUPPORTED;
  if (!ws) return MFAC_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  SamplePlan p;
  p.plan(ar, d, B);
  if (ar.overflow) return MFAC_ERR_WORKSPACE;
  Shadow sh(shadow, d);
  const int64_t n = B * d.Dp;
  const unsigned nblk = blocks_for(n, 256);
  pad_rows_kernel<<<blocks_for(B * d.Lp, 256), 256, 0, s>>>(latents, d.L, p.lat, nullptr, d.Lp, B);
  count_launch();
  if (noise) pad_rows_kernel<<<nblk, 256, 0, s>>>(noise, d.D, p.x, nullptr, d.Dp, B);
  else fill_normal_kernel<<<blocks_for(n / 4, 256), 256, 0, s>>>(p.x, d.Dp, d.D, B, seed);
  count_launch();

  const bool direct_out = d.D == d.Dp;   // no padded columns: the last update writes the caller's buffer, no un-padding pass
  // dst = f(src, [t, h]) with optional classifier-free guidance (sampling.py:62-81)
  auto eval_f = [&](const float* src, float t, float h, float* dst) -> int {
    cond_const_kernel<<<1, 128, 0, s>>>(t, h, p.cond, d);   // one row: (t, h) is the same for every sample
    count_launch();
    MFAC_OK(forward_pass(d, sh, p.cond, p.lat, dst, B, p.fs, s, /*uniform_cond=*/true, nullptr, src));
    if (guidance_scale != 1.0f) {
      MFAC_OK(forward_pass(d, sh, p.cond, nullptr, p.tmp, B, p.fs, s, /*uniform_cond=*/true, nullptr, src));
      axpy2_kernel<<<nblk, 256, 0, s>>>(nullptr, 1.0f, guidance_scale, dst, 1.0f - guidance_scale, p.tmp, dst, n);
      count_launch();
    }
    return MFAC_SUCCESS;
  };

  if (mode == MFAC_SAMPLE_HEUN) {
    const float dt = 1.0f / (float)n_steps;
    for (int i = 0; i < n_steps; ++i) {
      // jnp.linspace(1, 0, n)[i]; n = 1 gives [1.0]
      const float t = n_steps == 1 ? 1.0f : 1.0f - (float)i / (float)(n_steps - 1);
      MFAC_OK(eval_f(p.x, t, 0.f, p.k1));
      axpy2_kernel<<<nblk, 256, 0, s>>>(p.x, -dt, 1.0f, p.k1, 0.f, nullptr, p.x2, n);
      count_launch();
      MFAC_OK(eval_f(p.x2, t - dt, 0.f, p.k2));
      axpy2_kernel<<<nblk, 256, 0, s>>>(p.x, -0.5f * dt, 1.0f, p.k1, 1.0f, p.k2, (direct_out && i == n_steps - 1) ? out : p.x, n);
      count_launch();
    }
  } else {
    for (int i = 0; i < n_steps; ++i) {
      const float t = 1.0f - (float)i / (float)n_steps, r = 1.0f - (float)(i + 1) / (float)n_steps;
      MFAC_OK(eval_f(p.x, t, t - r, p.k1));
      axpy2_kernel<<<nblk, 256, 0, s>>>(p.x, -(t - r), 1.0f, p.k1, 0.f, nullptr, (direct_out && i == n_steps - 1) ? out : p.x, n);
      count_launch();
    }
  }
  if (!direct_out) {
    unpad_rows_kernel<<<blocks_for(B * d.D, 256), 256, 0, s>>>(p.x, d.Dp, out, d.D, B);
    count_launch();
  }
  return launch_status();
}

}  // extern "C"
