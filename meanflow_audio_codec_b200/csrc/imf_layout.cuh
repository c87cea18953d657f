// Parameter / shadow / activation layouts of the MLP velocity network (models/mlp_flow.py).
//
// Flat fp32 parameter order (== jax tree_flatten of the Flax tree, blocks in numeric order):
//   for k in 0..nb-1:  cond1.bias[C] cond1.kernel[C,C] cond2.bias[2I+D] cond2.kernel[C,2I+D]
//                      mlp1.bias[I]  mlp1.kernel[I,I]  mlp2.bias[D]     mlp2.kernel[I,D]
//   encoder:           enc1.bias[He] enc1.kernel[D,He] enc2.bias[L]     enc2.kernel[He,L]
// with I = L + D, He = (D + L) / 2.
//
// Internal (device) feature dimensions are padded to multiples of 64 so every GEMM operand is
// TMA/UMMA friendly for ANY model size (the reference's tests use D=8, L=64, C=32):
//   Dp, Lp, Cp, Hep = round_up(., 64);  Ip = Lp + Dp;  Mp = 2*Ip + Dp.
// The concatenated block input [latents | x] lives as [lat (Lp) | x (Dp)], i.e. real column
// j of the Flax concat maps to cm(j) = j < L ? j : Lp + (j - L).  The modulation vector
// [s1 (I) | shift (I) | s2 (D)] lives as [s1 (Ip) | shift (Ip) | s2 (Dp)] with the same map.
// Padded weight entries are zero and padded activation columns are written as zero.
#pragma once

#include "mfac_common.cuh"

namespace mfac {

struct Dims {
  int D, L, C, nb;
  int I, He;
  int Dp, Lp, Cp, Ip, Hep, Mp;
  int Ca;  // nb * Cp: width of the batched first modulation layer (all blocks' cond1 kernels side by side)
  // flat parameter offsets inside one block, the block stride, and the encoder base
  int64_t o_c1b, o_c1w, o_c2b, o_c2w, o_m1b, o_m1w, o_m2b, o_m2w, blk_stride;
  int64_t o_e1b, o_e1w, o_e2b, o_e2w, total;
  // bf16 shadow offsets (elements) inside one block, block stride, encoder offsets, total
  // The cond1 kernels of ALL blocks form one [Cp, nb*Cp] matrix at s_c1all (block k = columns [k*Cp, (k+1)*Cp)): every
  // block's first modulation layer sees the same input row, so one GEMM evaluates them all.
  int64_t s_c2w, s_m1w, s_m2w, s_blk_stride, s_e1w, s_e2w, s_c1all, s_total;
  // padded fp32 bias offsets (elements) inside the bias section of the shadow
  int64_t b_c2, b_m1, b_m2, b_blk_stride, b_e1, b_e2, b_c1all, b_total;
  int64_t bias_section_bytes_offset;  // byte offset of the fp32 bias section inside the shadow buffer

  __host__ __device__ int cm(int j) const { return j < L ? j : Lp + (j - L); }
  // padded concat column -> real column or -1
  __host__ __device__ int cm_inv(int p) const {
    if (p < L) return p;
    if (p < Lp) return -1;
    const int j = p - Lp;
    return j < D ? L + j : -1;
  }
  // real modulation column -> padded
  __host__ __device__ int mm(int j) const {
    if (j < I) return cm(j);
    if (j < 2 * I) return Ip + cm(j - I);
    return 2 * Ip + (j - 2 * I);
  }
  __host__ __device__ int mm_inv(int p) const {
    if (p < Ip) return cm_inv(p);
    if (p < 2 * Ip) {
      const int r = cm_inv(p - Ip);
      return r < 0 ? -1 : I + r;
    }
    const int j = p - 2 * Ip;
    return j < D ? 2 * I + j : -1;
  }
};

inline int make_dims(const MfacMlpDims* d, Dims* out) {
  if (!d) return MFAC_ERR_NULL;
  // nb == 0 is the encoder on its own (layout = the four encoder leaves): accepted by mfac_mlp_encode, the casts and AdamW;
  // every entry point that evaluates the velocity network asks for nb >= 1 itself.
  if (d->D <= 0 || d->L <= 0 || d->C <= 0 || d->nb < 0 || (d->C & 1)) return MFAC_ERR_BAD_SHAPE;
  Dims x{};
  x.D = d->D; x.L = d->L; x.C = d->C; x.nb = d->nb;
  x.I = x.L + x.D;
  x.He = (x.D + x.L) / 2;
  if (x.He <= 0) return MFAC_ERR_BAD_SHAPE;
  x.Dp = round_up(x.D, 64); x.Lp = round_up(x.L, 64); x.Cp = round_up(x.C, 64); x.Hep = round_up(x.He, 64);
  x.Ip = x.Lp + x.Dp;
  x.Mp = 2 * x.Ip + x.Dp;
  x.Ca = x.nb * x.Cp;
  const int64_t I = x.I, D = x.D, C = x.C, L = x.L, He = x.He;
  int64_t o = 0;
  x.o_c1b = o; o += C;
  x.o_c1w = o; o += C * C;
  x.o_c2b = o; o += 2 * I + D;
  x.o_c2w = o; o += C * (2 * I + D);
  x.o_m1b = o; o += I;
  x.o_m1w = o; o += I * I;
  x.o_m2b = o; o += D;
  x.o_m2w = o; o += I * D;
  x.blk_stride = o;
  o = x.blk_stride * x.nb;
  x.o_e1b = o; o += He;
  x.o_e1w = o; o += D * He;
  x.o_e2b = o; o += L;
  x.o_e2w = o; o += He * L;
  x.total = o;
  int64_t s = 0;
  auto take = [&](int64_t n) { int64_t at = s; s += round_up<int64_t>(n, 128); return at; };
  x.s_c2w = take((int64_t)x.Cp * x.Mp);
  x.s_m1w = take((int64_t)x.Ip * x.Ip);
  x.s_m2w = take((int64_t)x.Ip * x.Dp);
  x.s_blk_stride = s;
  s = x.s_blk_stride * x.nb;
  x.s_e1w = take((int64_t)x.Dp * x.Hep);
  x.s_e2w = take((int64_t)x.Hep * x.Lp);
  x.s_c1all = take((int64_t)x.Cp * x.Ca);
  x.s_total = s;
  int64_t bo = 0;
  auto takeb = [&](int64_t n) { int64_t at = bo; bo += round_up<int64_t>(n, 64); return at; };
  x.b_c2 = takeb(x.Mp);
  x.b_m1 = takeb(x.Ip);
  x.b_m2 = takeb(x.Dp);
  x.b_blk_stride = bo;
  bo = x.b_blk_stride * x.nb;
  x.b_e1 = takeb(x.Hep);
  x.b_e2 = takeb(x.Lp);
  x.b_c1all = takeb(x.Ca);
  x.b_total = bo;
  x.bias_section_bytes_offset = round_up<int64_t>(x.s_total * 2, 256);
  *out = x;
  return MFAC_SUCCESS;
}

// Maps a flat parameter index to its slot in the shadow: kernels -> bf16 element index in the weight
// section, biases -> fp32 element index in the bias section.  Used by the cast and AdamW kernels.
struct ShadowSlot {
  int64_t idx;
  bool is_bias;
};
__host__ __device__ inline ShadowSlot shadow_slot_of(const Dims& d, int64_t i) {
  if (i < d.blk_stride * d.nb) {
    const int64_t k = i / d.blk_stride;
    const int64_t w = i - k * d.blk_stride;
    const int64_t base_s = k * d.s_blk_stride, base_b = k * d.b_blk_stride;
    if (w < d.o_c1w) return {d.b_c1all + k * d.Cp + (w - d.o_c1b), true};
    if (w < d.o_c2b) {
      const int64_t q = w - d.o_c1w;
      const uint32_t qq = (uint32_t)q; const int r = (int)(qq / (uint32_t)d.C), c = (int)(qq % (uint32_t)d.C);
      return {d.s_c1all + (int64_t)r * d.Ca + k * d.Cp + c, false};
    }
    if (w < d.o_c2w) return {base_b + d.b_c2 + d.mm((int)(w - d.o_c2b)), true};
    if (w < d.o_m1b) {
      const int64_t q = w - d.o_c2w;
      const int W = 2 * d.I + d.D;
      const uint32_t qq = (uint32_t)q; const int r = (int)(qq / (uint32_t)W), c = (int)(qq % (uint32_t)W);
      return {base_s + d.s_c2w + (int64_t)r * d.Mp + d.mm(c), false};
    }
    if (w < d.o_m1w) return {base_b + d.b_m1 + d.cm((int)(w - d.o_m1b)), true};
    if (w < d.o_m2b) {
      const int64_t q = w - d.o_m1w;
      const uint32_t qq = (uint32_t)q; const int r = (int)(qq / (uint32_t)d.I), c = (int)(qq % (uint32_t)d.I);
      return {base_s + d.s_m1w + (int64_t)d.cm(r) * d.Ip + d.cm(c), false};
    }
    if (w < d.o_m2w) return {base_b + d.b_m2 + (w - d.o_m2b), true};
    const int64_t q = w - d.o_m2w;
    const uint32_t qq = (uint32_t)q; const int r = (int)(qq / (uint32_t)d.D), c = (int)(qq % (uint32_t)d.D);
    return {base_s + d.s_m2w + (int64_t)d.cm(r) * d.Dp + c, false};
  }
  if (i < d.o_e1w) return {d.b_e1 + (i - d.o_e1b), true};
  if (i < d.o_e2b) {
    const int64_t q = i - d.o_e1w;
    const uint32_t qq = (uint32_t)q; const int r = (int)(qq / (uint32_t)d.He), c = (int)(qq % (uint32_t)d.He);
    return {d.s_e1w + (int64_t)r * d.Hep + c, false};
  }
  if (i < d.o_e2w) return {d.b_e2 + (i - d.o_e2b), true};
  const int64_t q = i - d.o_e2w;
  const uint32_t qq = (uint32_t)q; const int r = (int)(qq / (uint32_t)d.L), c = (int)(qq % (uint32_t)d.L);
  return {d.s_e2w + (int64_t)r * d.Lp + c, false};
}

// Fused AdamW (adamw.cu), launchable over slices of the flat parameter vector
struct AdamWLaunch {
  float* params;
  const float* grads;
  float* mu;
  float* nu;
  void* shadow;              // may be null
  float lr, b1, b2, eps, weight_decay, grad_scale;
  float inv_bc1, inv_bc2;    // filled by adamw_prepare
  const float* bc_dev;       // device copy of the two factors (graph replay) or null
};
// advance = false leaves the device counter untouched (the fused step's prologue still has to read it as the RNG step;
// adamw_advance increments it once the step's launches are enqueued)
int adamw_prepare(AdamWLaunch& h, int64_t count, uint64_t* count_dev, float* scratch_dev, cudaStream_t s, bool advance);
int adamw_advance(uint64_t* count_dev, cudaStream_t s);
int adamw_segments(const Dims& d, const AdamWLaunch& h, int64_t seg_off, int64_t seg_cnt, int64_t seg_stride, int nseg,
                   cudaStream_t s);

// Simple bump allocator over the caller-provided workspace (256-byte aligned slices).
struct Arena {
  uint8_t* base;
  size_t cap;
  size_t off = 0;
  bool overflow = false;
  Arena(void* p, size_t n) : base(reinterpret_cast<uint8_t*>(p)), cap(n) {}
  template <typename T>
  T* take(size_t count) {
    const size_t bytes = round_up<size_t>(count * sizeof(T), 256);
    size_t at = round_up<size_t>(off, 256);
    off = at + bytes;
    if (base == nullptr) return nullptr;  // sizing pass
    if (off > cap) { overflow = true; return reinterpret_cast<T*>(base); }
    return reinterpret_cast<T*>(base + at);
  }
};

}  // namespace mfac
