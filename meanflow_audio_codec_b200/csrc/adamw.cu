// Parameter cast (fp32 Flax layout -> padded bf16 kernels + padded fp32 biases) and fused AdamW.
//
// ref: optax.adamw defaults (b1 0.9, b2 0.999, eps 1e-8, eps_root 0) with decoupled weight decay on
// every leaf, wired at trainers/train.py:236 and applied by TrainState.apply_gradients
// (trainers/training_steps.py:33):
//     m <- b1 m + (1-b1) g          v <- b2 v + (1-b2) g^2
//     p <- p - lr * ( (m / (1-b1^t)) / (sqrt(v / (1-b2^t)) + eps) + wd * p )
// HBM-bound: 16 B read + 12 B write per parameter (+2 B bf16 shadow write).
#include <cmath>

#include "imf_layout.cuh"

namespace mfac {
void count_launch();
void* profile_begin(int family, double work, cudaStream_t s, const char* label = nullptr, int M = 0, int N = 0, int K = 0);
void profile_end(void* token, cudaStream_t s);

namespace {

__global__ void cast_params_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ sw, float* __restrict__ sb,
                                   Dims d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.total) return;
  const ShadowSlot s = shadow_slot_of(d, i);
  const float v = params[i];
  if (s.is_bias) sb[s.idx] = v;
  else sw[s.idx] = __float2bfloat16(v);
}

struct AdamArgs {
  float* p;
  const float* g;
  float* mu;
  float* nu;
  __nv_bfloat16* sw;  // may be null
  float* sb;
  float lr, b1, b2, eps, wd, gscale, inv_bc1, inv_bc2;
  const float* bc_dev;  // optional {inv_bc1, inv_bc2} in device memory (graph replay)
};

// t = *count + 1;  bc = {1 / (1 - b1^t), 1 / (1 - b2^t)};  ++*count
__global__ void adamw_bias_correction_kernel(uint64_t* count, float* bc, float b1, float b2, int advance) {
  const double t = (double)(*count) + 1.0;
  bc[0] = (float)(1.0 / (1.0 - pow((double)b1, t)));
  bc[1] = (float)(1.0 / (1.0 - pow((double)b2, t)));
  if (advance) *count += 1;
}
__global__ void advance_count_kernel(uint64_t* count) { *count += 1; }

// Segment j (= blockIdx.y) covers the flat indices [seg_off + j * seg_stride, + seg_cnt): the whole vector is one segment;
// a training step that applies the update slice by slice (mfac_imf_train_step) passes one block's slice, or the nb short
// first-modulation-layer slices in one launch.  Thread 0 of a segment takes the unaligned head so that every other thread
// works on a 16-byte aligned quad.
__global__ void __launch_bounds__(256) adamw_kernel(AdamArgs a, Dims d, int64_t seg_off, int64_t seg_cnt, int64_t seg_stride) {
  const int64_t begin = seg_off + (int64_t)blockIdx.y * seg_stride, end = begin + seg_cnt;
  const int64_t a4 = (begin + 3) & ~(int64_t)3;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t i0 = tid == 0 ? begin : a4 + (tid - 1) * 4;
  int64_t lim = tid == 0 ? (a4 < end ? a4 : end) : end;
  if (i0 >= lim) return;
  float p[4], g[4], m[4], v[4];
  const float inv_bc1 = a.bc_dev ? a.bc_dev[0] : a.inv_bc1, inv_bc2 = a.bc_dev ? a.bc_dev[1] : a.inv_bc2;
  const bool full = tid != 0 && i0 + 4 <= lim;
  if (full) {
    const float4 p4 = *reinterpret_cast<const float4*>(a.p + i0), g4 = *reinterpret_cast<const float4*>(a.g + i0);
    const float4 m4 = *reinterpret_cast<const float4*>(a.mu + i0), v4 = *reinterpret_cast<const float4*>(a.nu + i0);
    p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w;
    g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
    m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
    v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
  } else {
    for (int q = 0; q < 4; ++q) {
      const bool ok = i0 + q < lim;
      p[q] = ok ? a.p[i0 + q] : 0.f; g[q] = ok ? a.g[i0 + q] : 0.f;
      m[q] = ok ? a.mu[i0 + q] : 0.f; v[q] = ok ? a.nu[i0 + q] : 0.f;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float gq = g[q] * a.gscale;
    m[q] = a.b1 * m[q] + (1.0f - a.b1) * gq;
    v[q] = a.b2 * v[q] + (1.0f - a.b2) * gq * gq;
    const float upd = (m[q] * inv_bc1) / (sqrtf(v[q] * inv_bc2) + a.eps) + a.wd * p[q];
    p[q] = p[q] - a.lr * upd;
  }
  if (full) {
    *reinterpret_cast<float4*>(a.p + i0) = make_float4(p[0], p[1], p[2], p[3]);
    *reinterpret_cast<float4*>(a.mu + i0) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(a.nu + i0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    for (int q = 0; q < 4 && i0 + q < lim; ++q) { a.p[i0 + q] = p[q]; a.mu[i0 + q] = m[q]; a.nu[i0 + q] = v[q]; }
  }
  if (a.sw) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (i0 + q >= lim) break;
      const ShadowSlot s = shadow_slot_of(d, i0 + q);
      if (s.is_bias) a.sb[s.idx] = p[q];
      else a.sw[s.idx] = __float2bfloat16(p[q]);
    }
  }
}

}  // namespace
}  // namespace mfac

extern "C" {

int64_t mfac_mlp_param_count(const MfacMlpDims* dims) {
  mfac::Dims d;
  const int rc = mfac::make_dims(dims, &d);
  return rc == MFAC_SUCCESS ? d.total : rc;
}

int mfac_mlp_param_offset(const MfacMlpDims* dims, int32_t block, int32_t which, int64_t* offset, int64_t* rows,
                          int64_t* cols) {
  mfac::Dims d;
  MFAC_OK(mfac::make_dims(dims, &d));
  if (!offset || !rows || !cols) return MFAC_ERR_NULL;
  const int64_t I = d.I, D = d.D, C = d.C, L = d.L, He = d.He;
  if (block < 0) {
    const int64_t off[4] = {d.o_e1b, d.o_e1w, d.o_e2b, d.o_e2w};
    const int64_t r[4] = {1, D, 1, He}, c[4] = {He, He, L, L};
    if (which < 0 || which > 3) return MFAC_ERR_BAD_SHAPE;
    *offset = off[which]; *rows = r[which]; *cols = c[which];
    return MFAC_SUCCESS;
  }
  if (block >= d.nb || which < 0 || which > 7) return MFAC_ERR_BAD_SHAPE;
  const int64_t off[8] = {d.o_c1b, d.o_c1w, d.o_c2b, d.o_c2w, d.o_m1b, d.o_m1w, d.o_m2b, d.o_m2w};
  const int64_t r[8] = {1, C, 1, C, 1, I, 1, I};
  const int64_t c[8] = {C, C, 2 * I + D, 2 * I + D, I, I, D, D};
  *offset = block * d.blk_stride + off[which]; *rows = r[which]; *cols = c[which];
  return MFAC_SUCCESS;
}

size_t mfac_mlp_shadow_bytes(const MfacMlpDims* dims) {
  mfac::Dims d;
  if (mfac::make_dims(dims, &d) != MFAC_SUCCESS) return 0;
  return (size_t)d.bias_section_bytes_offset + (size_t)d.b_total * 4;
}

int mfac_mlp_cast_params(const MfacMlpDims* dims, const float* params, void* shadow, void* stream) {
  mfac::Dims d;
  MFAC_OK(mfac::make_dims(dims, &d));
  if (!params || !shadow) return MFAC_ERR_NULL;
  cudaStream_t s = (cudaStream_t)stream;
  MFAC_CUDA_OK(cudaMemsetAsync(shadow, 0, mfac_mlp_shadow_bytes(dims), s));
  auto* sw = reinterpret_cast<__nv_bfloat16*>(shadow);
  auto* sb = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(shadow) + d.bias_section_bytes_offset);
  mfac::cast_params_kernel<<<(unsigned)mfac::ceil_div<int64_t>(d.total, 256), 256, 0, s>>>(params, sw, sb, d);
  mfac::count_launch();
  return mfac::launch_status();
}

}  // extern "C"

namespace mfac {
// One launch over nseg segments [seg_off + j * seg_stride, + seg_cnt) of the flat parameter vector (imf.cu: train step).
int adamw_segments(const Dims& d, const AdamWLaunch& h, int64_t seg_off, int64_t seg_cnt, int64_t seg_stride, int nseg,
                   cudaStream_t s) {
  if (seg_cnt <= 0 || nseg <= 0) return MFAC_SUCCESS;
  AdamArgs a;
  a.p = h.params; a.g = h.grads; a.mu = h.mu; a.nu = h.nu;
  a.sw = reinterpret_cast<__nv_bfloat16*>(h.shadow);
  a.sb = h.shadow ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(h.shadow) + d.bias_section_bytes_offset) : nullptr;
  a.lr = h.lr; a.b1 = h.b1; a.b2 = h.b2; a.eps = h.eps; a.wd = h.weight_decay; a.gscale = h.grad_scale;
  a.inv_bc1 = h.inv_bc1; a.inv_bc2 = h.inv_bc2; a.bc_dev = h.bc_dev;
  const int64_t threads = ceil_div<int64_t>(seg_cnt, 4) + 2;
  void* prof = profile_begin(MFAC_PROF_ADAMW, (h.shadow ? 30.0 : 28.0) * (double)seg_cnt * nseg, s);
  adamw_kernel<<<dim3((unsigned)ceil_div<int64_t>(threads, 256), (unsigned)nseg), 256, 0, s>>>(a, d, seg_off, seg_cnt, seg_stride);
  profile_end(prof, s);
  count_launch();
  return launch_status();
}
// host-side bias correction for step count `count` (steps already taken), or the device-side form for graph replay
int adamw_prepare(AdamWLaunch& h, int64_t count, uint64_t* count_dev, float* scratch_dev, cudaStream_t s, bool advance) {
  const double c = (double)count + 1.0;
  h.inv_bc1 = (float)(1.0 / (1.0 - std::pow((double)h.b1, c)));
  h.inv_bc2 = (float)(1.0 / (1.0 - std::pow((double)h.b2, c)));
  h.bc_dev = nullptr;
  if (count_dev) {
    if (!scratch_dev) return MFAC_ERR_NULL;
    adamw_bias_correction_kernel<<<1, 1, 0, s>>>(count_dev, scratch_dev, h.b1, h.b2, advance ? 1 : 0);
    count_launch();
    h.bc_dev = scratch_dev;
  }
  return launch_status();
}
int adamw_advance(uint64_t* count_dev, cudaStream_t s) {
  advance_count_kernel<<<1, 1, 0, s>>>(count_dev);
  count_launch();
  return launch_status();
}
}  // namespace mfac

extern "C" {

namespace {
int adamw_launch(const MfacMlpDims* dims, float* params, const float* grads, float* mu, float* nu, void* shadow, int64_t count,
                 uint64_t* count_dev, float* scratch_dev, float lr, float b1, float b2, float eps, float weight_decay,
                 float grad_scale, void* stream) {
  mfac::Dims d;
  MFAC_OK(mfac::make_dims(dims, &d));
  if (!params || !grads || !mu || !nu) return MFAC_ERR_NULL;
  if (count < 0) return MFAC_ERR_BAD_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  mfac::AdamWLaunch h{params, grads, mu, nu, shadow, lr, b1, b2, eps, weight_decay, grad_scale, 0.f, 0.f, nullptr};
  MFAC_OK(mfac::adamw_prepare(h, count, count_dev, scratch_dev, s, true));
  return mfac::adamw_segments(d, h, 0, d.total, 0, 1, s);
}
}  // namespace

int mfac_adamw_step(const MfacMlpDims* dims, float* params, const float* grads, float* mu, float* nu, void* shadow,
                    int64_t count, float lr, float b1, float b2, float eps, float weight_decay, float grad_scale,
                    void* stream) {
  return adamw_launch(dims, params, grads, mu, nu, shadow, count, nullptr, nullptr, lr, b1, b2, eps, weight_decay, grad_scale,
                      stream);
}

int mfac_adamw_step_dev(const MfacMlpDims* dims, float* params, const float* grads, float* mu, float* nu, void* shadow,
                        uint64_t* count_dev, float* scratch_dev, float lr, float b1, float b2, float eps, float weight_decay,
                        float grad_scale, void* stream) {
  if (!count_dev || !scratch_dev) return MFAC_ERR_NULL;
  return adamw_launch(dims, params, grads, mu, nu, shadow, 0, count_dev, scratch_dev, lr, b1, b2, eps, weight_decay, grad_scale,
                      stream);
}

}  // extern "C"
