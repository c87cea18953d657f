// XLA FFI (jax.ffi) handlers over the libmfac C ABI -- the "thin C-ABI layer (jax.ffi custom calls)" of the north star.
//
// Not BUILT in the build image (no jaxlib, hence no xla/ffi/api/ffi.h); there the translation unit is syntax- and type-checked
// against a minimal stand-in for that header (tests/stubs/xla/ffi/api/ffi.h, tests/test_abi_cpu.py) so that every call into
// the C ABI stays in step with include/mfac.h.  Built by meanflow_audio_codec_b200/jax_ffi/build.py on a box where
// `python -c "import jax.ffi; print(jax.ffi.include_dir())"` works:
//   g++ -O2 -std=c++17 -shared -fPIC -I$(jax include dir) -I<repo>/include -I/usr/local/cuda/include \
//       mfac_jax_ffi.cc -L<repo>/meanflow_audio_codec_b200 -lmfac -Wl,-rpath,'$ORIGIN/..' -o libmfac_jax_ffi.so
// Every handler only unpacks buffers and forwards to the C ABI on XLA's stream; shapes are validated on the Python side
// (jax_ffi/__init__.py) exactly as the reference validates them (preprocessing/mdct.py:189-194,247-250).
#include <cuda_runtime_api.h>

#include <cstdint>

#include "mfac.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error status_of(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, mfac_status_string(rc));
}

// x[B, T] -> X[B, nf, N]            (ref: preprocessing/mdct.py:143-198, direct branch :317-327)
static ffi::Error MdctImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> x, ffi::ResultBuffer<ffi::F32> X, int32_t window_size,
                           int32_t hop_size) {
  const auto d = x.dimensions();
  if (d.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "mfac_mdct expects x[B, T]");
  return status_of(mfac_mdct_f32(x.typed_data(), X->typed_data(), d[0], d[1], window_size, hop_size, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfacMdct, MdctImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("window_size")
                                  .Attr<int32_t>("hop_size"));

// X[B, nf, N] -> y[B, (nf-1) hop + 2N]   (ref: preprocessing/mdct.py:201-256, :330-340, :517-540)
static ffi::Error ImdctImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> X, ffi::ResultBuffer<ffi::F32> y, int32_t window_size,
                            int32_t hop_size) {
  const auto d = X.dimensions();
  if (d.size() != 3 || d[2] != window_size) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "mfac_imdct expects X[B, nf, N]");
  return status_of(mfac_imdct_f32(X.typed_data(), y->typed_data(), d[0], d[1], window_size, hop_size, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfacImdct, ImdctImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("window_size")
                                  .Attr<int32_t>("hop_size"));

// model.apply({"params": p}, x, time, latents)   (ref: models/mlp_flow.py:199-230)
// operands: flat params (jax.flatten_util.ravel_pytree order == the layout of include/mfac.h), bf16 shadow (bytes),
// x[B, D], time[B, 2], latents[B, L];  results: out[B, D], workspace scratch (bytes, sized by mfac_workspace_bytes)
static ffi::Error MlpForwardImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::U8> shadow,
                                 ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> time, ffi::Buffer<ffi::F32> latents,
                                 ffi::ResultBuffer<ffi::F32> out, ffi::ResultBuffer<ffi::U8> ws, int32_t D, int32_t L, int32_t C,
                                 int32_t nb) {
  MfacMlpDims dims{D, L, C, nb};
  const int64_t B = x.dimensions()[0];
  return status_of(mfac_mlp_forward(&dims, params.typed_data(), shadow.typed_data(), x.typed_data(), time.typed_data(),
                                    latents.typed_data(), out->typed_data(), B, ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfacMlpForward, MlpForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("D")
                                  .Attr<int32_t>("L")
                                  .Attr<int32_t>("C")
                                  .Attr<int32_t>("nb"));

// ImprovedMeanFlowLoss.compute_loss -> (loss, grads)   (ref: trainers/loss_strategies.py:227-280)
// operands: flat params, shadow, x[B, D], e[B, D], t[B], r[B];  results: loss[], grads[P], workspace scratch
static ffi::Error ImfLossGradImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::U8> shadow,
                                  ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> e, ffi::Buffer<ffi::F32> t,
                                  ffi::Buffer<ffi::F32> r, ffi::ResultBuffer<ffi::F32> loss, ffi::ResultBuffer<ffi::F32> grads,
                                  ffi::ResultBuffer<ffi::U8> ws, int32_t D, int32_t L, int32_t C, int32_t nb) {
  MfacMlpDims dims{D, L, C, nb};
  MfacImfConfig cfg{0.001f, 0.999f, -0.4f, 1.0f, 0.5f, 1e-3f, 1, 0, 0, 0, nullptr, MFAC_LOSS_IMPROVED_MEAN_FLOW, 0.5f, 0, 0};
  const int64_t B = x.dimensions()[0];
  return status_of(mfac_imf_loss_grad(&dims, &cfg, params.typed_data(), shadow.typed_data(), x.typed_data(), e.typed_data(),
                                      t.typed_data(), r.typed_data(), loss->typed_data(), grads->typed_data(), nullptr, B,
                                      ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfacImfLossGrad, ImfLossGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("D")
                                  .Attr<int32_t>("L")
                                  .Attr<int32_t>("C")
                                  .Attr<int32_t>("nb"));

// model.apply({"params": p}, x, method="encode")   (ref: models/mlp_flow.py:153-162)
// operands: flat params, shadow, x[B, D];  results: latents[B, L], workspace scratch
static ffi::Error MlpEncodeImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::U8> shadow,
                                ffi::Buffer<ffi::F32> x, ffi::ResultBuffer<ffi::F32> latents, ffi::ResultBuffer<ffi::U8> ws,
                                int32_t D, int32_t L, int32_t C, int32_t nb) {
  MfacMlpDims dims{D, L, C, nb};
  const int64_t B = x.dimensions()[0];
  return status_of(mfac_mlp_encode(&dims, params.typed_data(), shadow.typed_data(), x.typed_data(), latents->typed_data(), B,
                                   ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfacMlpEncode, MlpEncodeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("D")
                                  .Attr<int32_t>("L")
                                  .Attr<int32_t>("C")
                                  .Attr<int32_t>("nb"));

// sample(apply_fn, D, params, key, latents, n_steps, guidance_scale)   (ref: evaluators/sampling.py:5-95) and the mean-flow
// 1-/2-NFE rule (mode 1).  operands: flat params, shadow, latents[B, L], noise[B, D] (the JAX side draws it with its own key,
// sampling.py:47-48);  results: out[B, D], workspace scratch
static ffi::Error SampleImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::U8> shadow,
                             ffi::Buffer<ffi::F32> latents, ffi::Buffer<ffi::F32> noise, ffi::ResultBuffer<ffi::F32> out,
                             ffi::ResultBuffer<ffi::U8> ws, int32_t D, int32_t L, int32_t C, int32_t nb, int32_t mode,
                             int32_t n_steps, float guidance_scale) {
  MfacMlpDims dims{D, L, C, nb};
  const int64_t B = latents.dimensions()[0];
  return status_of(mfac_sample(&dims, params.typed_data(), shadow.typed_data(), latents.typed_data(), noise.typed_data(), mode,
                               n_steps, guidance_scale, 0, out->typed_data(), B, ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfacSample, SampleImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("D")
                                  .Attr<int32_t>("L")
                                  .Attr<int32_t>("C")
                                  .Attr<int32_t>("nb")
                                  .Attr<int32_t>("mode")
                                  .Attr<int32_t>("n_steps")
                                  .Attr<float>("guidance_scale"));

// mfac_mlp_cast_params: flat fp32 params -> the bf16 shadow the GEMMs read.  operand: params;  result: shadow (bytes)
static ffi::Error CastParamsImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::ResultBuffer<ffi::U8> shadow, int32_t D,
                                 int32_t L, int32_t C, int32_t nb) {
  MfacMlpDims dims{D, L, C, nb};
  return status_of(mfac_mlp_cast_params(&dims, params.typed_data(), shadow->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfacCastParams, CastParamsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("D")
                                  .Attr<int32_t>("L")
                                  .Attr<int32_t>("C")
                                  .Attr<int32_t>("nb"));

// train_step = compute_loss + TrainState.apply_gradients (optax.adamw)   (ref: trainers/training_steps.py:15-61,
// trainers/train.py:236).  XLA buffers are immutable, so the updated state comes back as results that ALIAS the operands
// (input_output_aliases on the Python side): operands params, shadow, mu, nu, x[B, D], e[B, D], t[B], r[B];
// results: params', shadow', mu', nu' (aliased in place), loss[], grads[P] (scratch / summed gradient), workspace scratch.
static ffi::Error ImfTrainStepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::U8> shadow,
                                   ffi::Buffer<ffi::F32> mu, ffi::Buffer<ffi::F32> nu, ffi::Buffer<ffi::F32> x,
                                   ffi::Buffer<ffi::F32> e, ffi::Buffer<ffi::F32> t, ffi::Buffer<ffi::F32> r,
                                   ffi::ResultBuffer<ffi::F32> params_out, ffi::ResultBuffer<ffi::U8> shadow_out,
                                   ffi::ResultBuffer<ffi::F32> mu_out, ffi::ResultBuffer<ffi::F32> nu_out,
                                   ffi::ResultBuffer<ffi::F32> loss, ffi::ResultBuffer<ffi::F32> grads,
                                   ffi::ResultBuffer<ffi::U8> ws, int32_t D, int32_t L, int32_t C, int32_t nb, int64_t count,
                                   float lr, float weight_decay, int32_t world) {
  // the aliased results ARE the operands' memory; refuse a call that was made without the aliases
  if (params_out->typed_data() != params.typed_data() || shadow_out->typed_data() != shadow.typed_data() ||
      mu_out->typed_data() != mu.typed_data() || nu_out->typed_data() != nu.typed_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "mfac_imf_train_step needs input_output_aliases for params/shadow/mu/nu");
  MfacMlpDims dims{D, L, C, nb};
  MfacImfConfig cfg{0.001f, 0.999f, -0.4f, 1.0f, 0.5f, 1e-3f, 1, 0, 0, 0, nullptr, MFAC_LOSS_IMPROVED_MEAN_FLOW, 0.5f, 0, 0};
  MfacAdamWConfig opt{lr, 0.9f, 0.999f, 1e-8f, weight_decay};
  const int64_t B = x.dimensions()[0];
  return status_of(mfac_imf_train_step(&dims, &cfg, &opt, params_out->typed_data(), shadow_out->typed_data(), mu_out->typed_data(),
                                       nu_out->typed_data(), count, nullptr, nullptr, x.typed_data(), e.typed_data(),
                                       t.typed_data(), r.typed_data(), loss->typed_data(), grads->typed_data(), nullptr, B, world,
                                       ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfacImfTrainStep, ImfTrainStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("D")
                                  .Attr<int32_t>("L")
                                  .Attr<int32_t>("C")
                                  .Attr<int32_t>("nb")
                                  .Attr<int64_t>("count")
                                  .Attr<float>("lr")
                                  .Attr<float>("weight_decay")
                                  .Attr<int32_t>("world"));
