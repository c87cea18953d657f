/* libmfac -- C ABI of the B200-native meanflow_audio_codec hot path.
 *
 * Every entry point takes plain device pointers and sizes (no torch / jax types),
 * launches asynchronously on the given CUDA stream (passed as void* so this header
 * needs no CUDA include), does no host synchronisation, owns no caller memory and
 * returns an int status (0 = ok, negative = error; never throws).  Constant tables
 * (window, twiddles, cosine basis) are built in fp64 on the host once per (N, hop)
 * and cached per device inside the library.
 *
 * "ref:" comments name the reference interface each function replaces
 * (paths inside /root/reference/meanflow_audio_codec/).
 */
#ifndef MFAC_H_
#define MFAC_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define MFAC_API __attribute__((visibility("default")))
#else
#define MFAC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define MFAC_SUCCESS 0
#define MFAC_ERR_BAD_SHAPE (-1)    /* non-positive size, inconsistent dims */
#define MFAC_ERR_UNSUPPORTED (-2)  /* configuration the kernels do not cover */
#define MFAC_ERR_WORKSPACE (-3)    /* workspace pointer null or too small */
#define MFAC_ERR_NULL (-4)         /* required pointer is null */
#define MFAC_ERR_DRIVER (-5)       /* CUDA driver entry point unavailable */
#define MFAC_ERR_NCCL (-6)         /* NCCL unavailable or returned an error */
/* CUDA runtime errors are returned as -(1000 + cudaError_t). */

MFAC_API int mfac_version(void);   /* 101: MfacImfConfig gained method, gamma, uniform_time, rows_r_equals_t */
MFAC_API const char* mfac_status_string(int status);

/* ------------------------------------------------------------------ MDCT / IMDCT
 * ref: preprocessing/mdct.py:143-198 (mdct), :201-256 (imdct); direct-cosine branch
 * :317-372 with framing :476-495 and overlap-add :517-540.  N = window_size
 * (coefficients per frame, a frame is 2N samples), hop = hop_size. */

/* ref: mdct.py:491  nf = 1 if T < N else (T - N) / hop + 1 */
MFAC_API int64_t mfac_mdct_num_frames(int64_t T, int32_t N, int32_t hop);
/* ref: mdct.py:513  L = (nf - 1) * hop + 2N */
MFAC_API int64_t mfac_imdct_length(int64_t nf, int32_t N, int32_t hop);

/* x[B, T] (row-major, fp32) -> X[B, nf, N].  Zero right-padding to (nf-1)*hop+2N is implicit. */
MFAC_API int mfac_mdct_f32(const float* x, float* X, int64_t B, int64_t T, int32_t N, int32_t hop, void* stream);
/* X[B, nf, N] -> y[B, (nf-1)*hop + 2N]. */
MFAC_API int mfac_imdct_f32(const float* X, float* y, int64_t B, int64_t nf, int32_t N, int32_t hop, void* stream);

/* Strided forms used for channel-interleaved audio [B, T, C] <-> tokens [B, nf, N*C]
 * (ref: preprocessing/tokenization.py:73-129, mdct.py:602-611,672-693): sample s of clip b is
 * x[b * x_clip_stride + s * x_elem_stride]; coefficient k of frame i is
 * X[b * X_clip_stride + i * X_frame_stride + k].  Strides are in elements. */
MFAC_API int mfac_mdct_strided_f32(const float* x, int64_t x_clip_stride, int64_t x_elem_stride, float* X,
                          int64_t X_clip_stride, int64_t X_frame_stride, int64_t B, int64_t T, int32_t N,
                          int32_t hop, void* stream);
MFAC_API int mfac_imdct_strided_f32(const float* X, int64_t X_clip_stride, int64_t X_frame_stride, float* y,
                           int64_t y_clip_stride, int64_t y_elem_stride, int64_t B, int64_t nf, int32_t N,
                           int32_t hop, void* stream);

/* ------------------------------------------------------------------ MLP velocity network
 * ref: models/mlp_flow.py:125-230 (ConditionalFlow), :63-117 (block), :39-55 (encoder).
 *
 * Parameters are ONE flat fp32 device array in jax tree_flatten order of the Flax tree
 * (blocks_0..blocks_{nb-1} in numeric order, then encoder; inside a block:
 * conditioning_layer/dense1/{bias,kernel}, conditioning_layer/dense2/{bias,kernel},
 * mlp/dense1/{bias,kernel}, mlp/dense2/{bias,kernel}); kernels are [in, out] row-major
 * exactly as flax.linen.Dense stores them.  Gradients and AdamW moments use the same layout.
 * The tensor-core GEMMs read a bf16 "shadow" of the kernels (mfac_mlp_cast_params, refreshed
 * by mfac_adamw_step); biases are read from the fp32 array. */
typedef struct MfacMlpDims {
  int32_t D;  /* noise_dimension (tokens flattened: nf * N) */
  int32_t L;  /* latent_dimension */
  int32_t C;  /* condition_dimension (even) */
  int32_t nb; /* num_blocks; 0 = the encoder on its own (layout = the four encoder leaves): accepted by the layout queries,
               * mfac_mlp_cast_params, mfac_mlp_encode and AdamW, MFAC_ERR_BAD_SHAPE everywhere the velocity network is evaluated */
} MfacMlpDims;

enum MfacParamId {
  MFAC_P_COND1_B = 0, MFAC_P_COND1_W = 1, MFAC_P_COND2_B = 2, MFAC_P_COND2_W = 3,
  MFAC_P_MLP1_B = 4, MFAC_P_MLP1_W = 5, MFAC_P_MLP2_B = 6, MFAC_P_MLP2_W = 7,
  /* encoder (block = -1) */
  MFAC_P_ENC1_B = 0, MFAC_P_ENC1_W = 1, MFAC_P_ENC2_B = 2, MFAC_P_ENC2_W = 3
};

MFAC_API int64_t mfac_mlp_param_count(const MfacMlpDims* dims);
/* Offset (elements) and shape of one leaf; block = -1 selects the encoder. rows = 1 for biases. */
MFAC_API int mfac_mlp_param_offset(const MfacMlpDims* dims, int32_t block, int32_t which, int64_t* offset, int64_t* rows,
                          int64_t* cols);
MFAC_API size_t mfac_mlp_shadow_bytes(const MfacMlpDims* dims);
MFAC_API int mfac_mlp_cast_params(const MfacMlpDims* dims, const float* params, void* shadow, void* stream);

enum MfacWorkspaceKind {
  MFAC_WS_FORWARD = 0, /* mfac_mlp_forward / mfac_mlp_encode */
  MFAC_WS_LOSS_GRAD = 1, /* mfac_imf_loss_grad */
  MFAC_WS_SAMPLE = 2     /* mfac_sample */
};
MFAC_API size_t mfac_workspace_bytes(int32_t kind, const MfacMlpDims* dims, int64_t B);

/* ref: model.apply({"params": p}, x, method="encode")  (mlp_flow.py:153-162).  x[B,D] -> latents[B,L] */
MFAC_API int mfac_mlp_encode(const MfacMlpDims* dims, const float* params, const void* shadow, const float* x,
                    float* latents, int64_t B, void* ws, size_t ws_bytes, void* stream);
/* ref: model.apply({"params": p}, x, time, latents)  (mlp_flow.py:199-230).
 * x[B,D], time[B,2] = (t,h), latents[B,L] or NULL (zeros, mlp_flow.py:223-228) -> out[B,D] */
MFAC_API int mfac_mlp_forward(const MfacMlpDims* dims, const float* params, const void* shadow, const float* x,
                     const float* time, const float* latents, float* out, int64_t B, void* ws, size_t ws_bytes,
                     void* stream);

/* ------------------------------------------------------------------ iMF loss + gradients
 * ref: ImprovedMeanFlowLoss.compute_loss (trainers/loss_strategies.py:227-280) with
 * LinearNoiseSchedule (trainers/noise_schedules.py:69-88), sample_tr (utils.py:36-45) and
 * weighted_l2_loss (utils.py:16-25).
 *
 * e[B,D], t[B], r[B] may be given explicitly (parity mode; the reference re-draws the same
 * values every step) or all NULL, in which case they are drawn on the device
 * from a Philox counter stream keyed by (seed, step, rank offset).  grads is the flat fp32
 * gradient (same layout as params; overwritten).  loss is one fp32.  The optional outputs
 * (any may be NULL) expose intermediates for parity tests. */
typedef struct MfacImfConfig {
  float noise_min, noise_max;          /* 0.001, 0.999 */
  float time_mean, time_std;           /* -0.4, 1.0 */
  float data_proportion;               /* 0.5 : first int(B*p) rows get r = t */
  float loss_c;                        /* 1e-3 (p = 1) */
  int32_t use_weighted_loss;           /* 1 */
  uint64_t seed;
  uint64_t step;
  uint64_t row_offset;                 /* global index of local row 0 (rank * B) for the RNG */
  const uint64_t* step_dev;            /* optional DEVICE counter read instead of `step` (CUDA-graph replay: the
                                          captured launch must see a step that advances); NULL = use `step` */
  int32_t method;                      /* MfacLossMethod: which loss strategy (trainers/loss_strategies.py) */
  float gamma;                         /* mean flow: adaptive weight 1 / (mean_D delta^2 + loss_c)^(1 - gamma) */
  int32_t uniform_time;                /* 1: t ~ U(0,1) (UniformTimeSampling) instead of logit-normal, internal RNG only */
  int64_t rows_r_equals_t;             /* improved mean flow with explicit (t, r): the caller guarantees r[b] == t[b] for
                                          b < this many leading rows (sample_tr's rule, utils.py:41-44).  On those rows the
                                          u evaluation f(z, [t, t-r]) IS the v evaluation f(z, [t, 0]) and du/dt only meets
                                          the factor (t - r) == 0, so one saved primal pass serves both and no tangent is
                                          computed for them (bit-identical loss and gradients; du/dt of those rows is
                                          reported as 0).  0 = no promise.  With the internal RNG the library applies
                                          the rule itself (int(B * data_proportion) rows); -1 switches the sharing off. */
} MfacImfConfig;

/* ref: trainers/loss_strategies.py -- ImprovedMeanFlowLoss :204-280 (tangent seed = the network's own v),
 * MeanFlowLoss :115-201 (z = (1-t)x + t e, tangent seed = e - x, adaptive weight; pass noise_min = 0, noise_max = 1),
 * FlowMatchingLoss :50-112 (single time, h = 0, no JVP; r is ignored). */
enum MfacLossMethod { MFAC_LOSS_IMPROVED_MEAN_FLOW = 0, MFAC_LOSS_MEAN_FLOW = 1, MFAC_LOSS_FLOW_MATCHING = 2 };

typedef struct MfacImfAux {
  float* v;           /* [B,D] */
  float* u;           /* [B,D] */
  float* dudt;        /* [B,D] */
  float* per_example; /* [B]  sum_d delta^2 */
  float* e;           /* [B,D] noise actually used */
  float* t;           /* [B] */
  float* r;           /* [B] */
  /* Optional host callback, invoked on the calling thread right after the launches that finalise the gradient
   * slice grads[offset, offset+count) have been enqueued on `stream` (blocks nb-1 .. 0, then the encoder).  A
   * data-parallel caller starts that bucket's all-reduce from it, so the exchange overlaps the rest of the backward
   * (the reference has no distributed code). */
  void (*grad_ready)(void* user, int64_t offset, int64_t count);
  void* grad_ready_user;
} MfacImfAux;

MFAC_API int mfac_imf_loss_grad(const MfacMlpDims* dims, const MfacImfConfig* cfg, const float* params, const void* shadow,
                       const float* x, const float* e, const float* t, const float* r, float* loss, float* grads,
                       const MfacImfAux* aux, int64_t B, void* ws, size_t ws_bytes, void* stream);

/* ref: optax.adamw(lr, weight_decay) + TrainState.apply_gradients
 * (trainers/train.py:236, trainers/training_steps.py:33).  Decoupled decay on ALL leaves.
 * count = steps already taken.  grad_scale multiplies the gradient first (1/world after a
 * sum all-reduce).  If shadow != NULL the bf16 kernels are refreshed in the same pass. */
MFAC_API int mfac_adamw_step(const MfacMlpDims* dims, float* params, const float* grads, float* mu, float* nu, void* shadow,
                    int64_t count, float lr, float b1, float b2, float eps, float weight_decay, float grad_scale,
                    void* stream);

/* Graph-capturable form: the step count lives in device memory.  *count_dev is read for the bias correction and
 * then incremented on the device, so one captured launch sequence can be replayed step after step.
 * scratch_dev: 2 floats of device scratch (bias-correction factors). */
MFAC_API int mfac_adamw_step_dev(const MfacMlpDims* dims, float* params, const float* grads, float* mu, float* nu, void* shadow,
                        uint64_t* count_dev, float* scratch_dev, float lr, float b1, float b2, float eps, float weight_decay,
                        float grad_scale, void* stream);

/* ------------------------------------------------------------------ fused training step
 * ref: train_step (trainers/training_steps.py:15-61) = compute_loss + TrainState.apply_gradients with optax.adamw
 * (trainers/train.py:236).  One call: mfac_imf_loss_grad's schedule, and the AdamW update (with the bf16 shadow refresh)
 * applied slice by slice as the backward finalises each block's gradient -- for small batches on the library's side stream,
 * under the rest of the backward -- instead of one pass over all parameters afterwards.  world > 1: every slice is first
 * sum-all-reduced over the communicator of mfac_comm_init (NCCL on the stream the slice became final on, so the exchange
 * overlaps the remaining backward) and scaled by 1 / world.  `count` = steps already taken; count_dev / scratch_dev select the
 * graph-capturable form of mfac_adamw_step_dev (count is then ignored).  grads is caller-provided scratch of the parameter
 * vector's size and holds the (summed) gradient afterwards. */
typedef struct MfacAdamWConfig {
  float lr, b1, b2, eps, weight_decay;   /* optax.adamw: 1e-4 (config base_lr), 0.9, 0.999, 1e-8, 1e-4 */
} MfacAdamWConfig;
MFAC_API int mfac_imf_train_step(const MfacMlpDims* dims, const MfacImfConfig* cfg, const MfacAdamWConfig* opt, float* params,
                        void* shadow, float* mu, float* nu, int64_t count, uint64_t* count_dev, float* scratch_dev,
                        const float* x, const float* e, const float* t, const float* r, float* loss, float* grads,
                        const MfacImfAux* aux, int64_t B, int32_t world, void* ws, size_t ws_bytes, void* stream);

/* Raw-audio forms (the hot loop tokenises and steps back to back, trainers/train.py:339-345):
 * audio[B, T] fp32 takes the place of x[B, D]; each clip's MDCT tokens (window_size, hop_size; direct cosine branch) are one
 * model row, nf * window_size == D.  For window 512 / hop 256 and up to 16 frames per row the tokeniser's store stage IS the
 * step's prologue -- no token tensor is written to HBM; other geometries tokenise into scratch first (same results). */
MFAC_API int mfac_imf_loss_grad_audio(const MfacMlpDims* dims, const MfacImfConfig* cfg, const float* params, const void* shadow,
                             const float* audio, int64_t T, int32_t window_size, int32_t hop_size, const float* e, const float* t,
                             const float* r, float* loss, float* grads, const MfacImfAux* aux, int64_t B, void* ws,
                             size_t ws_bytes, void* stream);
MFAC_API int mfac_imf_train_step_audio(const MfacMlpDims* dims, const MfacImfConfig* cfg, const MfacAdamWConfig* opt, float* params,
                              void* shadow, float* mu, float* nu, int64_t count, uint64_t* count_dev, float* scratch_dev,
                              const float* audio, int64_t T, int32_t window_size, int32_t hop_size, const float* e, const float* t,
                              const float* r, float* loss, float* grads, const MfacImfAux* aux, int64_t B, int32_t world, void* ws,
                              size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ samplers
 * MFAC_SAMPLE_HEUN  ref: evaluators/sampling.py:5-95 (h = 0, grid linspace(1,0,n), dt = 1/n).
 * MFAC_SAMPLE_MF    mean-flow few-step rule x_r = x_t - (t-r) u(x_t,[t,t-r]) on a uniform grid
 *                   (documentation/research/improved_meanflow/improved_meanflow_key_eqn.md:311-318);
 *                   n_steps = 1 or 2 are the 1-/2-NFE samplers.
 * noise[B,D] may be NULL (Philox, keyed by seed).  latents[B,L] required (sampling.py:42-45). */
enum MfacSampleMode { MFAC_SAMPLE_HEUN = 0, MFAC_SAMPLE_MF = 1 };
MFAC_API int mfac_sample(const MfacMlpDims* dims, const float* params, const void* shadow, const float* latents,
                const float* noise, int32_t mode, int32_t n_steps, float guidance_scale, uint64_t seed, float* out,
                int64_t B, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ MLP-Mixer / ConvNeXt velocity networks (forward)
 * ref: ConditionalMLPMixerFlow models/mlp_mixer.py:171-235; ConditionalConvFlow models/conv_flow.py:213-271.
 * Same call signature as the MLP flow: x[B,D], time[B,2], latents[B, latent_flat] (the reference's
 * [B, num_latent_tokens, latent_dim] flattened, mlp_mixer.py:224-229) or NULL -> out[B,D].
 * Weights are passed as device pointers: kernels bf16 [in,out] row-major (Flax layout), biases fp32.
 * The block arrays are HOST arrays of structs holding device pointers.  Forward only (the reference never
 * trains these models). */
typedef struct MfacDense {
  const void* w;   /* bf16 [in, out] */
  const float* b;  /* fp32 [out] */
} MfacDense;

typedef struct MfacMixerDims {
  int32_t D, C, nb;
  int32_t tokens;      /* int(sqrt(D))^2 */
  int32_t channels;    /* num_channels (8, 16 or 32) */
  int32_t token_mix, channel_mix;
  int32_t latent_flat; /* num_latent_tokens * latent_dim, 0 if latents are never passed */
} MfacMixerDims;
typedef struct MfacMixerBlockW {
  MfacDense input_proj, adaln1, tok1, tok2, adaln2, ch1, ch2, output_proj;
} MfacMixerBlockW;
typedef struct MfacMixerWeights {
  const MfacMixerBlockW* blocks; /* host array [nb] */
  MfacDense latent_proj;
} MfacMixerWeights;
MFAC_API size_t mfac_mixer_workspace_bytes(const MfacMixerDims* dims, int64_t B);
MFAC_API int mfac_mixer_forward(const MfacMixerDims* dims, const MfacMixerWeights* w, const float* x, const float* time,
                       const float* latents, float* out, int64_t B, void* ws, size_t ws_bytes, void* stream);

typedef struct MfacConvDims {
  int32_t D, C, nb;
  int32_t S;           /* int(sqrt(D)) */
  int32_t channels;    /* min(16, C / 4): 4, 8 or 16 */
  int32_t bottleneck;  /* 128 */
  int32_t latent_flat;
} MfacConvDims;
typedef struct MfacConvBlockW {
  MfacDense input_proj1, input_proj2, conditioning, output_proj1, output_proj2;
  /* ConvNeXt block, all fp32 (the image stays in shared memory; SIMT fp32 arithmetic) */
  const float* conv3_w;     /* [3, 3, ch, ch]  (kh, kw, in, out) */
  const float* conv3_b;     /* [ch] */
  const float* pw1_w;       /* [ch, 2ch] */
  const float* pw1_b;       /* [2ch] */
  const float* grn_gamma;   /* [2ch] or NULL (zeros) */
  const float* grn_beta;    /* [2ch] or NULL */
  const float* pw2_w;       /* [2ch, ch] */
  const float* pw2_b;       /* [ch] */
  const float* layer_scale; /* [ch] or NULL (ones) */
} MfacConvBlockW;
typedef struct MfacConvWeights {
  const MfacConvBlockW* blocks; /* host array [nb] */
  MfacDense latent_proj;
} MfacConvWeights;
MFAC_API size_t mfac_conv_workspace_bytes(const MfacConvDims* dims, int64_t B);
MFAC_API int mfac_conv_forward(const MfacConvDims* dims, const MfacConvWeights* w, const float* x, const float* time,
                      const float* latents, float* out, int64_t B, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ data-parallel gradient all-reduce
 * (absent in the reference).  id_bytes = the 128-byte ncclUniqueId. */
MFAC_API int mfac_comm_unique_id(void* id_bytes_out);
MFAC_API int mfac_comm_init(const void* id_bytes, int32_t rank, int32_t world);
MFAC_API int mfac_comm_allreduce_sum_f32(float* buf, int64_t count, void* stream);
MFAC_API int mfac_comm_destroy(void);

/* ------------------------------------------------------------------ scheduling knob
 * Batches of up to `rows` rows run the concurrent schedules of mfac_imf_loss_grad: the three forward evaluations of the iMF
 * loss (trainers/loss_strategies.py:250-267) as independent kernel chains on the library's own side streams, and the
 * weight-gradient GEMMs of the backward beside the critical dX chain; everything is joined back into the caller's stream
 * before the call returns.  Default 4096 (env MFAC_CONC_MAX_ROWS); 0 = always the single-stream schedule.  The value also
 * decides the layout mfac_workspace_bytes(MFAC_WS_LOSS_GRAD) sizes: change it before sizing the workspace, not between. */
MFAC_API int mfac_set_concurrency_max_rows(int32_t rows);
/* 1 when a batch of B rows of this geometry runs the concurrent schedules (B <= rows and B * D <= rows * 1024), else 0. */
MFAC_API int mfac_uses_concurrent_schedule(const MfacMlpDims* dims, int64_t B);

/* ------------------------------------------------------------------ test hooks
 * C[M,N] fp32 = A * B with bf16 operands through the production tcgen05 GEMM.
 * a_mn_major = 0: A stored [M,K] (K contiguous); 1: stored [K,M].
 * b_mn_major = 0: B stored [N,K] (K contiguous); 1: stored [K,N]. */
MFAC_API int mfac_debug_gemm_bf16(const void* A, const void* B, float* Cout, int64_t M, int64_t N, int64_t K,
                         int32_t a_mn_major, int32_t b_mn_major, int32_t block_n, void* stream);
/* 1 = route every GEMM through a plain SIMT kernel (debug triage only; never set in product use). */
MFAC_API int mfac_debug_set_simt_gemm(int32_t on);
/* 0 = slice K uniformly for the weight-gradient GEMMs; 1 (default) = stream-K (equal contiguous k-block ranges per CTA pair). */
MFAC_API int mfac_debug_set_stream_k(int32_t on);
/* 0 = never use the CTA-pair (cta_group::2) GEMM kernel; 1 (default) = use it where it pays. */
MFAC_API int mfac_debug_set_pair_gemm(int32_t on);
MFAC_API int mfac_debug_counters(int64_t* kernel_launches);
/* Phase timeline of the training step (debug): while on, mfac_imf_loss_grad / mfac_imf_train_step record an event at each
 * phase boundary (1 prologue+encoder done, 2 forward chains joined, 3 tangent done, 4 loss done, 5..: after each backward
 * block, 90 backward done, 99 end); collect returns ids and milliseconds since the first mark and clears the list. */
MFAC_API int mfac_debug_phase_marks(int32_t on);
MFAC_API int mfac_debug_phase_collect(int32_t* ids, float* ms_since_first, int32_t cap);

/* ------------------------------------------------------------------ per-kernel timing (bench.py roofline)
 * While enabled, every launch of a profiled kernel family is bracketed by CUDA events on its own
 * stream (no synchronisation, no serialisation).  mfac_profile_collect synchronises the device and
 * returns, per family, the number of launches, the summed event time (ms) and the summed algorithmic
 * work (FLOPs for MFAC_PROF_GEMM, bytes for the HBM-bound families). */
enum MfacProfFamily { MFAC_PROF_GEMM = 0, MFAC_PROF_MDCT = 1, MFAC_PROF_IMDCT = 2, MFAC_PROF_ADAMW = 3, MFAC_PROF_FAMILIES = 4 };
MFAC_API int mfac_profile_enable(int32_t on);
MFAC_API int mfac_profile_collect(int64_t* launches, double* ms, double* work);

#ifdef __cplusplus
}
#endif
#endif /* MFAC_H_ */
