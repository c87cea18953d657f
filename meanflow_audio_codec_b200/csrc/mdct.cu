// MDCT analysis / IMDCT overlap-add synthesis for sm_100a.
//
// Reference semantics: the direct-cosine branch of preprocessing/mdct.py
//   X[b,i,k] = sum_{n<2N} x[b, i*hop+n] w[n] cos(pi/N (n + N/2 + 1/2)(k + 1/2))        (mdct.py:317-327,347-358)
//   y[b,s]   = sum_i (2/N) w[s-i*hop] sum_k X[b,i,k] cos(pi/N (s-i*hop + N/2 + 1/2)(k + 1/2))   (:330-340,361-372,517-540)
// with w[n] = sin(pi (n+1/2) / 2N)  (:126-136), nf = 1 if T<N else (T-N)/hop+1, zero right padding (:487-494).
//
// Two code paths, both fp32 arithmetic with tables generated in fp64 on the host:
//  * N == 512 (every shipped config): the 2N->N TDAC fold is fused with the window, the N-point
//    DCT-IV is computed through one 256-point complex FFT per frame (pre/post twiddle), one warp per
//    frame, 8-8-4 radix passes in registers with two padded shared-memory exchanges.  Frames of a
//    clip share one staged input segment (framing costs no extra HBM traffic).  The inverse runs the
//    same DCT-IV, keeps the N-sample core of each frame in shared memory and every output sample
//    gathers its <= ceil(2N/hop) frames (unfold + window + 2/N) -- no scan, no atomics.
//  * any other N: dense contraction against the cached windowed cosine basis (exactly the
//    reference's einsum), tiled through shared memory.
// HBM-bound: algorithmic bytes are 4*(T + nf*N) per clip forward, 4*(nf*N + L) inverse.
#include <map>
#include <mutex>
#include <vector>
#include <cmath>
#include <cstdlib>

#include "mfac_common.cuh"
#include "imf_prep.cuh"   // prologue of the iMF step (PrepArgs, draw_tr, Philox, cond rows) for the fused tokenise + prologue kernel

namespace mfac {

void count_launch();
void* profile_begin(int family, double work, cudaStream_t s, const char* label = nullptr, int M = 0, int N = 0, int K = 0);
void profile_end(void* token, cudaStream_t s);

struct StridedIO {
  int64_t in_clip_stride, in_elem_stride;    // MDCT: x ; IMDCT: X (elem stride = frame stride)
  int64_t out_clip_stride, out_elem_stride;  // MDCT: X (elem stride = frame stride); IMDCT: y
};

namespace {

constexpr int FFT_N = 512;       // window_size handled by the FFT path
constexpr int FFT_H = 256;       // complex FFT length
constexpr int SCR_STRIDE = 36;   // padded row stride (float2) of the per-warp exchange buffer
constexpr int SCR_F2 = 8 * SCR_STRIDE;  // 288 float2 = 576 floats >= 512 staging floats
constexpr int MDCT_WARPS = 8;

struct FftTables {      // device pointers
  const float* window;  // [2N]
  const float2* pre;    // [H]   exp(-i pi (4m+1) / 4N)
  const float2* post;   // [H]   exp(-i pi k / N)
  const float2* w64;    // [64]  exp(-2 pi i q / 64)
  const float2* w256;   // [256] exp(-2 pi i q / 256)
};

struct DenseTables {
  const float* wc;   // [2N, N]  w[n] * C[n,k]
  const float* wct;  // [N, 2N]  (2/N) * w[n] * C[n,k], transposed
};

struct TableSet {
  FftTables fft{};
  DenseTables dense{};
  bool has_fft = false, has_dense = false;
};

std::mutex g_mu;
std::map<std::pair<int, int>, TableSet> g_tables;  // (device, N)
bool g_imdct_gather = getenv("MFAC_IMDCT_GATHER") != nullptr;  // debug: force the shared-memory gather kernel

template <typename T>
int upload(const std::vector<T>& h, const T** out) {
  T* d = nullptr;
  MFAC_CUDA_OK(cudaMalloc(&d, h.size() * sizeof(T)));
  MFAC_CUDA_OK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = d;
  return MFAC_SUCCESS;
}

int get_tables(int N, bool want_fft, TableSet* out) {
  int dev = 0;
  MFAC_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_mu);
  TableSet& ts = g_tables[{dev, N}];
  const double pi = 3.14159265358979323846;
  if (want_fft && !ts.has_fft) {
    std::vector<float> w(2 * N);
    for (int n = 0; n < 2 * N; ++n) w[n] = (float)std::sin(pi * (n + 0.5) / (2.0 * N));
    std::vector<float2> pre(N / 2), post(N / 2), w64(64), w256(256);
    for (int m = 0; m < N / 2; ++m) {
      pre[m] = make_float2((float)std::cos(pi * (4 * m + 1) / (4.0 * N)), (float)-std::sin(pi * (4 * m + 1) / (4.0 * N)));
      post[m] = make_float2((float)std::cos(pi * m / (double)N), (float)-std::sin(pi * m / (double)N));
    }
    for (int q = 0; q < 64; ++q) w64[q] = make_float2((float)std::cos(2 * pi * q / 64.0), (float)-std::sin(2 * pi * q / 64.0));
    for (int q = 0; q < 256; ++q) w256[q] = make_float2((float)std::cos(2 * pi * q / 256.0), (float)-std::sin(2 * pi * q / 256.0));
    MFAC_OK(upload(w, &ts.fft.window));
    MFAC_OK(upload(pre, &ts.fft.pre));
    MFAC_OK(upload(post, &ts.fft.post));
    MFAC_OK(upload(w64, &ts.fft.w64));
    MFAC_OK(upload(w256, &ts.fft.w256));
    ts.has_fft = true;
  }
  if (!want_fft && !ts.has_dense) {
    std::vector<float> wc((size_t)2 * N * N), wct((size_t)2 * N * N);
    for (int n = 0; n < 2 * N; ++n) {
      const double wn = std::sin(pi * (n + 0.5) / (2.0 * N));
      for (int k = 0; k < N; ++k) {
        // exact argument reduction: (2n + N + 1)(2k + 1) mod 8N in integers, then one fp64 cos
        const long long q = ((long long)(2 * n + N + 1) * (2 * k + 1)) % (8LL * N);
        const double c = std::cos(pi * (double)q / (4.0 * N));
        wc[(size_t)n * N + k] = (float)(wn * c);
        wct[(size_t)k * 2 * N + n] = (float)(2.0 / N * wn * c);
      }
    }
    MFAC_OK(upload(wc, &ts.dense.wc));
    MFAC_OK(upload(wct, &ts.dense.wct));
    ts.has_dense = true;
  }
  *out = ts;
  return MFAC_SUCCESS;
}

// ------------------------------------------------------------------ complex helpers
// complex add / subtract as ONE packed fp32x2 instruction (sm_100 FADD2): the butterflies are 40 % of the FFT's FP instructions
// (measured: -9 % / -11 % SASS in the forward / inverse kernels, +1.5 % throughput -- the kernels are bound by the shared-memory
// pipe and by latency at 12 warps per SM, not by FP issue)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return f2add(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long ua = *reinterpret_cast<unsigned long long*>(&a), ub = *reinterpret_cast<unsigned long long*>(&b), r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(ua), "l"(ub));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// forward 4-point DFT, natural order in and out
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = mul_mi(csub(x1, x3));
  x0 = cadd(t0, t2);
  x1 = cadd(t1, t3);
  x2 = csub(t0, t2);
  x3 = csub(t1, t3);
}
// forward 8-point DFT, natural order in and out (radix-2 DIF split, then two 4-point DFTs)
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  const float r = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
  b1 = make_float2(r * (b1.x + b1.y), r * (b1.y - b1.x));    // * (1 - i)/sqrt2
  b2 = mul_mi(b2);                                          // * (-i)
  b3 = make_float2(r * (b3.y - b3.x), -r * (b3.x + b3.y));   // * (-1 - i)/sqrt2
  dft4(a0, a1, a2, a3);
  dft4(b0, b1, b2, b3);
  v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
  v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// 256-point forward complex FFT of the 8 values each lane holds (element m = lane + 32 j),
// followed by the DCT-IV post twiddle.  Results are scattered to stage[] (512 floats, aliasing
// the exchange buffer) as the DCT-IV output: stage[2k] = Re y_k, stage[511-2k] = -Im y_k.
//   n = 32 n1 + 4 n2 + n3   (lane = 4 n2 + n3, j = n1)      k = k1 + 8 k2 + 64 k3
__device__ __forceinline__ void fft256_dct4_tail(float2 (&c)[8], float2* scr, const float2* s_post, const float2* s_w64,
                                                 const float2* s_w256, int lane) {
  // pass 1: radix-8 over n1, twiddle W64^(n2 k1)
  dft8(c);
  {
    const int n2 = lane >> 2;
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) c[k1] = cmul(c[k1], s_w64[(n2 * k1) & 63]);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) scr[k1 * SCR_STRIDE + lane] = c[k1];
  }
  __syncwarp();
  // pass 2: lane = (k1, n3); radix-8 over n2, twiddle W256^(n3 (k1 + 8 k2))
  {
    const int k1 = lane >> 2, n3 = lane & 3;
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) c[n2] = scr[k1 * SCR_STRIDE + 4 * n2 + n3];
    dft8(c);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) c[k2] = cmul(c[k2], s_w256[(n3 * (k1 + 8 * k2)) & 255]);
    __syncwarp();
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) scr[k1 * SCR_STRIDE + 4 * k2 + n3] = c[k2];
  }
  __syncwarp();
  // pass 3: two (k1,k2) pairs per lane, radix-4 over n3, post twiddle
  float2 y[8];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int q = lane + 32 * half;
    const int k1 = q >> 3, k2 = q & 7;
    const float4* p = reinterpret_cast<const float4*>(scr + k1 * SCR_STRIDE + 4 * k2);
    const float4 u0 = p[0], u1 = p[1];
    float2 x0 = make_float2(u0.x, u0.y), x1 = make_float2(u0.z, u0.w), x2 = make_float2(u1.x, u1.y),
           x3 = make_float2(u1.z, u1.w);
    dft4(x0, x1, x2, x3);
    const int k = k1 + 8 * k2;
    y[4 * half + 0] = cmul(x0, s_post[k]);
    y[4 * half + 1] = cmul(x1, s_post[k + 64]);
    y[4 * half + 2] = cmul(x2, s_post[k + 128]);
    y[4 * half + 3] = cmul(x3, s_post[k + 192]);
  }
  __syncwarp();
  float* stage = reinterpret_cast<float*>(scr);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int q = lane + 32 * half;
    const int k = (q >> 3) + 8 * (q & 7);
#pragma unroll
    for (int k3 = 0; k3 < 4; ++k3) {
      const int kk = k + 64 * k3;
      stage[2 * kk] = y[4 * half + k3].x;
      stage[FFT_N - 1 - 2 * kk] = -y[4 * half + k3].y;
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------ forward, N = 512
// grid (ceil(nf / frames_per_cta), B); 8 warps; dynamic smem:
//   tables (window 1024 f, pre/post 256 f2 each, w64 64 f2, w256 256 f2) | per-warp exchange | input segment
__global__ void __launch_bounds__(MDCT_WARPS * 32)
mdct512_kernel(const float* __restrict__ x, float* __restrict__ X, FftTables tab, StridedIO io, int64_t T, int64_t nf,
               int hop, int frames_per_cta, int seg_len) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* s_win = reinterpret_cast<float*>(smem_raw);
  float2* s_pre = reinterpret_cast<float2*>(s_win + 2 * FFT_N);
  float2* s_post = s_pre + FFT_H;
  float2* s_w64 = s_post + FFT_H;
  float2* s_w256 = s_w64 + 64;
  float2* s_scr = s_w256 + 256;
  float* s_seg = reinterpret_cast<float*>(s_scr + MDCT_WARPS * SCR_F2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.y;
  const int64_t f0 = (int64_t)blockIdx.x * frames_per_cta;
  const int nframes = (int)min((int64_t)frames_per_cta, nf - f0);

  for (int i = tid; i < 2 * FFT_N; i += blockDim.x) s_win[i] = tab.window[i];
  for (int i = tid; i < FFT_H; i += blockDim.x) {
    s_pre[i] = tab.pre[i];
    s_post[i] = tab.post[i];
    s_w256[i] = tab.w256[i];
  }
  if (tid < 64) s_w64[tid] = tab.w64[tid];
  // stage the input segment once; frames overlap inside it.  Zero beyond T (implicit padding).
  {
    const int64_t s0 = f0 * hop;
    const int need = (nframes - 1) * hop + 2 * FFT_N;
    const float* xb = x + b * io.in_clip_stride;
    for (int i = tid; i < need; i += blockDim.x) {
      const int64_t s = s0 + i;
      s_seg[i] = (s < T) ? __ldg(xb + s * io.in_elem_stride) : 0.f;
    }
  }
  __syncthreads();

  float2* scr = s_scr + warp * SCR_F2;
  for (int f = warp; f < nframes; f += MDCT_WARPS) {
    const float* z = s_seg + f * hop;
    float2 c[8];
    // fold + window + pre-twiddle: c[m] = (u[2m] + i u[N-1-2m]) * pre[m],  h = N/2 = 256
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = lane + 32 * j;
      float re, im;
      if (j < 4) {  // m < 128
        const int a0 = 3 * FFT_H - 1 - 2 * m, a1 = 3 * FFT_H + 2 * m, a2 = FFT_H - 1 - 2 * m, a3 = FFT_H + 2 * m;
        re = -z[a0] * s_win[a0] - z[a1] * s_win[a1];
        im = z[a2] * s_win[a2] - z[a3] * s_win[a3];
      } else {
        const int a0 = 2 * m - FFT_H, a1 = 3 * FFT_H - 1 - 2 * m, a2 = FFT_H + 2 * m, a3 = 5 * FFT_H - 1 - 2 * m;
        re = z[a0] * s_win[a0] - z[a1] * s_win[a1];
        im = -z[a2] * s_win[a2] - z[a3] * s_win[a3];
      }
      c[j] = cmul(make_float2(re, im), s_pre[m]);
    }
    fft256_dct4_tail(c, scr, s_post, s_w64, s_w256, lane);
    const float4* st = reinterpret_cast<const float4*>(scr);
    float* dst = X + b * io.out_clip_stride + (f0 + f) * io.out_elem_stride;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(dst)[lane + 32 * i] = st[lane + 32 * i];
    } else {
      const float* sf = reinterpret_cast<const float*>(scr);
      for (int i = lane; i < FFT_N; i += 32) dst[i] = sf[i];
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ inverse, N = 512
// grid (ceil(L / samples_per_cta), B).  Phase 1: DCT-IV of every frame touching the CTA's output range
// into s_v[frame][N].  Phase 2: each output sample gathers unfold(v_i)[s - i*hop] * w * 2/N over its frames.
__global__ void __launch_bounds__(MDCT_WARPS * 32)
imdct512_kernel(const float* __restrict__ X, float* __restrict__ y, FftTables tab, StridedIO io, int64_t nf, int64_t L,
                int hop, int samples_per_cta, int max_frames) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* s_win = reinterpret_cast<float*>(smem_raw);
  float2* s_pre = reinterpret_cast<float2*>(s_win + 2 * FFT_N);
  float2* s_post = s_pre + FFT_H;
  float2* s_w64 = s_post + FFT_H;
  float2* s_w256 = s_w64 + 64;
  float2* s_scr = s_w256 + 256;
  float* s_v = reinterpret_cast<float*>(s_scr + MDCT_WARPS * SCR_F2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.y;
  const int64_t s0 = (int64_t)blockIdx.x * samples_per_cta;
  const int64_t s1 = min(L, s0 + samples_per_cta);
  // frames i with i*hop <= s < i*hop + 2N for some s in [s0, s1)
  int64_t i_lo = (s0 - 2 * FFT_N + 1 + hop - 1);
  i_lo = i_lo <= 0 ? 0 : i_lo / hop;
  const int64_t i_hi = min(nf - 1, (s1 - 1) / hop);
  const int nframes = (int)(i_hi - i_lo + 1);  // <= max_frames by construction

  for (int i = tid; i < 2 * FFT_N; i += blockDim.x) s_win[i] = tab.window[i];
  for (int i = tid; i < FFT_H; i += blockDim.x) {
    s_pre[i] = tab.pre[i];
    s_post[i] = tab.post[i];
    s_w256[i] = tab.w256[i];
  }
  if (tid < 64) s_w64[tid] = tab.w64[tid];
  __syncthreads();

  float2* scr = s_scr + warp * SCR_F2;
  for (int f = warp; f < nframes; f += MDCT_WARPS) {
    const float* src = X + b * io.in_clip_stride + (i_lo + f) * io.in_elem_stride;
    float* stage = reinterpret_cast<float*>(scr);
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<float4*>(stage)[lane + 32 * i] = __ldg(reinterpret_cast<const float4*>(src) + lane + 32 * i);
    } else {
      for (int i = lane; i < FFT_N; i += 32) stage[i] = __ldg(src + i);
    }
    __syncwarp();
    float2 c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = lane + 32 * j;
      c[j] = cmul(make_float2(stage[2 * m], stage[FFT_N - 1 - 2 * m]), s_pre[m]);
    }
    __syncwarp();
    fft256_dct4_tail(c, scr, s_post, s_w64, s_w256, lane);
    float4* dst = reinterpret_cast<float4*>(s_v + f * FFT_N);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[lane + 32 * i] = reinterpret_cast<const float4*>(scr)[lane + 32 * i];
    __syncwarp();
  }
  __syncthreads();

  const float scale = 2.0f / FFT_N;
  float* yb = y + b * io.out_clip_stride;
  for (int64_t s = s0 + tid; s < s1; s += blockDim.x) {
    int64_t lo = s - 2 * FFT_N + 1 + hop - 1;
    lo = lo <= 0 ? 0 : lo / hop;
    const int64_t hi = min(nf - 1, s / hop);
    float acc = 0.f;
    for (int64_t i = lo; i <= hi; ++i) {
      const int n = (int)(s - i * hop);
      const float* v = s_v + (i - i_lo) * FFT_N;
      float val;
      if (n < FFT_H) val = v[FFT_H + n];
      else if (n < 3 * FFT_H) val = -v[3 * FFT_H - 1 - n];
      else val = -v[n - 3 * FFT_H];
      acc = fmaf(val, s_win[n], acc);
    }
    yb[s * io.out_elem_stride] = acc * scale;
  }
}

// ------------------------------------------------------------------ N = 512, hop = 256 fast path
// The shipped configuration (window_size 512, hop N/2).  Sixteen lanes own one frame; each lane keeps 16 complex
// points in registers and the 256-point FFT is two radix-16 passes with ONE 16x16 transpose through shared
// memory (17-float2 row pitch: conflict free).  Every per-lane constant (window x pre-twiddle products, inter-pass
// twiddles) lives in registers for the whole CTA; the post twiddles sit in shared memory.  The input segment of a
// CTA's 32 frames is staged once, split into even / odd samples with a 16-word skew per 128 words, so the fold's
// stride-2 reads become unit-stride and the two frames of a warp hit disjoint banks.  The interleaved DCT-IV output
// X[2k] = Re y_k, X[2k+1] = -Im y_{255-k} is assembled with one lane-mirror shuffle per value and leaves the SM as
// full 128-byte lines.  ~75 shared-memory wavefronts and ~400 issue slots per frame: HBM is the bound.
constexpr int F2_THREADS = 128;
constexpr int F2_FRAMES = 32;                  // frames per CTA (forward)
constexpr int F2_EX = 16 * 17;                 // float2 per half-warp exchange buffer
constexpr int F2_SEG_WORDS = ((F2_FRAMES - 1) * 128 + 512) / 128 * 144;  // skewed even (or odd) plane of a segment
constexpr int I2_BLOCKS = 29;                  // 256-sample output blocks per CTA (inverse): 29 + 3 halo frames = 32
constexpr int I2_FRAMES = I2_BLOCKS + 3;

// forward 16-point DFT in registers: natural order in, output bin k at index 4 * (k % 4) + k / 4
__device__ __forceinline__ void fft16(float2 (&x)[16]) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r = 0.70710678118654752f;
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4(x[b], x[4 + b], x[8 + b], x[12 + b]);   // over a; t[b][c] at x[4c + b]
  // t[b][c] *= W16^(b c)
  x[5] = cmul(x[5], make_float2(c1, -s1));    // c=1,b=1
  x[6] = cmul(x[6], make_float2(r, -r));      // c=1,b=2
  x[7] = cmul(x[7], make_float2(s1, -c1));    // c=1,b=3
  x[9] = cmul(x[9], make_float2(r, -r));      // c=2,b=1
  x[10] = mul_mi(x[10]);                      // c=2,b=2  W16^4 = -i
  x[11] = cmul(x[11], make_float2(-r, -r));   // c=2,b=3  W16^6
  x[13] = cmul(x[13], make_float2(s1, -c1));  // c=3,b=1  W16^3
  x[14] = cmul(x[14], make_float2(-r, -r));   // c=3,b=2  W16^6
  x[15] = cmul(x[15], make_float2(-c1, s1));  // c=3,b=3  W16^9
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);  // over b; X[c + 4d] at x[4c + d]
}
__host__ __device__ constexpr int fidx(int k) { return 4 * (k & 3) + (k >> 2); }

// c[j] holds point m = 16 j + ln on entry; on exit c[fidx(k2)] = FFT256(c)[ln + 16 k2] * post[ln + 16 k2]
template <class PostT>
__device__ __forceinline__ void fft256_lanes16(float2 (&c)[16], const float2 (&tw)[16], float2* ex, const PostT& post, int ln) {
  fft16(c);
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) c[fidx(k1)] = cmul(c[fidx(k1)], tw[k1]);
  __syncwarp();
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) ex[k1 * 17 + ln] = c[fidx(k1)];
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) c[n2] = ex[ln * 17 + n2];
  fft16(c);
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) c[fidx(k2)] = cmul(c[fidx(k2)], post(k2));
}
struct PostSmem {   // post twiddles in shared memory (forward: registers are spent on the fold constants)
  const float2* p;
  __device__ __forceinline__ float2 operator()(int k2) const { return p[16 * k2]; }
};
struct PostRegs {   // post twiddles in registers
  float2 v[16];
  __device__ __forceinline__ float2 operator()(int k2) const { return v[k2]; }
};

// fold + window + pre-twiddle of one frame whose even / odd sample planes start at E / O (constants folded into kf):
// plane index r -> r + 16 (r >> 7); a 16-lane group never straddles a 128-word block
__device__ __forceinline__ void fold_frame(const float* E, const float* O, const float (&kf)[16][4], float2 (&c)[16], int ln) {
#define MFAC_SK(r) ((r) + 16 * ((r) >> 7))
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float a = O[MFAC_SK(383 - 16 * j) - ln], bb = E[MFAC_SK(384 + 16 * j) + ln];
    const float cc = O[MFAC_SK(127 - 16 * j) - ln], d = E[MFAC_SK(128 + 16 * j) + ln];
    c[j].x = -a * kf[j][0] - bb * kf[j][2] - cc * kf[j][3] + d * kf[j][1];
    c[j].y = -a * kf[j][1] - bb * kf[j][3] + cc * kf[j][2] - d * kf[j][0];
  }
#pragma unroll
  for (int j = 8; j < 16; ++j) {
    const float p = E[MFAC_SK(16 * j - 128) + ln], q = O[MFAC_SK(383 - 16 * j) - ln];
    const float r = E[MFAC_SK(128 + 16 * j) + ln], t = O[MFAC_SK(639 - 16 * j) - ln];
    c[j].x = p * kf[j][0] - q * kf[j][2] + r * kf[j][3] + t * kf[j][1];
    c[j].y = p * kf[j][1] - q * kf[j][3] - r * kf[j][2] - t * kf[j][0];
  }
#undef MFAC_SK
}
// per-lane constants of the forward transform: (window x pre-twiddle) products of the 16 points this lane folds
__device__ __forceinline__ void load_fold_constants(const FftTables& tab, float (&kf)[16][4], float2 (&tw)[16], int ln) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int m = 16 * j + ln;
    const float2 pre = tab.pre[m];
    const float w1 = j < 8 ? tab.window[FFT_H + 2 * m] : tab.window[2 * m - FFT_H];
    const float w2 = j < 8 ? tab.window[FFT_H - 1 - 2 * m] : tab.window[3 * FFT_H - 1 - 2 * m];
    kf[j][0] = w1 * pre.x; kf[j][1] = w1 * pre.y; kf[j][2] = w2 * pre.x; kf[j][3] = w2 * pre.y;
    tw[j] = tab.w256[(ln * j) & 255];
  }
}

__global__ void __launch_bounds__(F2_THREADS, 3)
mdct512h256_kernel(const float* __restrict__ x, float* __restrict__ X, FftTables tab, int64_t T, int64_t nf,
                   int64_t x_clip_stride, int64_t X_clip_stride) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* sE = reinterpret_cast<float*>(smem_raw);
  float* sO = sE + F2_SEG_WORDS;
  float2* sEx = reinterpret_cast<float2*>(sO + F2_SEG_WORDS);
  float2* sPost = sEx + (F2_THREADS / 16) * F2_EX;
  const int tid = threadIdx.x, lane = tid & 31, ln = tid & 15, hw = tid >> 4;
  const int64_t b = blockIdx.y;
  const int64_t f0 = (int64_t)blockIdx.x * F2_FRAMES;
  const int nframes = (int)min((int64_t)F2_FRAMES, nf - f0);

  // stage the segment: every thread first issues ALL of its 16-byte loads (one round trip), then splits them into
  // the even plane sE and the odd plane sO; word q of a plane sits at q + 16 (q >> 7).  Samples >= T are zero.
  {
    const int64_t s0 = f0 * FFT_H;
    const int need = (nframes - 1) * FFT_H + 2 * FFT_N;           // multiple of 256
    const float* xb = x + b * x_clip_stride + s0;
    constexpr int NV = (((F2_FRAMES - 1) * FFT_H + 2 * FFT_N) / 4 + F2_THREADS - 1) / F2_THREADS;  // float4 per thread
    float4 v[NV];
    if ((reinterpret_cast<uintptr_t>(xb) & 15) == 0 && s0 + need <= T) {
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int i4 = (tid + u * F2_THREADS) * 4;
        v[u] = i4 < need ? __ldg(reinterpret_cast<const float4*>(xb + i4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int i4 = (tid + u * F2_THREADS) * 4;
        const int64_t rem = T - s0 - i4;                          // samples available from this position
        v[u].x = (i4 < need && rem > 0) ? __ldg(xb + i4) : 0.f;
        v[u].y = (i4 < need && rem > 1) ? __ldg(xb + i4 + 1) : 0.f;
        v[u].z = (i4 < need && rem > 2) ? __ldg(xb + i4 + 2) : 0.f;
        v[u].w = (i4 < need && rem > 3) ? __ldg(xb + i4 + 3) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int q = (tid + u * F2_THREADS) * 2;
      if (2 * q < need) {
        const int ph = q + 16 * (q >> 7);
        *reinterpret_cast<float2*>(sE + ph) = make_float2(v[u].x, v[u].z);
        *reinterpret_cast<float2*>(sO + ph) = make_float2(v[u].y, v[u].w);
      }
    }
  }
  float kf[16][4];
  float2 tw[16];
  load_fold_constants(tab, kf, tw, ln);
  sPost[tid] = tab.post[tid];
  sPost[tid + 128] = tab.post[tid + 128];

  __syncthreads();

  float2* ex = sEx + hw * F2_EX;
  // CTA-uniform trip count (so the shuffles below need no reconvergence code); a half-warp whose frame lies beyond
  // the clip shadows the last valid frame and skips the store
  const int n_it = (nframes + F2_THREADS / 16 - 1) / (F2_THREADS / 16);
#pragma unroll 1
  for (int it = 0; it < n_it; ++it) {
    const int f = it * (F2_THREADS / 16) + hw;
    const int fc = f < nframes ? f : nframes - 1;
    float2 c[16];
    fold_frame(sE + fc * 144, sO + fc * 144, kf, c, ln);
    fft256_lanes16(c, tw, ex, PostSmem{sPost + ln}, ln);
    // X[2k] = Re y_k, X[2k+1] = -Im y_{255-k};  y_{255-k} lives in lane 15 - ln, register 15 - k2
    float2* dst = reinterpret_cast<float2*>(X + b * X_clip_stride + (f0 + fc) * FFT_N) + ln;
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const float im = __shfl_xor_sync(0xffffffffu, c[fidx(15 - k2)].y, 15);
      if (f < nframes) dst[16 * k2] = make_float2(c[fidx(k2)].x, -im);
    }
  }
}

// Short clips (nf <= 16, e.g. the 784-sample MNIST-shaped rows of the training step): one CTA packs several clips so
// that its 32 frame slots stay busy.  Slot s -> clip s / nf, frame s % nf; every clip owns its own skewed segment.
__global__ void __launch_bounds__(F2_THREADS, 2)
mdct512h256_short_kernel(const float* __restrict__ x, float* __restrict__ X, FftTables tab, int64_t T, int nf, int cpc, int64_t B,
                         int64_t x_clip_stride, int64_t X_clip_stride) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int need = (nf - 1) * FFT_H + 2 * FFT_N;     // samples a clip's frames touch
  const int pw = need / 2 / 128 * 144;               // skewed words per plane per clip
  float* sE = reinterpret_cast<float*>(smem_raw);
  float* sO = sE + cpc * pw;
  float2* sEx = reinterpret_cast<float2*>(sO + cpc * pw);
  float2* sPost = sEx + (F2_THREADS / 16) * F2_EX;
  const int tid = threadIdx.x, ln = tid & 15, hw = tid >> 4;
  const int64_t b0 = (int64_t)blockIdx.x * cpc;
  const int nclips = (int)min((int64_t)cpc, B - b0);
  const int nframes = nclips * nf;
  const int per_clip4 = need / 4;
  const int total4 = nclips * per_clip4;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((x_clip_stride & 3) == 0);
  constexpr int BATCH = 10;
  for (int g0 = 0; g0 < total4; g0 += BATCH * F2_THREADS) {
    float4 v[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int g = g0 + tid + u * F2_THREADS;
      const int cl = g / per_clip4, i4 = (g - cl * per_clip4) * 4;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < total4) {
        const float* src = x + (b0 + cl) * x_clip_stride + i4;
        if (aligned && i4 + 3 < T) v[u] = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (i4 < T) v[u].x = __ldg(src);
          if (i4 + 1 < T) v[u].y = __ldg(src + 1);
          if (i4 + 2 < T) v[u].z = __ldg(src + 2);
          if (i4 + 3 < T) v[u].w = __ldg(src + 3);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int g = g0 + tid + u * F2_THREADS;
      if (g < total4) {
        const int cl = g / per_clip4, q = (g - cl * per_clip4) * 2;
        const int ph = cl * pw + q + 16 * (q >> 7);
        *reinterpret_cast<float2*>(sE + ph) = make_float2(v[u].x, v[u].z);
        *reinterpret_cast<float2*>(sO + ph) = make_float2(v[u].y, v[u].w);
      }
    }
  }
  float kf[16][4];
  float2 tw[16];
  load_fold_constants(tab, kf, tw, ln);
  sPost[tid] = tab.post[tid];
  sPost[tid + 128] = tab.post[tid + 128];
  __syncthreads();

  float2* ex = sEx + hw * F2_EX;
  const int n_it = (nframes + F2_THREADS / 16 - 1) / (F2_THREADS / 16);
#pragma unroll 1
  for (int it = 0; it < n_it; ++it) {
    const int sl = it * (F2_THREADS / 16) + hw;
    const int sc = sl < nframes ? sl : nframes - 1;
    const int cl = sc / nf, fr = sc - cl * nf;
    float2 c[16];
    fold_frame(sE + cl * pw + fr * 144, sO + cl * pw + fr * 144, kf, c, ln);
    fft256_lanes16(c, tw, ex, PostSmem{sPost + ln}, ln);
    float2* dst = reinterpret_cast<float2*>(X + (b0 + cl) * X_clip_stride + (int64_t)fr * FFT_N) + ln;
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const float im = __shfl_xor_sync(0xffffffffu, c[fidx(15 - k2)].y, 15);
      if (sl < nframes) dst[16 * k2] = make_float2(c[fidx(k2)].x, -im);
    }
  }
}


// ------------------------------------------------------------------ fused tokenise + iMF prologue
// The training loop tokenises a batch and hands the tokens to the loss (trainers/train.py:339-345).  For short clips
// (nf <= 16 frames per row: the 784- / 2048- / 4096-sample rows of the benchmarks) this kernel is the short-clip MDCT above
// with the iMF prologue (imf_prep_kernel) as its store stage: a coefficient pair never goes to HBM as a token -- it leaves the
// registers as z_t, as the loss target noise_max e - x, and as the encoder's bf16 input row.  (e, t, r) are the same draws
// imf_prep_kernel makes for the same (seed, step, row): the noise quad of four consecutive columns is computed by one lane of
// an even / odd lane pair and shared with a shuffle.
__global__ void __launch_bounds__(F2_THREADS, 3)
tokenize_prep_short_kernel(const float* __restrict__ x, FftTables tab, int64_t T, int nf, int cpc, int64_t x_clip_stride,
                           PrepArgs a, Dims d) {
  MFAC_PDL_SYNC();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int need = (nf - 1) * FFT_H + 2 * FFT_N;     // samples a clip's frames touch
  const int pw = need / 2 / 128 * 144;               // skewed words per plane per clip
  float* sE = reinterpret_cast<float*>(smem_raw);
  float* sO = sE + cpc * pw;
  float2* sEx = reinterpret_cast<float2*>(sO + cpc * pw);
  float2* sPost = sEx + (F2_THREADS / 16) * F2_EX;
  __shared__ float s_t[32], s_r[32];
  const int tid = threadIdx.x, ln = tid & 15, hw = tid >> 4;
  const int64_t B = a.B;
  const int64_t b0 = (int64_t)blockIdx.x * cpc;
  const int nclips = (int)min((int64_t)cpc, B - b0);
  const int nframes = nclips * nf;
  const int per_clip4 = need / 4;
  const int total4 = nclips * per_clip4;
  const uint64_t step = a.cfg.step_dev ? *a.cfg.step_dev : a.cfg.step;
  if (tid < nclips) {
    float t, r;
    draw_tr(a, b0 + tid, step, t, r);
    s_t[tid] = t; s_r[tid] = r;
    a.t[b0 + tid] = t; a.r[b0 + tid] = r;
  }
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((x_clip_stride & 3) == 0);
  constexpr int BATCH = 10;
  for (int g0 = 0; g0 < total4; g0 += BATCH * F2_THREADS) {
    float4 v[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int g = g0 + tid + u * F2_THREADS;
      const int cl = g / per_clip4, i4 = (g - cl * per_clip4) * 4;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < total4) {
        const float* src = x + (b0 + cl) * x_clip_stride + i4;
        if (aligned && i4 + 3 < T) v[u] = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (i4 < T) v[u].x = __ldg(src);
          if (i4 + 1 < T) v[u].y = __ldg(src + 1);
          if (i4 + 2 < T) v[u].z = __ldg(src + 2);
          if (i4 + 3 < T) v[u].w = __ldg(src + 3);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int g = g0 + tid + u * F2_THREADS;
      if (g < total4) {
        const int cl = g / per_clip4, q = (g - cl * per_clip4) * 2;
        const int ph = cl * pw + q + 16 * (q >> 7);
        *reinterpret_cast<float2*>(sE + ph) = make_float2(v[u].x, v[u].z);
        *reinterpret_cast<float2*>(sO + ph) = make_float2(v[u].y, v[u].w);
      }
    }
  }
  float kf[16][4];
  float2 tw[16];
  load_fold_constants(tab, kf, tw, ln);
  sPost[tid] = tab.post[tid];
  sPost[tid + 128] = tab.post[tid + 128];
  __syncthreads();
  // conditioning rows of this CTA's clips (every thread takes part in each row)
  for (int cl = 0; cl < nclips; ++cl) {
    const float t = s_t[cl], r = s_r[cl];
    if (a.cond_v) write_cond_row(t, 0.f, d.C, d.Cp, a.cond_v + (b0 + cl) * d.Cp, nullptr);
    write_cond_row(t, t - r, d.C, d.Cp, a.cond_u + (b0 + cl) * d.Cp, a.dcond_u + (b0 + cl) * d.Cp);
  }

  float2* ex = sEx + hw * F2_EX;
  const int n_it = (nframes + F2_THREADS / 16 - 1) / (F2_THREADS / 16);
  const bool odd = (ln & 1) != 0;
#pragma unroll 1
  for (int it = 0; it < n_it; ++it) {
    const int sl = it * (F2_THREADS / 16) + hw;
    const int sc = sl < nframes ? sl : nframes - 1;
    const int cl = sc / nf, fr = sc - cl * nf;
    float2 c[16];
    fold_frame(sE + cl * pw + fr * 144, sO + cl * pw + fr * 144, kf, c, ln);
    fft256_lanes16(c, tw, ex, PostSmem{sPost + ln}, ln);
    const int64_t b = b0 + cl;
    const float t = s_t[cl];
    const float omt = 1.0f - t, nscale = a.cfg.noise_min + a.cfg.noise_max * t, nmax = a.cfg.noise_max;
    const int64_t row = b * d.Dp + (int64_t)fr * FFT_N + 2 * ln;     // + 32 k2: this lane's coefficient pair k2
    // tokens X[2k] = Re y_k, X[2k+1] = -Im y_{255-k};  y_{255-k} lives in lane 15 - ln, register 15 - k2.  The spectrum goes
    // back through the half-warp's exchange buffer so that the store stage is a ROLLED loop (one Philox call, six stores per
    // trip): unrolled it was 70-100 KB of SASS, far beyond the instruction cache.
    __syncwarp();
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) ex[k2 * 17 + ln] = c[fidx(k2)];
    __syncwarp();
#pragma unroll 1
    for (int kp = 0; kp < 16; kp += 2) {
      float2 xv[2], ev[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int k2 = kp + u;
        xv[u] = make_float2(ex[k2 * 17 + ln].x, -ex[(15 - k2) * 17 + (15 - ln)].y);
      }
      if (a.e_in) {
#pragma unroll
        for (int u = 0; u < 2; ++u) ev[u] = *reinterpret_cast<const float2*>(a.e_in + row + 32 * (kp + u));
      } else {
        // noise quad of columns [4 q, 4 q + 4) = this lane pair's two coefficient pairs of k2: the even lane draws the quad of
        // k2 = kp, the odd lane the quad of k2 = kp + 1; each keeps its half and hands the other half to its neighbour
        const int k2m = kp + (odd ? 1 : 0);
        const uint64_t col4 = (uint64_t)((int64_t)fr * FFT_N + 32 * k2m + 2 * (ln & ~1));
        const uint64_t idx = ((a.cfg.row_offset + (uint64_t)b) * (uint64_t)d.Dp + col4) >> 2;
        const float4 n4 = philox_normal4(idx, 0u, a.cfg.seed, step);
        const float2 keep = odd ? make_float2(n4.z, n4.w) : make_float2(n4.x, n4.y);
        const float2 give = odd ? make_float2(n4.x, n4.y) : make_float2(n4.z, n4.w);
        const float gx = __shfl_xor_sync(0xffffffffu, give.x, 1), gy = __shfl_xor_sync(0xffffffffu, give.y, 1);
        ev[0] = odd ? make_float2(gx, gy) : keep;
        ev[1] = odd ? keep : make_float2(gx, gy);
      }
      if (sl < nframes) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int64_t at = row + 32 * (kp + u);
          const float2 zt = make_float2(zt_of(omt, xv[u].x, nscale, ev[u].x), zt_of(omt, xv[u].y, nscale, ev[u].y));
          const float2 tg = make_float2(target_of(nmax, ev[u].x, xv[u].x), target_of(nmax, ev[u].y, xv[u].y));
          if (a.e) *reinterpret_cast<float2*>(a.e + at) = ev[u];
          if (a.z) *reinterpret_cast<float2*>(a.z + at) = zt;
          *reinterpret_cast<float2*>(a.z2 + at) = zt;
          *reinterpret_cast<float2*>(a.target + at) = tg;
          if (a.seed) *reinterpret_cast<float2*>(a.seed + at) = tg;
          *reinterpret_cast<uint32_t*>(a.xb + at) = pack_bf16(xv[u].x, xv[u].y);
        }
      }
    }
  }
}

// inverse: CTA = clip b, output blocks [q0, q0 + I2_BLOCKS) of 256 samples; phase 1 DCT-IV of the <= 32 frames that
// touch them into shared memory, phase 2 unfold + window + overlap-add as a float4 gather (4 frames per sample).
__global__ void __launch_bounds__(F2_THREADS, 2)
imdct512h256_kernel(const float* __restrict__ X, float* __restrict__ y, FftTables tab, int64_t nf, int64_t L,
                    int64_t X_clip_stride, int64_t y_clip_stride) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* sV = reinterpret_cast<float*>(smem_raw);                      // [I2_FRAMES][512]
  float2* sEx = reinterpret_cast<float2*>(sV + I2_FRAMES * FFT_N);
  const int tid = threadIdx.x, ln = tid & 15, hw = tid >> 4;
  const int64_t b = blockIdx.y;
  const int64_t q0 = (int64_t)blockIdx.x * I2_BLOCKS;
  const int64_t nblocks = nf + 3;                                      // L / 256
  const int64_t q1 = min(nblocks, q0 + I2_BLOCKS);
  const int64_t i_lo = max((int64_t)0, q0 - 3), i_hi = min(nf - 1, q1 - 1);
  const int nframes = (int)(i_hi - i_lo + 1);

  float2* ex = sEx + hw * F2_EX;
  constexpr int HW = F2_THREADS / 16;
  // the first frame's coefficients are requested before anything else; the constant tables load behind them
  float2 in[16];
  {
    const int fc0 = hw < nframes ? hw : nframes - 1;
    const float2* src = reinterpret_cast<const float2*>(X + b * X_clip_stride + (i_lo + fc0) * FFT_N) + ln;
#pragma unroll
    for (int j = 0; j < 16; ++j) in[j] = __ldg(src + 16 * j);          // {X[2m], X[2m+1]}, m = 16 j + ln
  }
  float2 pre[16], tw[16];
  PostRegs post;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    pre[j] = tab.pre[16 * j + ln];
    tw[j] = tab.w256[(ln * j) & 255];
    const float2 p = tab.post[ln + 16 * j];
    post.v[j] = make_float2(p.x * (2.0f / FFT_N), p.y * (2.0f / FFT_N));
  }
  const int n_it = (nframes + HW - 1) / HW;                          // CTA-uniform
#pragma unroll 1
  for (int it = 0; it < n_it; ++it) {
    const int f = it * HW + hw;
    const int fc = f < nframes ? f : nframes - 1;
    float2 c[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // X[511 - 2m] = X[2 (255 - m) + 1]: second half of lane 15 - ln, register 15 - j
      const float hi = __shfl_xor_sync(0xffffffffu, in[15 - j].y, 15);
      c[j] = cmul(make_float2(in[j].x, hi), pre[j]);
    }
    if (it + 1 < n_it) {                                               // next frame's coefficients fly during this FFT
      const int fn = (it + 1) * HW + hw;
      const int fcn = fn < nframes ? fn : nframes - 1;
      const float2* src = reinterpret_cast<const float2*>(X + b * X_clip_stride + (i_lo + fcn) * FFT_N) + ln;
#pragma unroll
      for (int j = 0; j < 16; ++j) in[j] = __ldg(src + 16 * j);
    }
    fft256_lanes16(c, tw, ex, post, ln);
    float2* dst = reinterpret_cast<float2*>(sV + fc * FFT_N) + ln;    // v[2k] = Re y_k, v[2k+1] = -Im y_{255-k}
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const float im = __shfl_xor_sync(0xffffffffu, c[fidx(15 - k2)].y, 15);
      if (f < nframes) dst[16 * k2] = make_float2(c[fidx(k2)].x, -im);
    }
  }
  __syncthreads();

  // gather: thread -> 4 consecutive samples r..r+3 of a block, two blocks per pass
  const int r = 4 * (tid & 63);
  float4 w0 = *reinterpret_cast<const float4*>(tab.window + r);
  float4 w1 = *reinterpret_cast<const float4*>(tab.window + 256 + r);
  float4 w2 = *reinterpret_cast<const float4*>(tab.window + 512 + r);
  float4 w3 = *reinterpret_cast<const float4*>(tab.window + 768 + r);
  float* yb = y + b * y_clip_stride;
  const bool vec = (reinterpret_cast<uintptr_t>(yb) & 15) == 0;
  // local block index lb = q - q0; frame slot of (q - t) in sV is lb + off - t with off = q0 - i_lo (0..3)
  const int nb = (int)(q1 - q0), off = (int)(q0 - i_lo);
  const int lastf = (int)(nf - 1 - i_lo);                  // last valid frame slot
  float* dst = yb + (q0 + (tid >> 6)) * 256 + r;
  for (int lb = tid >> 6; lb < nb; lb += 2, dst += 512) {
    const int sl = lb + off;                               // slot of frame q
    const float* vb = sV + sl * FFT_N;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sl <= lastf) {                                     // t = 0:  v[256 + r + j]
      const float4 v = *reinterpret_cast<const float4*>(vb + 256 + r);
      acc.x = v.x * w0.x; acc.y = v.y * w0.y; acc.z = v.z * w0.z; acc.w = v.w * w0.w;
    }
    if (sl >= 1 && sl - 1 <= lastf) {                      // t = 1: -v[511 - r - j]
      const float4 v = *reinterpret_cast<const float4*>(vb - FFT_N + 508 - r);
      acc.x -= v.w * w1.x; acc.y -= v.z * w1.y; acc.z -= v.y * w1.z; acc.w -= v.x * w1.w;
    }
    if (sl >= 2 && sl - 2 <= lastf) {                      // t = 2: -v[255 - r - j]
      const float4 v = *reinterpret_cast<const float4*>(vb - 2 * FFT_N + 252 - r);
      acc.x -= v.w * w2.x; acc.y -= v.z * w2.y; acc.z -= v.y * w2.z; acc.w -= v.x * w2.w;
    }
    if (sl >= 3 && sl - 3 <= lastf) {                      // t = 3: -v[r + j]
      const float4 v = *reinterpret_cast<const float4*>(vb - 3 * FFT_N + r);
      acc.x -= v.x * w3.x; acc.y -= v.y * w3.y; acc.z -= v.z * w3.z; acc.w -= v.w * w3.w;
    }
    if (vec) *reinterpret_cast<float4*>(dst) = acc;
    else { dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w; }
  }
}

// ------------------------------------------------------------------ inverse, streaming overlap-add in registers
// A half-warp walks along consecutive frames of one clip.  After the DCT-IV of frame i, lane ln holds v_i[p] at the 32
// positions p = 2 ln + e + 32 k2.  Each v_i[p] lands on exactly two output samples, and both belong either to this lane
// or to its mirror lane 15 - ln:
//     p >= 256:  block i   at r = p - 256  (+w[p-256])      block i+1 at r = 511 - p  (-w[767-p], mirror lane)
//     p <  256:  block i+3 at r = p        (-w[768+p])      block i+2 at r = 255 - p  (-w[767-p], mirror lane)
// so the overlap-add is three 16-value register accumulators per lane (blocks i+1..i+3) and one lane-mirror shuffle
// per value; block i is complete after frame i and leaves as eight 128-byte lines per half-warp.  No shared-memory
// frame store, no gather pass, no CTA barrier; a stream re-does 3 halo frames.  Shared memory holds only the exchange
// buffers and the per-lane constant tables (pre/post twiddles, window products).
constexpr int S2_TABLE_F2 = 256 + 256 + 128 + 128 + 256;   // pre | post | Wl_hi | Wl_lo | Wm   (float2 each)

__global__ void __launch_bounds__(F2_THREADS, 3)
imdct512h256_stream_kernel(const float* __restrict__ X, float* __restrict__ y, FftTables tab, int64_t nf, int64_t X_clip_stride,
                           int64_t y_clip_stride, int S, int streams_per_clip, int64_t total_streams) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float2* sEx = reinterpret_cast<float2*>(smem_raw);
  float2* sPre = sEx + (F2_THREADS / 16) * F2_EX;   // [j][ln]   pre[16 j + ln]
  float2* sPost = sPre + 256;                       // [k2][ln]  post[ln + 16 k2] * 2/N
  float2* sWhi = sPost + 256;                       // [kk][ln]  { w[r], w[r+1] },               r = 2 ln + 32 kk
  float2* sWlo = sWhi + 128;                        // [kk][ln]  {-w[768+r], -w[769+r]}
  float2* sWm = sWlo + 128;                         // [k2][ln]  {-w[767-p], -w[766-p]},         p = 2 ln + 32 k2
  const int tid = threadIdx.x, ln = tid & 15, hw = tid >> 4;
  for (int i = tid; i < 256; i += F2_THREADS) {
    const int l = i & 15, k = i >> 4;
    sPre[i] = tab.pre[16 * k + l];
    const float2 po = tab.post[l + 16 * k];
    sPost[i] = make_float2(po.x * (2.0f / FFT_N), po.y * (2.0f / FFT_N));
    const int p = 2 * l + 32 * k;
    sWm[i] = make_float2(-tab.window[767 - p], -tab.window[766 - p]);
    if (k < 8) {
      sWhi[i] = make_float2(tab.window[p], tab.window[p + 1]);
      sWlo[i] = make_float2(-tab.window[768 + p], -tab.window[769 + p]);
    }
  }
  float2 tw[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) tw[j] = tab.w256[(ln * j) & 255];
  __syncthreads();

  const int64_t sid = (int64_t)blockIdx.x * (F2_THREADS / 16) + hw;
  const bool valid = sid < total_streams;
  const int64_t sidc = valid ? sid : total_streams - 1;      // idle half-warps shadow a real stream (no stores)
  const int64_t b = sidc / streams_per_clip;
  const int64_t i0 = (sidc % streams_per_clip) * (int64_t)S;
  const int64_t nblocks = nf + 3;
  const float* Xb = X + b * X_clip_stride;
  float* yb = y + b * y_clip_stride;
  float2* ex = sEx + hw * F2_EX;

  float2 accA[8], accB[8], accC[8];                          // partial blocks i, i+1, i+2 (float2 over e)
#pragma unroll
  for (int k = 0; k < 8; ++k) accA[k] = accB[k] = accC[k] = make_float2(0.f, 0.f);

  auto load_frame = [&](int64_t fi, float2 (&in)[16]) {
    const bool ok = fi >= 0 && fi < nf;
    const float2* src = reinterpret_cast<const float2*>(Xb + (ok ? fi : 0) * FFT_N) + ln;
#pragma unroll
    for (int j = 0; j < 16; ++j) in[j] = ok ? __ldg(src + 16 * j) : make_float2(0.f, 0.f);
  };
  float2 in[16];
  load_frame(i0 - 3, in);

#pragma unroll 1
  for (int j = 0; j < S + 3; ++j) {
    const int64_t fi = i0 - 3 + j;
    float2 c[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float hi = __shfl_xor_sync(0xffffffffu, in[15 - q].y, 15);   // X[511 - 2m]
      c[q] = cmul(make_float2(in[q].x, hi), sPre[16 * q + ln]);
    }
    load_frame(fi + 1, in);                                              // flies during this frame's FFT
    fft256_lanes16(c, tw, ex, PostSmem{sPost + ln}, ln);
    // v[k2] = { v_i[2 ln + 32 k2], v_i[2 ln + 32 k2 + 1] }
    float2 v[16];
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const float im = __shfl_xor_sync(0xffffffffu, c[fidx(15 - k2)].y, 15);
      v[k2] = make_float2(c[fidx(k2)].x, -im);
    }
    // block i is complete: t = 0 term of this frame on top of accA
    if (valid && j >= 3 && fi < nblocks) {
      float2* dst = reinterpret_cast<float2*>(yb + fi * 256) + ln;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const float2 w = sWhi[16 * kk + ln];
        dst[16 * kk] = make_float2(fmaf(w.x, v[kk + 8].x, accA[kk].x), fmaf(w.y, v[kk + 8].y, accA[kk].y));
      }
    }
    // mirror terms u[k2] = -w[767 - p] v[p]; the value for (kk, e) comes from the mirror lane's (15 - kk | 7 - kk, 1 - e)
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const float2 wm1 = sWm[16 * (15 - kk) + ln], wm2 = sWm[16 * (7 - kk) + ln], wl = sWlo[16 * kk + ln];
      const float u1x = __shfl_xor_sync(0xffffffffu, wm1.y * v[15 - kk].y, 15);   // -> e' = 0 of block i+1
      const float u1y = __shfl_xor_sync(0xffffffffu, wm1.x * v[15 - kk].x, 15);   // -> e' = 1
      const float u2x = __shfl_xor_sync(0xffffffffu, wm2.y * v[7 - kk].y, 15);    // -> block i+2
      const float u2y = __shfl_xor_sync(0xffffffffu, wm2.x * v[7 - kk].x, 15);
      accA[kk] = make_float2(accB[kk].x + u1x, accB[kk].y + u1y);
      accB[kk] = make_float2(accC[kk].x + u2x, accC[kk].y + u2y);
      accC[kk] = make_float2(wl.x * v[kk].x, wl.y * v[kk].y);                     // t = 3 term opens block i+3
    }
  }
}

constexpr int S2_SMEM = ((F2_THREADS / 16) * F2_EX + S2_TABLE_F2) * 8;

constexpr int F2_SMEM = 2 * F2_SEG_WORDS * 4 + (F2_THREADS / 16) * F2_EX * 8 + 256 * 8;
constexpr int I2_SMEM = I2_FRAMES * FFT_N * 4 + (F2_THREADS / 16) * F2_EX * 8;

// ------------------------------------------------------------------ generic N: dense contraction
constexpr int DENSE_FR = 4;       // frames per CTA
constexpr int DENSE_THREADS = 128;

__global__ void __launch_bounds__(DENSE_THREADS)
mdct_dense_kernel(const float* __restrict__ x, float* __restrict__ X, const float* __restrict__ wc, StridedIO io, int64_t T,
                  int64_t nf, int N, int hop) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* s_x = reinterpret_cast<float*>(smem_raw);  // [DENSE_FR][2N]
  const int64_t b = blockIdx.z;
  const int64_t f0 = (int64_t)blockIdx.y * DENSE_FR;
  const int nframes = (int)min((int64_t)DENSE_FR, nf - f0);
  const float* xb = x + b * io.in_clip_stride;
  for (int i = threadIdx.x; i < DENSE_FR * 2 * N; i += blockDim.x) {
    const int f = i / (2 * N), n = i % (2 * N);
    const int64_t s = (f0 + f) * hop + n;
    s_x[i] = (f < nframes && s < T) ? __ldg(xb + s * io.in_elem_stride) : 0.f;
  }
  __syncthreads();
  const int k = blockIdx.x * DENSE_THREADS + threadIdx.x;
  if (k >= N) return;
  float acc[DENSE_FR];
#pragma unroll
  for (int f = 0; f < DENSE_FR; ++f) acc[f] = 0.f;
  for (int n = 0; n < 2 * N; ++n) {
    const float c = __ldg(wc + (int64_t)n * N + k);
#pragma unroll
    for (int f = 0; f < DENSE_FR; ++f) acc[f] = fmaf(s_x[f * 2 * N + n], c, acc[f]);
  }
  for (int f = 0; f < nframes; ++f) X[b * io.out_clip_stride + (f0 + f) * io.out_elem_stride + k] = acc[f];
}

__global__ void __launch_bounds__(256)
imdct_dense_kernel(const float* __restrict__ X, float* __restrict__ y, const float* __restrict__ wct, StridedIO io,
                   int64_t nf, int64_t L, int N, int hop) {
  const int64_t b = blockIdx.y;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= L) return;
  int64_t lo = s - 2 * N + 1 + hop - 1;
  lo = lo <= 0 ? 0 : lo / hop;
  const int64_t hi = min(nf - 1, s / hop);
  float acc = 0.f;
  for (int64_t i = lo; i <= hi; ++i) {
    const int n = (int)(s - i * hop);
    const float* Xi = X + b * io.in_clip_stride + i * io.in_elem_stride;
    float a = 0.f;
    for (int k = 0; k < N; ++k) a = fmaf(__ldg(Xi + k), __ldg(wct + (int64_t)k * 2 * N + n), a);
    acc += a;
  }
  y[b * io.out_clip_stride + s * io.out_elem_stride] = acc;
}

constexpr int TABLE_BYTES = 2 * FFT_N * 4 + (FFT_H * 2 + 64 + 256) * 8;
constexpr int SCRATCH_BYTES = MDCT_WARPS * SCR_F2 * 8;

}  // namespace

int mdct_forward(const float* x, float* X, StridedIO io, int64_t B, int64_t T, int N, int hop, cudaStream_t stream) {
  if (!x || !X) return MFAC_ERR_NULL;
  if (B <= 0 || T <= 0 || N <= 0 || hop <= 0) return MFAC_ERR_BAD_SHAPE;
  if (B > 65535) return MFAC_ERR_UNSUPPORTED;
  const int64_t nf = T < N ? 1 : (T - N) / hop + 1;
  TableSet ts;
  // frames per CTA: as many as fit a ~48 KB input segment, at most 32
  int fpc = (int)((12288 - 2 * FFT_N) / hop + 1);
  fpc = fpc < 1 ? 1 : (fpc > 32 ? 32 : fpc);
  if (nf < fpc) fpc = (int)nf;
  const int64_t seg_len = (int64_t)(fpc - 1) * hop + 2 * FFT_N;
  if (N == FFT_N && hop == FFT_H && io.in_elem_stride == 1 && io.out_elem_stride == FFT_N &&
      (io.out_clip_stride % 2) == 0 && (reinterpret_cast<uintptr_t>(X) & 7) == 0) {
    MFAC_OK(get_tables(N, true, &ts));
    static PerDeviceOnce configured2;
    if (configured2.need()) {
      MFAC_CUDA_OK(cudaFuncSetAttribute(mdct512h256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
      configured2.done();
    }
    void* prof = profile_begin(MFAC_PROF_MDCT, 4.0 * (double)B * ((double)T + (double)nf * N), stream);
    if (nf <= F2_FRAMES / 2) {
      // short clips: pack clips into a CTA until its 32 frame slots are used
      const int cpc = F2_FRAMES / (int)nf;
      const int need = ((int)nf - 1) * FFT_H + 2 * FFT_N;
      const size_t smem_s = (size_t)2 * cpc * (need / 2 / 128 * 144) * 4 + (F2_THREADS / 16) * F2_EX * 8 + 256 * 8;
      static PerDeviceOnce configured3;
      if (configured3.need()) {
        MFAC_CUDA_OK(cudaFuncSetAttribute(mdct512h256_short_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured3.done();
      }
      mdct512h256_short_kernel<<<(unsigned)ceil_div<int64_t>(B, cpc), F2_THREADS, smem_s, stream>>>(
          x, X, ts.fft, T, (int)nf, cpc, B, io.in_clip_stride, io.out_clip_stride);
    } else {
      dim3 grid((unsigned)ceil_div<int64_t>(nf, F2_FRAMES), (unsigned)B);
      mdct512h256_kernel<<<grid, F2_THREADS, F2_SMEM, stream>>>(x, X, ts.fft, T, nf, io.in_clip_stride, io.out_clip_stride);
    }
    profile_end(prof, stream);
    count_launch();
    return launch_status();
  }
  if (N == FFT_N && seg_len <= 40960) {
    MFAC_OK(get_tables(N, true, &ts));
    const size_t smem = TABLE_BYTES + SCRATCH_BYTES + (size_t)seg_len * 4;
    static PerDeviceOnce configured;
    if (configured.need()) {
      MFAC_CUDA_OK(cudaFuncSetAttribute(mdct512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured.done();
    }
    dim3 grid((unsigned)ceil_div<int64_t>(nf, fpc), (unsigned)B);
    void* prof = profile_begin(MFAC_PROF_MDCT, 4.0 * (double)B * ((double)T + (double)nf * N), stream);
    mdct512_kernel<<<grid, MDCT_WARPS * 32, smem, stream>>>(x, X, ts.fft, io, T, nf, hop, fpc, (int)seg_len);
    profile_end(prof, stream);
    count_launch();
    return launch_status();
  }
  if (N > 4096) return MFAC_ERR_UNSUPPORTED;
  MFAC_OK(get_tables(N, false, &ts));
  const size_t smem = (size_t)DENSE_FR * 2 * N * 4;
  static PerDeviceOnce configured_dense;
  if (configured_dense.need()) {
    MFAC_CUDA_OK(cudaFuncSetAttribute(mdct_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured_dense.done();
  }
  const int64_t fy = ceil_div<int64_t>(nf, DENSE_FR);
  if (fy > 65535) return MFAC_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(N, DENSE_THREADS), (unsigned)fy, (unsigned)B);
  mdct_dense_kernel<<<grid, DENSE_THREADS, smem, stream>>>(x, X, ts.dense.wc, io, T, nf, N, hop);
  count_launch();
  return launch_status();
}

int mdct_inverse(const float* X, float* y, StridedIO io, int64_t B, int64_t nf, int N, int hop, cudaStream_t stream) {
  if (!X || !y) return MFAC_ERR_NULL;
  if (B <= 0 || nf <= 0 || N <= 0 || hop <= 0) return MFAC_ERR_BAD_SHAPE;
  if (B > 65535) return MFAC_ERR_UNSUPPORTED;
  const int64_t L = (nf - 1) * hop + 2 * (int64_t)N;
  TableSet ts;
  // output samples per CTA: 32 hops (bounded), frames touching them: spc/hop + 2N/hop + 2
  int64_t spc = 32LL * hop;
  if (spc > 16384) spc = 16384;
  if (spc < 1024) spc = 1024;
  const int max_frames = (int)((spc - 1) / hop + (2 * FFT_N - 1) / hop + 2);
  const size_t smem = TABLE_BYTES + SCRATCH_BYTES + (size_t)max_frames * FFT_N * 4;
  if (N == FFT_N && hop == FFT_H && io.out_elem_stride == 1 && io.in_elem_stride == FFT_N &&
      (io.in_clip_stride % 2) == 0 && (reinterpret_cast<uintptr_t>(X) & 7) == 0) {
    MFAC_OK(get_tables(N, true, &ts));
    static PerDeviceOnce configured2;
    if (configured2.need()) {
      MFAC_CUDA_OK(cudaFuncSetAttribute(imdct512h256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, I2_SMEM));
      configured2.done();
    }
    void* prof = profile_begin(MFAC_PROF_IMDCT, 4.0 * (double)B * ((double)L + (double)nf * N), stream);
    const int64_t nblocks = nf + 3;
    if (nblocks >= 16 && (io.out_clip_stride % 2) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0 && !g_imdct_gather) {
      // stream length S: whole waves of (SMs x 3 CTAs x 8 streams) with the 3-frame halo charged to every stream
      const double slots = (double)num_sms() * 3 * (F2_THREADS / 16);
      int best_S = 32;
      double best = -1.0;
      for (int S = 16; S <= 96; ++S) {
        const double streams = (double)B * (double)ceil_div<int64_t>(nblocks, S);
        const double waves = streams / slots;
        const double eff = (waves / std::ceil(waves)) * ((double)S / (S + 3));
        if (eff > best + 1e-9) { best = eff; best_S = S; }
      }
      if (const char* e = getenv("MFAC_IMDCT_S")) best_S = atoi(e) > 0 ? atoi(e) : best_S;   // tuning hook
      const int spc = (int)ceil_div<int64_t>(nblocks, best_S);
      best_S = (int)ceil_div<int64_t>(nblocks, spc);   // same stream count, balanced stream lengths
      const int64_t total = B * (int64_t)spc;
      imdct512h256_stream_kernel<<<(unsigned)ceil_div<int64_t>(total, F2_THREADS / 16), F2_THREADS, S2_SMEM, stream>>>(
          X, y, ts.fft, nf, io.in_clip_stride, io.out_clip_stride, best_S, spc, total);
    } else {
      dim3 grid((unsigned)ceil_div<int64_t>(nf + 3, I2_BLOCKS), (unsigned)B);
      imdct512h256_kernel<<<grid, F2_THREADS, I2_SMEM, stream>>>(X, y, ts.fft, nf, L, io.in_clip_stride, io.out_clip_stride);
    }
    profile_end(prof, stream);
    count_launch();
    return launch_status();
  }
  if (N == FFT_N && smem <= 160 * 1024) {  // tiny hops (< ~16) fall through to the dense path
    MFAC_OK(get_tables(N, true, &ts));
    static PerDeviceOnce configured;
    if (configured.need()) {
      MFAC_CUDA_OK(cudaFuncSetAttribute(imdct512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured.done();
    }
    dim3 grid((unsigned)ceil_div<int64_t>(L, spc), (unsigned)B);
    void* prof = profile_begin(MFAC_PROF_IMDCT, 4.0 * (double)B * ((double)L + (double)nf * N), stream);
    imdct512_kernel<<<grid, MDCT_WARPS * 32, smem, stream>>>(X, y, ts.fft, io, nf, L, hop, (int)spc, max_frames);
    profile_end(prof, stream);
    count_launch();
    return launch_status();
  }
  if (N > 4096) return MFAC_ERR_UNSUPPORTED;
  MFAC_OK(get_tables(N, false, &ts));
  dim3 grid((unsigned)ceil_div<int64_t>(L, 256), (unsigned)B);
  imdct_dense_kernel<<<grid, 256, 0, stream>>>(X, y, ts.dense.wct, io, nf, L, N, hop);
  count_launch();
  return launch_status();
}


// Fused tokenise + prologue launch (imf.cu).  MFAC_ERR_UNSUPPORTED when the geometry is outside the short-clip kernel: the
// caller then tokenises into a scratch buffer and runs the plain prologue.
int tokenize_prep_launch(const float* audio, int64_t T, int N, int hop, const PrepArgs& pa, const Dims& d, cudaStream_t stream) {
  if (!audio) return MFAC_ERR_NULL;
  if (N != FFT_N || hop != FFT_H || T <= 0) return MFAC_ERR_UNSUPPORTED;
  const int64_t nf = T < N ? 1 : (T - N) / hop + 1;
  if (nf > F2_FRAMES / 2 || nf * N != d.D || d.D != d.Dp) return MFAC_ERR_UNSUPPORTED;
  if (pa.e_in && (reinterpret_cast<uintptr_t>(pa.e_in) & 7)) return MFAC_ERR_UNSUPPORTED;
  static const bool off = getenv("MFAC_NO_FUSED_TOKENIZE") != nullptr;
  if (off) return MFAC_ERR_UNSUPPORTED;
  TableSet ts;
  MFAC_OK(get_tables(N, true, &ts));
  // clips per CTA: at most 16 frame slots (two trips of the 8 half-warps), so that three CTAs fit an SM's shared memory -- the
  // store stage (Philox + six output streams) wants the occupancy more than the staging wants long segments
  int cpc = F2_FRAMES / 2 / (int)nf;
  if (cpc < 1) cpc = 1;
  const int need = ((int)nf - 1) * FFT_H + 2 * FFT_N;
  const size_t smem_s = (size_t)2 * cpc * (need / 2 / 128 * 144) * 4 + (F2_THREADS / 16) * F2_EX * 8 + 256 * 8;
  static PerDeviceOnce configured;
  if (configured.need()) {
    MFAC_CUDA_OK(cudaFuncSetAttribute(tokenize_prep_short_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured.done();
  }
  void* prof = profile_begin(MFAC_PROF_MDCT, 4.0 * (double)pa.B * ((double)T + (double)nf * N), stream);
  tokenize_prep_short_kernel<<<(unsigned)ceil_div<int64_t>(pa.B, cpc), F2_THREADS, smem_s, stream>>>(audio, ts.fft, T, (int)nf, cpc, T,
                                                                                                    pa, d);
  profile_end(prof, stream);
  count_launch();
  return launch_status();
}
}  // namespace mfac

// ------------------------------------------------------------------ C ABI
using mfac::StridedIO;

extern "C" {

int64_t mfac_mdct_num_frames(int64_t T, int32_t N, int32_t hop) {
  if (T <= 0 || N <= 0 || hop <= 0) return MFAC_ERR_BAD_SHAPE;
  return T < N ? 1 : (T - N) / hop + 1;
}
int64_t mfac_imdct_length(int64_t nf, int32_t N, int32_t hop) {
  if (nf <= 0 || N <= 0 || hop <= 0) return MFAC_ERR_BAD_SHAPE;
  return (nf - 1) * hop + 2 * (int64_t)N;
}

int mfac_mdct_strided_f32(const float* x, int64_t x_clip_stride, int64_t x_elem_stride, float* X, int64_t X_clip_stride,
                          int64_t X_frame_stride, int64_t B, int64_t T, int32_t N, int32_t hop, void* stream) {
  StridedIO io{x_clip_stride, x_elem_stride, X_clip_stride, X_frame_stride};
  return mfac::mdct_forward(x, X, io, B, T, N, hop, (cudaStream_t)stream);
}
int mfac_imdct_strided_f32(const float* X, int64_t X_clip_stride, int64_t X_frame_stride, float* y, int64_t y_clip_stride,
                           int64_t y_elem_stride, int64_t B, int64_t nf, int32_t N, int32_t hop, void* stream) {
  StridedIO io{X_clip_stride, X_frame_stride, y_clip_stride, y_elem_stride};
  return mfac::mdct_inverse(X, y, io, B, nf, N, hop, (cudaStream_t)stream);
}
int mfac_mdct_f32(const float* x, float* X, int64_t B, int64_t T, int32_t N, int32_t hop, void* stream) {
  const int64_t nf = mfac_mdct_num_frames(T, N, hop);
  if (nf < 0) return (int)nf;
  return mfac_mdct_strided_f32(x, T, 1, X, nf * N, N, B, T, N, hop, stream);
}
int mfac_imdct_f32(const float* X, float* y, int64_t B, int64_t nf, int32_t N, int32_t hop, void* stream) {
  const int64_t L = mfac_imdct_length(nf, N, hop);
  if (L < 0) return (int)L;
  return mfac_imdct_strided_f32(X, nf * N, N, y, L, 1, B, nf, N, hop, stream);
}

}  // extern "C"
