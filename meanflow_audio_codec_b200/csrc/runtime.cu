// Process-wide runtime helpers of libmfac: device properties, TMA tensor-map encoding
// (driver entry point resolved at run time so the library links against cudart only),
// launch counters and the GEMM test hook.
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "gemm.cuh"

namespace mfac {

namespace {
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_simt{0};
std::atomic<int> g_pair{getenv("MFAC_NO_PAIR_GEMM") ? 0 : 1};
std::atomic<int> g_streamk{getenv("MFAC_NO_STREAM_K") ? 0 : 1};
std::atomic<int> g_num_sms[128] = {};   // per device

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;

void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess) {
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
}
}  // namespace

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional per-launch event timing -------------------------------------------------
namespace {
struct ProfRecord {
  int family;
  double work;
  cudaEvent_t e0, e1;
  const char* label;
  int M, N, K;
};
std::mutex g_prof_mu;
std::vector<ProfRecord> g_prof;
std::atomic<int> g_prof_on{0};
}  // namespace

bool profile_enabled() { return g_prof_on.load(std::memory_order_relaxed) != 0; }
void* profile_begin(int family, double work, cudaStream_t s, const char* label, int M, int N, int K) {
  if (!profile_enabled()) return nullptr;
  ProfRecord r{family, work, nullptr, nullptr, label, M, N, K};
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return nullptr;
  cudaEventRecord(r.e0, s);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof.push_back(r);
  return reinterpret_cast<void*>(g_prof.size());  // 1-based index
}
void profile_end(void* token, cudaStream_t s) {
  if (!token) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  cudaEventRecord(g_prof[reinterpret_cast<size_t>(token) - 1].e1, s);
}
// ---- phase marks (debug): events at the phase boundaries of one mfac_imf_loss_grad / train_step call
namespace {
std::atomic<int> g_phase_on{0};
std::mutex g_phase_mu;
std::vector<std::pair<int, cudaEvent_t>> g_phase;
}  // namespace
void phase_mark(int id, cudaStream_t s) {
  if (!g_phase_on.load(std::memory_order_relaxed)) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, s);
  std::lock_guard<std::mutex> lock(g_phase_mu);
  g_phase.push_back({id, e});
}

bool simt_gemm_enabled() { return g_simt.load(std::memory_order_relaxed) != 0; }
bool pair_gemm_enabled() { return g_pair.load(std::memory_order_relaxed) != 0; }
bool stream_k_enabled() { return g_streamk.load(std::memory_order_relaxed) != 0; }

int num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 127) dev = 0;
  int n = g_num_sms[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  g_num_sms[dev].store(n, std::memory_order_relaxed);
  return n;
}

namespace {
struct ForkSlot {
  int dev = -1;
  bool ok = false;
  ForkCtx ctx;
};
}  // namespace

ForkCtx* fork_ctx() {
  static thread_local std::vector<ForkSlot*> slots;   // one per device this thread has used (never freed: process lifetime)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  for (ForkSlot* sl : slots)
    if (sl->dev == dev) return sl->ok ? &sl->ctx : nullptr;
  ForkSlot* sl = new ForkSlot();
  sl->dev = dev;
  sl->ok = true;
  for (int i = 0; i < ForkCtx::kStreams && sl->ok; ++i)
    sl->ok = cudaStreamCreateWithFlags(&sl->ctx.side[i], cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < ForkCtx::kEvents && sl->ok; ++i)
    sl->ok = cudaEventCreateWithFlags(&sl->ctx.ev[i], cudaEventDisableTiming) == cudaSuccess;
  if (!sl->ok) cudaGetLastError();
  slots.push_back(sl);
  return sl->ok ? &sl->ctx : nullptr;
}

namespace {
std::atomic<int> g_conc_rows{getenv("MFAC_CONC_MAX_ROWS") ? atoi(getenv("MFAC_CONC_MAX_ROWS")) : 4096};
}
int concurrency_max_rows() { return g_conc_rows.load(std::memory_order_relaxed); }

int make_tmap_bf16_sw(CUtensorMap* out, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                      int box_outer, int swizzle_bytes) {
  std::call_once(g_encode_once, resolve_encode);
  if (!g_encode) return MFAC_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0) return MFAC_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MFAC_SUCCESS : MFAC_ERR_DRIVER;
}
int make_tmap_bf16(CUtensorMap* out, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                   int box_outer) {
  return make_tmap_bf16_sw(out, ptr, inner, outer, ld, box_inner, box_outer, 128);
}

}  // namespace mfac

extern "C" {

int mfac_version(void) { return 101; }   // 101: MfacImfConfig gained method / gamma / uniform_time / rows_r_equals_t

const char* mfac_status_string(int status) {
  switch (status) {
    case MFAC_SUCCESS: return "success";
    case MFAC_ERR_BAD_SHAPE: return "bad shape";
    case MFAC_ERR_UNSUPPORTED: return "unsupported configuration";
    case MFAC_ERR_WORKSPACE: return "workspace missing or too small";
    case MFAC_ERR_NULL: return "null pointer";
    case MFAC_ERR_DRIVER: return "CUDA driver entry point unavailable";
    case MFAC_ERR_NCCL: return "NCCL unavailable or failed";
    default: break;
  }
  if (status <= -1000) return cudaGetErrorString((cudaError_t)(-status - 1000));
  return "unknown status";
}

int mfac_debug_gemm_bf16(const void* A, const void* B, float* Cout, int64_t M, int64_t N, int64_t K, int32_t a_mn_major,
                         int32_t b_mn_major, int32_t block_n, void* stream) {
  using namespace mfac;
  if (!A || !B || !Cout) return MFAC_ERR_NULL;
  GemmOperandDesc a{A, a_mn_major ? M : K, a_mn_major != 0};
  GemmOperandDesc b{B, b_mn_major ? N : K, b_mn_major != 0};
  EpiStoreF32 epi{Cout, N};
  cudaStream_t s = (cudaStream_t)stream;
  if (!a_mn_major && !b_mn_major) return launch_gemm<false, false>(a, b, (int)M, (int)N, (int)K, epi, s, block_n);
  if (!a_mn_major && b_mn_major) return launch_gemm<false, true>(a, b, (int)M, (int)N, (int)K, epi, s, block_n);
  if (a_mn_major && !b_mn_major) return launch_gemm<true, false>(a, b, (int)M, (int)N, (int)K, epi, s, block_n);
  return launch_gemm<true, true>(a, b, (int)M, (int)N, (int)K, epi, s, block_n);
}

int mfac_debug_set_simt_gemm(int32_t on) {
  mfac::g_simt.store(on ? 1 : 0);
  return MFAC_SUCCESS;
}

int mfac_debug_set_stream_k(int32_t on) {
  mfac::g_streamk.store(on ? 1 : 0);
  return MFAC_SUCCESS;
}

int mfac_set_concurrency_max_rows(int32_t rows) {
  mfac::g_conc_rows.store(rows < 0 ? 0 : rows);
  return MFAC_SUCCESS;
}

int mfac_debug_set_pair_gemm(int32_t on) {
  mfac::g_pair.store(on ? 1 : 0);
  return MFAC_SUCCESS;
}

int mfac_profile_enable(int32_t on) {
  mfac::g_prof_on.store(on ? 1 : 0);
  return MFAC_SUCCESS;
}

int mfac_profile_collect(int64_t* launches, double* ms, double* work) {
  using namespace mfac;
  if (!launches || !ms || !work) return MFAC_ERR_NULL;
  MFAC_CUDA_OK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (int f = 0; f < MFAC_PROF_FAMILIES; ++f) { launches[f] = 0; ms[f] = 0.0; work[f] = 0.0; }
  FILE* csv = nullptr;
  if (const char* path = getenv("MFAC_PROFILE_CSV")) csv = fopen(path, "a");  // per-launch dump for profiles/
  for (auto& r : g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess && r.family >= 0 && r.family < MFAC_PROF_FAMILIES) {
      if (csv) fprintf(csv, "%d,%s,%d,%d,%d,%.3f,%.6g\n", r.family, r.label ? r.label : "", r.M, r.N, r.K, t * 1e3, r.work);
      launches[r.family] += 1;
      ms[r.family] += t;
      work[r.family] += r.work;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  if (csv) fclose(csv);
  g_prof.clear();
  return MFAC_SUCCESS;
}

int mfac_debug_phase_marks(int32_t on) {
  mfac::g_phase_on.store(on ? 1 : 0);
  return MFAC_SUCCESS;
}

// ids[i], ms_since_first[i] of the marks recorded so far (cleared afterwards); returns the number of marks
int mfac_debug_phase_collect(int32_t* ids, float* ms_since_first, int32_t cap) {
  using namespace mfac;
  if (!ids || !ms_since_first) return MFAC_ERR_NULL;
  if (cudaDeviceSynchronize() != cudaSuccess) return MFAC_ERR_DRIVER;
  std::lock_guard<std::mutex> lock(g_phase_mu);
  int n = 0;
  for (auto& pr : g_phase) {
    if (n < cap) {
      float t = 0.f;
      cudaEventElapsedTime(&t, g_phase[0].second, pr.second);
      ids[n] = pr.first;
      ms_since_first[n] = t;
      ++n;
    }
  }
  for (auto& pr : g_phase) cudaEventDestroy(pr.second);
  g_phase.clear();
  return n;
}

int mfac_debug_counters(int64_t* kernel_launches) {
  if (kernel_launches) *kernel_launches = mfac::g_launches.load();
  return MFAC_SUCCESS;
}

}  // extern "C"
