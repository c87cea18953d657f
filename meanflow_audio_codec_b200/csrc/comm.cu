// Data-parallel gradient all-reduce over NCCL (NVLink 5 / NVSwitch), one communicator per process.
// The reference has no distributed code; this is the single exchange step
// of the batch-sharded iMF training step.  NCCL is resolved with dlopen at first use so libmfac
// links against cudart only and picks up the NCCL the host process (PyTorch) already loaded.
#include <dlfcn.h>

#include <mutex>

#include "mfac_common.cuh"

namespace mfac {
namespace {
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm);
typedef int (*GroupFn)(void);

struct Api {
  void* handle = nullptr;
  GetUniqueIdFn get_id = nullptr;
  CommInitRankFn init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn destroy = nullptr;
  GroupFn group_start = nullptr, group_end = nullptr;
};
Api g_api;
NcclComm g_comm = nullptr;
std::mutex g_mu;

bool load_api() {
  if (g_api.all_reduce) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    void* h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (!h) continue;
    g_api.handle = h;
    g_api.get_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
    g_api.init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
    g_api.all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
    g_api.destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
    g_api.group_start = (GroupFn)dlsym(h, "ncclGroupStart");
    g_api.group_end = (GroupFn)dlsym(h, "ncclGroupEnd");
    if (g_api.get_id && g_api.init_rank && g_api.all_reduce && g_api.destroy && g_api.group_start && g_api.group_end) return true;
  }
  g_api = Api{};
  return false;
}
}  // namespace

// ---- internal entry points of the fused training step (imf.cu) ----
bool comm_ready() { return g_comm != nullptr; }
// sum all-reduce, in place, of nseg strided fp32 segments [buf + j * stride, + count) as ONE NCCL group (one fused launch)
int comm_allreduce_segments_f32(float* buf, int64_t count, int64_t stride, int nseg, cudaStream_t stream) {
  if (!g_comm) return MFAC_ERR_NCCL;
  if (count <= 0 || nseg <= 0) return MFAC_SUCCESS;
  if (nseg > 1 && g_api.group_start() != 0) return MFAC_ERR_NCCL;
  int rc = 0;
  for (int j = 0; j < nseg && rc == 0; ++j) {
    float* b = buf + (int64_t)j * stride;
    rc = g_api.all_reduce(b, b, (size_t)count, 7 /*ncclFloat32*/, 0 /*ncclSum*/, g_comm, stream);
  }
  if (nseg > 1 && g_api.group_end() != 0) return MFAC_ERR_NCCL;
  return rc == 0 ? MFAC_SUCCESS : MFAC_ERR_NCCL;
}
}  // namespace mfac

extern "C" {

int mfac_comm_unique_id(void* id_bytes_out) {
  using namespace mfac;
  if (!id_bytes_out) return MFAC_ERR_NULL;
  std::lock_guard<std::mutex> lock(g_mu);
  if (!load_api()) return MFAC_ERR_NCCL;
  return g_api.get_id(reinterpret_cast<NcclUniqueId*>(id_bytes_out)) == 0 ? MFAC_SUCCESS : MFAC_ERR_NCCL;
}

int mfac_comm_init(const void* id_bytes, int32_t rank, int32_t world) {
  using namespace mfac;
  if (!id_bytes) return MFAC_ERR_NULL;
  if (world <= 0 || rank < 0 || rank >= world) return MFAC_ERR_BAD_SHAPE;
  std::lock_guard<std::mutex> lock(g_mu);
  if (!load_api()) return MFAC_ERR_NCCL;
  if (g_comm) return MFAC_ERR_UNSUPPORTED;  // one communicator per process
  NcclUniqueId id = *reinterpret_cast<const NcclUniqueId*>(id_bytes);
  return g_api.init_rank(&g_comm, world, id, rank) == 0 ? MFAC_SUCCESS : MFAC_ERR_NCCL;
}

int mfac_comm_allreduce_sum_f32(float* buf, int64_t count, void* stream) {
  using namespace mfac;
  if (!buf) return MFAC_ERR_NULL;
  if (count <= 0) return MFAC_ERR_BAD_SHAPE;
  if (!g_comm) return MFAC_ERR_NCCL;
  // ncclFloat32 = 7, ncclSum = 0
  return g_api.all_reduce(buf, buf, (size_t)count, 7, 0, g_comm, (cudaStream_t)stream) == 0 ? MFAC_SUCCESS : MFAC_ERR_NCCL;
}

int mfac_comm_destroy(void) {
  using namespace mfac;
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_comm) {
    g_api.destroy(g_comm);
    g_comm = nullptr;
  }
  return MFAC_SUCCESS;
}

}  // extern "C"
