// Data-parallel exchange step of the training path: the gradient all-reduce over NCCL (NVLink 5 / NVSwitch).
// The reference has no distributed code; BASELINE.json north_star asks for "bucketed gradient allreduce
// over NCCL on NVLink" at the step boundary of src/Trainer.py:65-76. This is the C-ABI side of it (SURVEY 8b:
// nv_dp_init / nv_dp_allreduce_bucket / nv_dp_destroy): the library owns ONE communicator per device, created from
// a 128-byte unique id the host exchanges however it likes (neurovit_b200/dp.py uses torch.distributed), and issues
// ncclAllReduce on the caller's stream — plain kernel launches from CUDA's point of view, so they can be captured
// into the training step's CUDA graph together with the backward kernels they overlap with.
//
// NCCL is resolved at run time with dlopen (the copy PyTorch ships), so the library has no link-time dependency and
// still loads on a box without NCCL (the entry points then return NV_ERR_NOT_INIT).
#include "nv_common.cuh"
#include <dlfcn.h>
#include <mutex>
#include <string.h>

namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI: stable since 2.0)
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_p;
typedef int ncclResult;  // 0 = ncclSuccess
// ncclConfig_t as NCCL 2.17 declared it; newer libraries accept it (they look at .version and default the fields
// added later). Per-communicator settings: unlike the NCCL_* environment variables, which the library reads once per
// process — PyTorch's own communicator has usually done that already — these cannot be pre-empted.
struct ncclConfig_v21700 {
  size_t size;
  unsigned int magic;
  unsigned int version;
  int blocking;
  int cgaClusterSize;
  int minCTAs;
  int maxCTAs;
  const char* netName;
};
constexpr int NCCL_UNDEF_INT = -2147483647 - 1;
enum { NCCL_SUM = 0, NCCL_AVG = 4, NCCL_FLOAT32 = 7, NCCL_BFLOAT16 = 9 };

struct NcclApi {
  void* lib = nullptr;
  ncclResult (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
  ncclResult (*CommInitRank)(ncclComm_p*, int, ncclUniqueId_t, int) = nullptr;
  ncclResult (*CommInitRankConfig)(ncclComm_p*, int, ncclUniqueId_t, int, void*) = nullptr;
  ncclResult (*CommDestroy)(ncclComm_p) = nullptr;
  ncclResult (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  ncclResult (*CommRegister)(ncclComm_p, void*, size_t, void**) = nullptr;
  ncclResult (*CommDeregister)(ncclComm_p, void*) = nullptr;
  const char* (*GetErrorString)(ncclResult) = nullptr;
  ncclResult (*GetVersion)(int*) = nullptr;
};

std::mutex g_mu;
NcclApi g_api;
struct DevComm { ncclComm_p comm = nullptr; int rank = 0, world = 1; void* reg[8] = {nullptr}; int nreg = 0; };
DevComm g_comm[64];

int cur_dev() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return -1;
  return d;
}

template <typename F>
bool sym(void* lib, const char* name, F& out) {
  out = reinterpret_cast<F>(dlsym(lib, name));
  return out != nullptr;
}

int nccl_fail(ncclResult r, const char* what) {
  nv_set_error("NCCL error %d (%s) in %s", (int)r, g_api.GetErrorString ? g_api.GetErrorString(r) : "?", what);
  return NV_ERR_CUDA;
}

}  // namespace

extern "C" {

// Load NCCL. path: a libnccl.so.2 to dlopen (NULL = "libnccl.so.2" through the normal search; when PyTorch is in
// the process its copy is already mapped and is the one returned).
int nv_dp_load(const char* path) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_api.lib != nullptr) return NV_OK;
  void* lib = dlopen(path != nullptr && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (lib == nullptr) {
    nv_set_error("nv_dp_load: cannot dlopen NCCL (%s)", dlerror());
    return NV_ERR_NOT_INIT;
  }
  NcclApi a;
  a.lib = lib;
  if (!(sym(lib, "ncclGetUniqueId", a.GetUniqueId) && sym(lib, "ncclCommInitRank", a.CommInitRank) &&
        sym(lib, "ncclCommDestroy", a.CommDestroy) && sym(lib, "ncclAllReduce", a.AllReduce) &&
        sym(lib, "ncclGetErrorString", a.GetErrorString) && sym(lib, "ncclGetVersion", a.GetVersion))) {
    nv_set_error("nv_dp_load: NCCL library lacks a required symbol");
    return NV_ERR_NOT_INIT;
  }
  sym(lib, "ncclCommInitRankConfig", a.CommInitRankConfig);   // optional (NCCL >= 2.14; maxCTAs since 2.17)
  sym(lib, "ncclCommRegister", a.CommRegister);      // optional (NCCL >= 2.19)
  sym(lib, "ncclCommDeregister", a.CommDeregister);
  g_api = a;
  return NV_OK;
}

// NCCL version code (e.g. 22809), 0 when not loaded
int nv_dp_nccl_version() {
  int v = 0;
  if (g_api.GetVersion == nullptr || g_api.GetVersion(&v) != 0) return 0;
  return v;
}

// rank 0: fill the 128-byte id that every rank passes to nv_dp_init
int nv_dp_unique_id(void* out128) {
  NV_REQUIRE(out128 != nullptr, "nv_dp_unique_id: null buffer");
  if (g_api.lib == nullptr) { nv_set_error("nv_dp: NCCL not loaded (nv_dp_load)"); return NV_ERR_NOT_INIT; }
  ncclUniqueId_t id;
  ncclResult r = g_api.GetUniqueId(&id);
  if (r != 0) return nccl_fail(r, "ncclGetUniqueId");
  memcpy(out128, id.internal, 128);
  return NV_OK;
}

// Create this process's communicator on the CURRENT device (collective: every rank calls it). One per device.
// max_ctas > 0: the communicator's collectives use at most that many CTAs (ncclConfig_t.maxCTAs) — the all-reduce
// runs under backward, where every SM it takes is taken from the persistent GEMMs.
int nv_dp_init(const void* uid128, int rank, int world, int max_ctas) {
  NV_REQUIRE(uid128 != nullptr && world >= 1 && rank >= 0 && rank < world, "nv_dp_init: bad rank %d / world %d", rank, world);
  if (g_api.lib == nullptr) { nv_set_error("nv_dp: NCCL not loaded (nv_dp_load)"); return NV_ERR_NOT_INIT; }
  const int d = cur_dev();
  NV_REQUIRE(d >= 0, "nv_dp_init: no current CUDA device");
  NV_REQUIRE(g_comm[d].comm == nullptr, "nv_dp_init: device %d already has a communicator (nv_dp_destroy first)", d);
  ncclUniqueId_t id;
  memcpy(id.internal, uid128, 128);
  ncclComm_p comm = nullptr;
  ncclResult r;
  int ver = 0;
  if (max_ctas > 0 && g_api.CommInitRankConfig != nullptr && g_api.GetVersion(&ver) == 0 && ver >= 21700) {
    ncclConfig_v21700 cfg;
    cfg.size = sizeof(cfg); cfg.magic = 0xcafebeef; cfg.version = 21700;
    cfg.blocking = NCCL_UNDEF_INT; cfg.cgaClusterSize = NCCL_UNDEF_INT; cfg.minCTAs = NCCL_UNDEF_INT;
    cfg.maxCTAs = max_ctas; cfg.netName = nullptr;
    r = g_api.CommInitRankConfig(&comm, world, id, rank, &cfg);
    if (r != 0) {   // argument validation fails identically (and immediately) on every rank: plain init instead
      comm = nullptr;
      r = g_api.CommInitRank(&comm, world, id, rank);
      if (r != 0) return nccl_fail(r, "ncclCommInitRank (after ncclCommInitRankConfig was rejected)");
    }
  } else {
    r = g_api.CommInitRank(&comm, world, id, rank);
    if (r != 0) return nccl_fail(r, "ncclCommInitRank");
  }
  g_comm[d].comm = comm; g_comm[d].rank = rank; g_comm[d].world = world; g_comm[d].nreg = 0;
  return NV_OK;
}

// Register a long-lived buffer (the flat gradient buffer) with the communicator: lets NCCL use zero-copy / NVLS
// paths on it. Optional; a no-op with an NCCL that lacks ncclCommRegister.
int nv_dp_register(void* buf, int64_t bytes) {
  const int d = cur_dev();
  NV_REQUIRE(d >= 0 && g_comm[d].comm != nullptr, "nv_dp_register: no communicator on this device");
  if (g_api.CommRegister == nullptr || g_comm[d].nreg >= 8) return NV_OK;
  void* h = nullptr;
  ncclResult r = g_api.CommRegister(g_comm[d].comm, buf, (size_t)bytes, &h);
  if (r != 0) return nccl_fail(r, "ncclCommRegister");
  g_comm[d].reg[g_comm[d].nreg++] = h;
  return NV_OK;
}

// In-place all-reduce of `count` elements at buf (dtype 0 = fp32, 1 = bf16; op 0 = sum, 1 = average over ranks),
// enqueued on `stream`. With world == 1 it is a no-op (and needs no communicator).
int nv_dp_allreduce_bucket(void* buf, int64_t count, int dtype, int op, void* stream) {
  NV_REQUIRE(buf != nullptr && count >= 0, "nv_dp_allreduce_bucket: bad buffer");
  NV_REQUIRE((dtype == 0 || dtype == 1) && (op == 0 || op == 1), "nv_dp_allreduce_bucket: bad dtype %d / op %d", dtype, op);
  const int d = cur_dev();
  NV_REQUIRE(d >= 0, "nv_dp_allreduce_bucket: no current CUDA device");
  if (g_comm[d].comm == nullptr) {
    nv_set_error("nv_dp_allreduce_bucket: no communicator on device %d (nv_dp_init)", d);
    return NV_ERR_NOT_INIT;
  }
  if (count == 0 || g_comm[d].world == 1) return NV_OK;
  ncclResult r = g_api.AllReduce(buf, buf, (size_t)count, dtype == 0 ? NCCL_FLOAT32 : NCCL_BFLOAT16,
                                 op == 0 ? NCCL_SUM : NCCL_AVG, g_comm[d].comm, static_cast<cudaStream_t>(stream));
  if (r != 0) return nccl_fail(r, "ncclAllReduce");
  return NV_OK;
}

int nv_dp_world(int* rank, int* world) {
  const int d = cur_dev();
  NV_REQUIRE(d >= 0 && g_comm[d].comm != nullptr, "nv_dp_world: no communicator on this device");
  if (rank) *rank = g_comm[d].rank;
  if (world) *world = g_comm[d].world;
  return NV_OK;
}

int nv_dp_destroy() {
  const int d = cur_dev();
  if (d < 0 || g_comm[d].comm == nullptr) return NV_OK;
  if (g_api.CommDeregister != nullptr)
    for (int i = 0; i < g_comm[d].nreg; ++i) g_api.CommDeregister(g_comm[d].comm, g_comm[d].reg[i]);
  ncclResult r = g_api.CommDestroy(g_comm[d].comm);
  g_comm[d] = DevComm();
  if (r != 0) return nccl_fail(r, "ncclCommDestroy");
  return NV_OK;
}

}  // extern "C"
