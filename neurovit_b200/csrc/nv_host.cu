// Host-side plumbing shared by all kernels: error text, device query, TMA descriptor encoding.
#include "nv_common.cuh"
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>
#include <mutex>

namespace {
thread_local char g_err[1024] = "";
std::mutex g_mu;
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
}  // namespace

void nv_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* nv_last_error_impl() { return g_err; }

int nv_check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return NV_OK;
  nv_set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return NV_ERR_CUDA;
}

namespace {
int g_sm_reserve[64] = {0};
}
// SMs the persistent kernels may fill: the device's SM count minus the SMs reserved for a concurrent communication
// kernel (nv_set_sm_reserve). A persistent grid that asked for every SM while NCCL's CTAs hold a few would leave
// some of its CTAs queued until others finish — a whole extra wave.
int nv_num_sms() {
  static int cached[64] = {0};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached[dev] = n;
  }
  const int avail = cached[dev] - g_sm_reserve[dev];
  return avail >= 2 ? avail : 2;
}

int nv_set_sm_reserve_impl(int n) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { nv_set_error("nv_set_sm_reserve: no current device"); return NV_ERR_CUDA; }
  NV_REQUIRE(n >= 0 && n <= 64, "nv_set_sm_reserve: %d out of range [0, 64]", n);
  g_sm_reserve[dev] = (n + 1) & ~1;   // even: the CTA-pair GEMM needs an even number of SMs
  return NV_OK;
}

bool nv_first_on_device(uint64_t* flags) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;  // unknown device: just redo the work
  const uint64_t bit = 1ull << dev;
  return (__atomic_fetch_or(flags, bit, __ATOMIC_ACQ_REL) & bit) == 0;
}

int nv_encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base,
                   const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                   CUtensorMapSwizzle swizzle) {
  if (g_encode == nullptr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_encode == nullptr) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
      if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
        nv_set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
        return NV_ERR_CUDA;
      }
      g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
  }
  cuuint64_t gdims[5], gstr[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = g_encode(map, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    nv_set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank=%d dims=[%llu,%llu,..] box=[%u,%u,..] "
                 "stride0=%llu base=%p",
                 (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                 box[0], rank > 1 ? box[1] : 0,
                 (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), base);
    return NV_ERR_CUDA;
  }
  return NV_OK;
}

// ---- dropout epoch counter -----------------------------------------------------------------------------
namespace {
std::mutex g_epoch_mu;
uint64_t* g_epoch[64] = {nullptr};
__global__ void epoch_add_kernel(uint64_t* e, uint64_t inc) { *e += inc; }
__global__ void counter_add_kernel(float* c, float inc) { *c += inc; }
}  // namespace

const uint64_t* nv_rng_epoch_dev() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (g_epoch[dev] == nullptr) {
    std::lock_guard<std::mutex> lk(g_epoch_mu);
    if (g_epoch[dev] == nullptr) {
      uint64_t* p = nullptr;
      if (cudaMalloc(&p, sizeof(uint64_t)) != cudaSuccess) return nullptr;
      cudaMemset(p, 0, sizeof(uint64_t));
      g_epoch[dev] = p;
    }
  }
  return g_epoch[dev];
}

int nv_rng_epoch_advance_launch(cudaStream_t stream) {
  uint64_t* e = const_cast<uint64_t*>(nv_rng_epoch_dev());
  NV_REQUIRE(e != nullptr, "rng epoch: allocation failed");
  epoch_add_kernel<<<1, 1, 0, stream>>>(e, 1);
  NV_LAUNCH_CHECK("epoch_add_kernel");
  return NV_OK;
}

int nv_rng_epoch_read(uint64_t* out, cudaStream_t stream) {
  NV_REQUIRE(out != nullptr, "rng epoch: null output");
  const uint64_t* e = nv_rng_epoch_dev();
  NV_REQUIRE(e != nullptr, "rng epoch: allocation failed");
  NV_CUDA(cudaMemcpyAsync(out, e, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  NV_CUDA(cudaStreamSynchronize(stream));
  return NV_OK;
}

int nv_counter_add_launch(float* counter, float inc, cudaStream_t stream) {
  NV_REQUIRE(counter != nullptr, "counter_add: null pointer");
  counter_add_kernel<<<1, 1, 0, stream>>>(counter, inc);
  NV_LAUNCH_CHECK("counter_add_kernel");
  return NV_OK;
}
