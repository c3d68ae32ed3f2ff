// 3D patch gather + LayerNorm(patch_dim): the front of ViT.to_patch_embedding
// (reference src/models/vit_3d.py:91-93; layout adapter src/models/NeuroEncoder.py:197-204).
//
// Rearrange('b c (f pf) (h p1) (w p2) -> b (f h w) (p1 p2 pf c)'):
//   token   t = (fi * (H/p1) + hi) * (W/p2) + wi
//   feature j = ((p1i * p2 + p2i) * pf + pfi) * C + c
//   source  video[b, c, fi*pf + pfi, hi*p1 + p1i, wi*p2 + p2i]            
// The kernel takes the five element strides of the *view* it is given, so the non-contiguous
// [B,1,D,H,W] view that ViT3DEncoder builds from a [B,H,W,D] tensor is gathered in place (for that
// view consecutive j are consecutive addresses: the box lands in weight-K order with no permutation).
// One warp owns one token: gather into shared memory once, two-pass statistics, coalesced store.
#include "nv_common.cuh"
#include <stdlib.h>

namespace {

struct PatchGeom {
  int B, C, F, H, W;       // view shape
  int pf, p1, p2;          // patch sizes along F, H, W
  int nf, nh, nw;          // patch grid
  int P;                   // patch_dim = C*pf*p1*p2
  int64_t sb, sc, sf, sh, sw;  // element strides of the view
};

__device__ __forceinline__ int64_t patch_src_offset(const PatchGeom& g, int b, int fi, int hi, int wi, int j) {
  const int c = j % g.C;
  int t = j / g.C;
  const int pfi = t % g.pf; t /= g.pf;
  const int p2i = t % g.p2;
  const int p1i = t / g.p2;
  return (int64_t)b * g.sb + (int64_t)c * g.sc + (int64_t)(fi * g.pf + pfi) * g.sf +
         (int64_t)(hi * g.p1 + p1i) * g.sh + (int64_t)(wi * g.p2 + p2i) * g.sw;
}

template <typename OutT> __device__ __forceinline__ void store_out(OutT* p, float v);
template <> __device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_out<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }

// out[r, 0:P] = LN(patch r) * gamma + beta ; out[r, P:ld_out] = 0 ; raw (optional) = un-normalised patch
template <typename OutT>
__global__ void patch_gather_ln_kernel(const float* __restrict__ video, PatchGeom g,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       OutT* __restrict__ out, int64_t ld_out, float* __restrict__ raw,
                                       float* __restrict__ mean_out, float* __restrict__ rstd_out, float eps) {
  extern __shared__ float sm[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* buf = sm + (size_t)warp * g.P;
  const int n_tok = g.nf * g.nh * g.nw;
  const int rows = g.B * n_tok;
  for (int r = blockIdx.x * warps + warp; r < rows; r += gridDim.x * warps) {
    const int b = r / n_tok;
    int t = r % n_tok;
    const int wi = t % g.nw; t /= g.nw;
    const int hi = t % g.nh;
    const int fi = t / g.nh;
    float s = 0.f;
    for (int j = lane; j < g.P; j += 32) {
      const float v = video[patch_src_offset(g, b, fi, hi, wi, j)];
      buf[j] = v;
      s += v;
    }
    const float mean = warp_sum(s) / (float)g.P;
    float q = 0.f;
    for (int j = lane; j < g.P; j += 32) { const float d = buf[j] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)g.P + eps);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
    OutT* o = out ? out + (int64_t)r * ld_out : nullptr;
    for (int j = lane; j < (int)ld_out; j += 32) {
      if (j < g.P) {
        if (raw) raw[(int64_t)r * g.P + j] = buf[j];
        if (o) store_out<OutT>(o + j, (buf[j] - mean) * rstd * gamma[j] + beta[j]);
      } else if (o) {
        store_out<OutT>(o + j, 0.f);
      }
    }
    __syncwarp();
  }
}

// dgamma[j] += sum_r dP[r,j] * xhat[r,j],  dbeta[j] += sum_r dP[r,j]; xhat re-gathered from the volume
// (no patch tensor is kept for the backward pass).
__global__ void patch_ln_param_grad_kernel(const float* __restrict__ video, PatchGeom g,
                                           const float* __restrict__ dP, int64_t ld_dp,
                                           const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                           float* __restrict__ dgamma, float* __restrict__ dbeta, int rows_per_block) {
  extern __shared__ float sm[];  // [2][P]
  float* ag = sm;
  float* ab = sm + g.P;
  for (int j = threadIdx.x; j < g.P; j += blockDim.x) { ag[j] = 0.f; ab[j] = 0.f; }
  const int n_tok = g.nf * g.nh * g.nw;
  const int rows = g.B * n_tok;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  for (int r = r0; r < r1; ++r) {
    const int b = r / n_tok;
    int t = r % n_tok;
    const int wi = t % g.nw; t /= g.nw;
    const int hi = t % g.nh;
    const int fi = t / g.nh;
    const float mean = mean_in[r], rstd = rstd_in[r];
    for (int j = threadIdx.x; j < g.P; j += blockDim.x) {  // slot j is private to this thread
      const float xh = (video[patch_src_offset(g, b, fi, hi, wi, j)] - mean) * rstd;
      const float d = dP[(int64_t)r * ld_dp + j];
      ag[j] += d * xh;
      ab[j] += d;
    }
  }
  for (int j = threadIdx.x; j < g.P; j += blockDim.x) {
    atomicAdd(dgamma + j, ag[j]);
    atomicAdd(dbeta + j, ab[j]);
  }
}

// ---- TMA variants: one 5-D box {pf, p2, p1, 1, 1} per token ---------------------------
// For the view ViT3DEncoder builds ([B,1,D,H,W] over a contiguous [B,H,W,D] tensor: sf = 1 < sw < sh, C = 1) the
// box lands in shared memory in exactly the Rearrange feature order j = (p1i * p2 + p2i) * pf + pfi, so the
// gather is a single cp.async.bulk.tensor.5d per token and no per-element index arithmetic. A warp owns a
// token and double-buffers: the box of its next token is in flight while it normalises the current one.
// Other layouts (contiguous [B,C,F,H,W], C > 1, patch rows that are not 16-byte multiples such as patch 9) keep
// the strided-load kernels above — same results bit for bit.
constexpr int PG_WARPS = 8;

template <typename OutT>
__global__ void __launch_bounds__(PG_WARPS * 32)
patch_gather_ln_tma_kernel(const __grid_constant__ CUtensorMap tmap, PatchGeom g, const float* __restrict__ gamma,
                           const float* __restrict__ beta, OutT* __restrict__ out, int64_t ld_out,
                           float* __restrict__ raw, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                           float eps) {
  extern __shared__ __align__(128) uint8_t pg_smem[];
  __shared__ __align__(8) uint64_t bars[PG_WARPS][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t box_bytes = (uint32_t)g.P * 4u;
  const uint32_t buf_bytes = (box_bytes + 127u) & ~127u;
  float* buf0 = reinterpret_cast<float*>(pg_smem + (size_t)warp * 2 * buf_bytes);
  const int n_tok = g.nf * g.nh * g.nw;
  const int rows = g.B * n_tok;
  const int stride = gridDim.x * PG_WARPS;
  if (lane == 0) {
    mbar_init(&bars[warp][0], 1);
    mbar_init(&bars[warp][1], 1);
    fence_mbar_init();
  }
  __syncwarp();
  auto issue = [&](int r, int slot) {  // lane 0 only
    const int b = r / n_tok;
    int t = r % n_tok;
    const int wi = t % g.nw; t /= g.nw;
    const int hi = t % g.nh;
    const int fi = t / g.nh;
    mbar_arrive_expect_tx(&bars[warp][slot], box_bytes);
    tma_load_5d(reinterpret_cast<uint8_t*>(buf0) + slot * buf_bytes, &tmap, &bars[warp][slot], fi * g.pf, wi * g.p2,
                hi * g.p1, 0, b);
  };
  int r = blockIdx.x * PG_WARPS + warp;
  if (r < rows && lane == 0) issue(r, 0);
  for (int it = 0; r < rows; r += stride, ++it) {
    const int slot = it & 1;
    if (r + stride < rows && lane == 0) issue(r + stride, slot ^ 1);  // that buffer was released by the __syncwarp below
    mbar_wait(&bars[warp][slot], (it >> 1) & 1);
    const float* buf = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(buf0) + slot * buf_bytes);
    float s = 0.f;
    for (int j = lane; j < g.P; j += 32) s += buf[j];
    const float mean = warp_sum(s) / (float)g.P;
    float q = 0.f;
    for (int j = lane; j < g.P; j += 32) { const float d = buf[j] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)g.P + eps);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
    OutT* o = out ? out + (int64_t)r * ld_out : nullptr;
    for (int j = lane; j < (int)ld_out; j += 32) {
      if (j < g.P) {
        if (raw) raw[(int64_t)r * g.P + j] = buf[j];
        if (o) store_out<OutT>(o + j, (buf[j] - mean) * rstd * gamma[j] + beta[j]);
      } else if (o) {
        store_out<OutT>(o + j, 0.f);
      }
    }
    __syncwarp();  // every lane is done reading this buffer before lane 0 re-arms it two iterations later
  }
}

// dgamma / dbeta of the patch LayerNorm with the patches re-gathered by TMA: a CTA walks its rows with the
// volume box and the dP row (1-D bulk copy) double-buffered in shared memory; thread t owns slots t, t + 256, ..
constexpr int PGB_THREADS = 256;
constexpr int PGB_MAX_SLOTS = 8;  // patch_dim <= 2048

__global__ void __launch_bounds__(PGB_THREADS)
patch_ln_param_grad_tma_kernel(const __grid_constant__ CUtensorMap tmap, PatchGeom g, const float* __restrict__ dP,
                               int64_t ld_dp, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                               float* __restrict__ dgamma, float* __restrict__ dbeta, int rows_per_block) {
  extern __shared__ __align__(128) uint8_t pg_smem[];
  __shared__ __align__(8) uint64_t bars[2];
  const uint32_t box_bytes = (uint32_t)g.P * 4u;
  const uint32_t buf_bytes = (box_bytes + 127u) & ~127u;
  const int n_tok = g.nf * g.nh * g.nw;
  const int rows = g.B * n_tok;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  auto vbuf = [&](int slot) { return reinterpret_cast<float*>(pg_smem + (size_t)slot * 2 * buf_bytes); };
  auto dbuf = [&](int slot) { return reinterpret_cast<float*>(pg_smem + (size_t)slot * 2 * buf_bytes + buf_bytes); };
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int r, int slot) {  // thread 0 only
    const int b = r / n_tok;
    int t = r % n_tok;
    const int wi = t % g.nw; t /= g.nw;
    const int hi = t % g.nh;
    const int fi = t / g.nh;
    mbar_arrive_expect_tx(&bars[slot], 2 * box_bytes);
    tma_load_5d(vbuf(slot), &tmap, &bars[slot], fi * g.pf, wi * g.p2, hi * g.p1, 0, b);
    bulk_load_1d(dbuf(slot), dP + (int64_t)r * ld_dp, box_bytes, &bars[slot]);
  };
  float ag[PGB_MAX_SLOTS], ab[PGB_MAX_SLOTS];
#pragma unroll
  for (int k = 0; k < PGB_MAX_SLOTS; ++k) { ag[k] = 0.f; ab[k] = 0.f; }
  if (threadIdx.x == 0 && r0 < r1) issue(r0, 0);
  for (int r = r0, it = 0; r < r1; ++r, ++it) {
    const int slot = it & 1;
    if (threadIdx.x == 0 && r + 1 < r1) issue(r + 1, slot ^ 1);  // released by the __syncthreads of the previous iteration
    const float mean = __ldg(mean_in + r), rstd = __ldg(rstd_in + r);
    mbar_wait(&bars[slot], (it >> 1) & 1);
    const float* v = vbuf(slot);
    const float* d = dbuf(slot);
#pragma unroll
    for (int k = 0; k < PGB_MAX_SLOTS; ++k) {
      const int j = threadIdx.x + k * PGB_THREADS;
      if (j < g.P) {
        const float dd = d[j];
        ag[k] = fmaf(dd, (v[j] - mean) * rstd, ag[k]);
        ab[k] += dd;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < PGB_MAX_SLOTS; ++k) {
    const int j = threadIdx.x + k * PGB_THREADS;
    if (j < g.P) {
      atomicAdd(dgamma + j, ag[k]);
      atomicAdd(dbeta + j, ab[k]);
    }
  }
}

// The 5-D tensor map over the view, innermost-first (F, W, H, C, B); usable when the box order equals the
// Rearrange feature order and TMA's alignment rules hold.
bool make_patch_tmap(CUtensorMap* m, const float* video, const PatchGeom& g) {
  if (g.C != 1 || g.sf != 1 || !(g.sw < g.sh)) return false;
  if ((g.pf * 4) % 16 != 0 || g.pf > 256 || g.p1 > 256 || g.p2 > 256) return false;
  if ((g.sw * 4) % 16 != 0 || (g.sh * 4) % 16 != 0 || (g.sb * 4) % 16 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(video) & 15) != 0 || g.P > 2048) return false;
  const uint64_t dims[5] = {(uint64_t)g.F, (uint64_t)g.W, (uint64_t)g.H, 1, (uint64_t)g.B};
  const uint64_t sc_bytes = (uint64_t)g.sb * 4;  // C = 1: any 16-byte multiple is a valid stride for the unit dim
  const uint64_t strides[4] = {(uint64_t)g.sw * 4, (uint64_t)g.sh * 4, sc_bytes, (uint64_t)g.sb * 4};
  const uint32_t box[5] = {(uint32_t)g.pf, (uint32_t)g.p2, (uint32_t)g.p1, 1, 1};
  return nv_encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, video, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE) == NV_OK;
}

int fill_geom(PatchGeom& g, const int64_t* dims, const int64_t* strides, const int64_t* patch) {
  g.B = (int)dims[0]; g.C = (int)dims[1]; g.F = (int)dims[2]; g.H = (int)dims[3]; g.W = (int)dims[4];
  g.pf = (int)patch[0]; g.p1 = (int)patch[1]; g.p2 = (int)patch[2];
  NV_REQUIRE(g.B >= 0 && g.C > 0 && g.F > 0 && g.H > 0 && g.W > 0 && g.pf > 0 && g.p1 > 0 && g.p2 > 0,
             "patch_embed: non-positive dimension");
  NV_REQUIRE(g.F % g.pf == 0, "Frames must be divisible by frame patch size");
  NV_REQUIRE(g.H % g.p1 == 0 && g.W % g.p2 == 0, "Image dimensions must be divisible by the patch size.");
  g.nf = g.F / g.pf; g.nh = g.H / g.p1; g.nw = g.W / g.p2;
  g.P = g.C * g.pf * g.p1 * g.p2;
  g.sb = strides[0]; g.sc = strides[1]; g.sf = strides[2]; g.sh = strides[3]; g.sw = strides[4];
  return NV_OK;
}

}  // namespace

int nv_patch_gather_ln_launch(const float* video, const int64_t* dims, const int64_t* strides,
                              const int64_t* patch, const float* gamma, const float* beta, void* out,
                              int out_is_bf16, int64_t ld_out, float* raw, float* mean, float* rstd, float eps,
                              cudaStream_t stream) {
  PatchGeom g;
  int s = fill_geom(g, dims, strides, patch);
  if (s != NV_OK) return s;
  const int rows = g.B * g.nf * g.nh * g.nw;
  if (rows == 0) return NV_OK;
  NV_REQUIRE(out == nullptr || ld_out >= g.P, "patch_embed: ld_out %lld < patch_dim %d", (long long)ld_out, g.P);
  NV_REQUIRE((size_t)g.P * 4 <= 200 * 1024, "patch_embed: patch_dim %d too large for shared memory staging", g.P);
  CUtensorMap tmap;
  if (!getenv("NV_PATCH_NO_TMA") && make_patch_tmap(&tmap, video, g)) {
    const size_t buf_bytes = ((size_t)g.P * 4 + 127) & ~(size_t)127;
    const size_t smem_t = (size_t)PG_WARPS * 2 * buf_bytes;
    int grid_t = (rows + PG_WARPS - 1) / PG_WARPS;
    if (grid_t > nv_num_sms() * 4) grid_t = nv_num_sms() * 4;
    if (out_is_bf16) {
      if (smem_t > 48 * 1024)
        NV_CUDA(cudaFuncSetAttribute(patch_gather_ln_tma_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      patch_gather_ln_tma_kernel<bf16><<<grid_t, PG_WARPS * 32, smem_t, stream>>>(tmap, g, gamma, beta, (bf16*)out, ld_out,
                                                                                  raw, mean, rstd, eps);
    } else {
      if (smem_t > 48 * 1024)
        NV_CUDA(cudaFuncSetAttribute(patch_gather_ln_tma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      patch_gather_ln_tma_kernel<float><<<grid_t, PG_WARPS * 32, smem_t, stream>>>(tmap, g, gamma, beta, (float*)out, ld_out,
                                                                                   raw, mean, rstd, eps);
    }
    NV_LAUNCH_CHECK("patch_gather_ln_tma_kernel");
    return NV_OK;
  }
  // (a failed tensor-map encode above is not an error: the strided-load kernel handles the layout)
  int warps = 8;
  while (warps > 1 && (size_t)warps * g.P * 4 > 96 * 1024) warps >>= 1;
  const size_t smem = (size_t)warps * g.P * 4;
  int grid = (rows + warps - 1) / warps;
  if (grid > nv_num_sms() * 8) grid = nv_num_sms() * 8;
  if (out_is_bf16) {
    if (smem > 48 * 1024)
      NV_CUDA(cudaFuncSetAttribute(patch_gather_ln_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    patch_gather_ln_kernel<bf16><<<grid, warps * 32, smem, stream>>>(video, g, gamma, beta, (bf16*)out, ld_out, raw,
                                                                      mean, rstd, eps);
  } else {
    if (smem > 48 * 1024)
      NV_CUDA(cudaFuncSetAttribute(patch_gather_ln_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    patch_gather_ln_kernel<float><<<grid, warps * 32, smem, stream>>>(video, g, gamma, beta, (float*)out, ld_out, raw,
                                                                       mean, rstd, eps);
  }
  NV_LAUNCH_CHECK("patch_gather_ln_kernel");
  return NV_OK;
}

int nv_patch_ln_param_grad_launch(const float* video, const int64_t* dims, const int64_t* strides,
                                  const int64_t* patch, const float* dP, int64_t ld_dp, const float* mean,
                                  const float* rstd, float* dgamma, float* dbeta, cudaStream_t stream) {
  PatchGeom g;
  int s = fill_geom(g, dims, strides, patch);
  if (s != NV_OK) return s;
  const int rows = g.B * g.nf * g.nh * g.nw;
  if (rows == 0) return NV_OK;
  int blocks = nv_num_sms() * 2;
  int rows_per_block = (rows + blocks - 1) / blocks;
  if (rows_per_block < 4) rows_per_block = 4;
  blocks = (rows + rows_per_block - 1) / rows_per_block;
  CUtensorMap tmap;
  if (!getenv("NV_PATCH_NO_TMA") && (ld_dp * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(dP) & 15) == 0 &&
      (g.P * 4) % 16 == 0 && make_patch_tmap(&tmap, video, g)) {
    const size_t buf_bytes = ((size_t)g.P * 4 + 127) & ~(size_t)127;
    patch_ln_param_grad_tma_kernel<<<blocks, PGB_THREADS, 4 * buf_bytes, stream>>>(tmap, g, dP, ld_dp, mean, rstd, dgamma,
                                                                                   dbeta, rows_per_block);
    NV_LAUNCH_CHECK("patch_ln_param_grad_tma_kernel");
    return NV_OK;
  }

  const size_t smem = (size_t)2 * g.P * 4;
  NV_REQUIRE(smem <= 48 * 1024, "patch_embed bwd: patch_dim %d too large", g.P);
  patch_ln_param_grad_kernel<<<blocks, 256, smem, stream>>>(video, g, dP, ld_dp, mean, rstd, dgamma, dbeta,
                                                            rows_per_block);
  NV_LAUNCH_CHECK("patch_ln_param_grad_kernel");
  return NV_OK;
}

// ---- LayerNorm(patch_dim) folded into the Linear that follows it (vit_3d.py:93-94) ------------------------------
// Linear(LN(x)) = xhat (W o gamma)^T + (W beta + b), xhat = (x - mean) rstd: the gather kernel writes xhat (gamma = 1,
// beta = 0) and the patch GEMM runs on the folded weight. Backward then needs ONE weight-gradient GEMM,
// G = de^T xhat [D, P], from which every parameter gradient of the pair follows in closed form:
//     dW[k, j] = G[k, j] gamma[j] + cs[k] beta[j]        dgamma[j] = sum_k W[k, j] G[k, j]
//     db[k]    = cs[k] = sum_rows de[., k]               dbeta[j]  = sum_k cs[k] W[k, j]
// instead of a second [B n, D] x [D, P] GEMM (dP = de W) and a pass that gathers every patch of the volume again to
// reduce dP against xhat (25.6 + 53.4 us per cfgA step).
namespace {

constexpr int FOLD_THREADS = 128;
template <typename T> struct OutStore1;
template <> struct OutStore1<float> { static __device__ __forceinline__ void st(float* p, float v) { *p = v; } };
template <> struct OutStore1<bf16> { static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16(v); } };
// one CTA per output row k: Wf[k, :] = W[k, :] o gamma (columns P .. Pp-1 zero), bias_f[k] = b[k] + W[k, :] . beta
template <typename OutT>
__global__ void __launch_bounds__(FOLD_THREADS)
ln_fold_kernel(const float* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ b, OutT* __restrict__ Wf, int64_t ld_wf, float* __restrict__ bias_f, int P,
               int Pp) {
  __shared__ float red[FOLD_THREADS / 32];
  const int k = blockIdx.x;
  const float* wr = W + (int64_t)k * P;
  OutT* fr = Wf + (int64_t)k * ld_wf;
  float acc = 0.f;
  for (int j = threadIdx.x; j < Pp; j += FOLD_THREADS) {
    float v = 0.f;
    if (j < P) {
      const float w = wr[j];
      v = w * gamma[j];
      acc = fmaf(w, beta[j], acc);
    }
    OutStore1<OutT>::st(fr + j, v);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < FOLD_THREADS / 32; ++i) s += red[i];
    bias_f[k] = (b ? b[k] : 0.f) + s;
  }
}

constexpr int FG_TX = 32, FG_TY = 8;
// grid (ceil(P / 32), slices of the D rows); block (32 columns, 8 row lanes). Every output but dW is accumulated (+=).
__global__ void __launch_bounds__(FG_TX * FG_TY)
ln_fold_grads_kernel(const float* __restrict__ G, int64_t ld_g, const float* __restrict__ W,
                     const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ cs,
                     float* __restrict__ dW, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ db, int D, int P, int rows_per_slice) {
  __shared__ float red_g[FG_TY][FG_TX], red_b[FG_TY][FG_TX];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int j = blockIdx.x * FG_TX + tx;
  const int k0 = blockIdx.y * rows_per_slice, k1 = min(D, k0 + rows_per_slice);
  const bool ok = j < P;
  const float gj = ok ? gamma[j] : 0.f, bj = ok ? beta[j] : 0.f;
  float acc_g = 0.f, acc_b = 0.f;
  for (int k = k0 + ty; k < k1; k += FG_TY) {
    if (!ok) continue;
    const float g = G[(int64_t)k * ld_g + j], w = W[(int64_t)k * P + j], c = cs[k];
    dW[(int64_t)k * P + j] += fmaf(g, gj, c * bj);
    acc_g = fmaf(w, g, acc_g);
    acc_b = fmaf(c, w, acc_b);
  }
  red_g[ty][tx] = acc_g;
  red_b[ty][tx] = acc_b;
  __syncthreads();
  if (ty == 0 && ok) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < FG_TY; ++i) { sg += red_g[i][tx]; sb += red_b[i][tx]; }
    atomicAdd(dgamma + j, sg);
    atomicAdd(dbeta + j, sb);
  }
  if (db != nullptr && blockIdx.x == 0)   // the Linear's bias gradient is cs itself
    for (int k = k0 + ty * FG_TX + tx; k < k1; k += FG_TX * FG_TY) db[k] += cs[k];
}

}  // namespace

int nv_ln_fold_launch(const float* W, const float* gamma, const float* beta, const float* b, void* Wf, int wf_is_bf16,
                      int64_t ld_wf, float* bias_f, int D, int P, cudaStream_t stream) {
  NV_REQUIRE(D > 0 && P > 0 && ld_wf >= P, "ln_fold: bad sizes D=%d P=%d ld=%lld", D, P, (long long)ld_wf);
  NV_REQUIRE(W && gamma && beta && Wf && bias_f, "ln_fold: null pointer");
  const int Pp = (int)ld_wf;   // pad columns of the folded weight are written as zeros (K padded to a multiple of 8)
  if (wf_is_bf16) ln_fold_kernel<bf16><<<D, FOLD_THREADS, 0, stream>>>(W, gamma, beta, b, (bf16*)Wf, ld_wf, bias_f, P, Pp);
  else ln_fold_kernel<float><<<D, FOLD_THREADS, 0, stream>>>(W, gamma, beta, b, (float*)Wf, ld_wf, bias_f, P, Pp);
  NV_LAUNCH_CHECK("ln_fold_kernel");
  return NV_OK;
}

int nv_ln_fold_grads_launch(const float* G, int64_t ld_g, const float* W, const float* gamma, const float* beta,
                            const float* cs, float* dW, float* dgamma, float* dbeta, float* db, int D, int P,
                            cudaStream_t stream) {
  NV_REQUIRE(D > 0 && P > 0 && ld_g >= P, "ln_fold_grads: bad sizes D=%d P=%d ld=%lld", D, P, (long long)ld_g);
  NV_REQUIRE(G && W && gamma && beta && cs && dW && dgamma && dbeta, "ln_fold_grads: null pointer");
  const int slices = D >= 512 ? 8 : 1;
  const int rows_per_slice = (D + slices - 1) / slices;
  dim3 grid((P + FG_TX - 1) / FG_TX, slices), block(FG_TX, FG_TY);
  ln_fold_grads_kernel<<<grid, block, 0, stream>>>(G, ld_g, W, gamma, beta, cs, dW, dgamma, dbeta, db, D, P,
                                                   rows_per_slice);
  NV_LAUNCH_CHECK("ln_fold_grads_kernel");
  return NV_OK;
}
