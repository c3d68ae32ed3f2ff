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
  const size_t smem = (size_t)2 * g.P * 4;
  NV_REQUIRE(smem <= 48 * 1024, "patch_embed bwd: patch_dim %d too large", g.P);
  patch_ln_param_grad_kernel<<<blocks, 256, smem, stream>>>(video, g, dP, ld_dp, mean, rstd, dgamma, dbeta,
                                                            rows_per_block);
  NV_LAUNCH_CHECK("patch_ln_param_grad_kernel");
  return NV_OK;
}
