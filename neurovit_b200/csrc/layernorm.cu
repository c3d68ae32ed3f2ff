// LayerNorm forward/backward and the embedding finalisation (HBM-bound kernels).
// Reference call sites: nn.LayerNorm at src/models/vit_3d.py:18,37,93,95,108; cls/pos add at :116-118.
// Math: y=(x-mu)*rstd*g+b, biased variance, eps inside the sqrt.
//   dx = rstd*(gh - mean(gh) - xh*mean(gh*xh)),  gh = dy*g,  dg = sum_rows dy*xh,  db = sum_rows dy.
// One warp owns one row; a lane owns the float4 chunks {lane + 32*j}, so per-column partial sums for
// dg/db stay in registers across the grid-stride row loop and are reduced once per CTA.
#include "nv_common.cuh"

namespace {

constexpr int LN_WARPS = 8;

template <typename T> struct OutStore;
template <> struct OutStore<float> {
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct OutStore<bf16> {
  static __device__ __forceinline__ void st(bf16* p, float4 v) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
};

// Row addressing with an optional "grouped" map: logical row r -> (r / group) * stride + off + r % group.
// Used to walk the patch rows of x[B, n+1, D] while skipping each sample's cls row.
struct RowMap {
  int group, stride, off;
  __device__ __forceinline__ int64_t operator()(int r) const {
    return group > 0 ? (int64_t)(r / group) * stride + off + (r % group) : (int64_t)r;
  }
};

template <int NV>
__device__ __forceinline__ void row_stats(const float4 (&v)[NV], int nvec, int lane, int D, float eps,
                                          float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (lane + 32 * j < nvec) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (lane + 32 * j < nvec) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  rstd = rsqrtf(warp_sum(q) / (float)D + eps);
}

// y[ymap(r)] = LN(x[xmap(r)]) * gamma + beta (+ add[(r % add_mod) ...])
template <int NV, typename OutT>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const float* __restrict__ x, int64_t ld_x, RowMap xmap, const float* __restrict__ gamma,
              const float* __restrict__ beta, const float* __restrict__ add, int64_t ld_add, int add_mod,
              int add_off, OutT* __restrict__ y, int64_t ld_y, RowMap ymap, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, int M, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  for (int r = blockIdx.x * LN_WARPS + warp; r < M; r += gridDim.x * LN_WARPS) {
    const float* xr = x + xmap(r) * ld_x;
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (lane + 32 * j < nvec) v[j] = *reinterpret_cast<const float4*>(xr + 4 * (lane + 32 * j));
    float mean, rstd;
    row_stats<NV>(v, nvec, lane, D, eps, mean, rstd);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
    OutT* yr = y + ymap(r) * ld_y;
    const float* ar = add ? add + (int64_t)(r % add_mod + add_off) * ld_add : nullptr;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      if (c < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c);
        const float4 b = *reinterpret_cast<const float4*>(beta + 4 * c);
        float4 o;
        o.x = (v[j].x - mean) * rstd * g.x + b.x;
        o.y = (v[j].y - mean) * rstd * g.y + b.y;
        o.z = (v[j].z - mean) * rstd * g.z + b.z;
        o.w = (v[j].w - mean) * rstd * g.w + b.w;
        if (ar) {
          const float4 a = *reinterpret_cast<const float4*>(ar + 4 * c);
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        OutStore<OutT>::st(yr + 4 * c, o);
      }
    }
  }
}

// dx[dxmap(r)] = LNbwd(dy[dymap(r)]; x[xmap(r)]) (+ dres[r]);   dgamma/dbeta (+= via atomics);
// optional colsum_out[D] += sum_rows dx (bias gradient of the linear that produced x's residual branch)
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const float* __restrict__ dy, int64_t ld_dy, RowMap dymap, const float* __restrict__ x,
              int64_t ld_x, RowMap xmap, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const float* __restrict__ gamma, const float* __restrict__ dres, int64_t ld_dres,
              float* __restrict__ dx, int64_t ld_dx, RowMap dxmap, bf16* __restrict__ dx_bf16,
              int64_t ld_dxb, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ colsum_out, int M, int D) {
  extern __shared__ float red[];  // [LN_WARPS][D]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  float4 acc_g[NV], acc_b[NV], acc_c[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    acc_g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_b[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_c[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 g[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (lane + 32 * j < nvec) g[j] = *reinterpret_cast<const float4*>(gamma + 4 * (lane + 32 * j));

  for (int r = blockIdx.x * LN_WARPS + warp; r < M; r += gridDim.x * LN_WARPS) {
    const float* xr = x + xmap(r) * ld_x;
    const float* dyr = dy + dymap(r) * ld_dy;
    const float mean = mean_in[r], rstd = rstd_in[r];
    float4 xh[NV], gh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      if (c < nvec) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + 4 * c);
        const float4 d = *reinterpret_cast<const float4*>(dyr + 4 * c);
        xh[j].x = (xv.x - mean) * rstd; xh[j].y = (xv.y - mean) * rstd;
        xh[j].z = (xv.z - mean) * rstd; xh[j].w = (xv.w - mean) * rstd;
        acc_g[j].x += d.x * xh[j].x; acc_g[j].y += d.y * xh[j].y;
        acc_g[j].z += d.z * xh[j].z; acc_g[j].w += d.w * xh[j].w;
        acc_b[j].x += d.x; acc_b[j].y += d.y; acc_b[j].z += d.z; acc_b[j].w += d.w;
        gh[j].x = d.x * g[j].x; gh[j].y = d.y * g[j].y; gh[j].z = d.z * g[j].z; gh[j].w = d.w * g[j].w;
        s1 += (gh[j].x + gh[j].y) + (gh[j].z + gh[j].w);
        s2 += (gh[j].x * xh[j].x + gh[j].y * xh[j].y) + (gh[j].z * xh[j].z + gh[j].w * xh[j].w);
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
    if (dx != nullptr || dx_bf16 != nullptr) {
      const int64_t orow = dxmap(r);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = lane + 32 * j;
        if (c < nvec) {
          float4 o;
          o.x = rstd * (gh[j].x - s1 - xh[j].x * s2);
          o.y = rstd * (gh[j].y - s1 - xh[j].y * s2);
          o.z = rstd * (gh[j].z - s1 - xh[j].z * s2);
          o.w = rstd * (gh[j].w - s1 - xh[j].w * s2);
          if (dres) {
            const float4 a = *reinterpret_cast<const float4*>(dres + (int64_t)r * ld_dres + 4 * c);
            o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
          }
          acc_c[j].x += o.x; acc_c[j].y += o.y; acc_c[j].z += o.z; acc_c[j].w += o.w;
          if (dx) *reinterpret_cast<float4*>(dx + orow * ld_dx + 4 * c) = o;
          if (dx_bf16) OutStore<bf16>::st(dx_bf16 + orow * ld_dxb + 4 * c, o);
        }
      }
    }
  }
  // CTA-level reduction of the per-warp column partials, then one atomic per column per CTA
  for (int pass = 0; pass < 3; ++pass) {
    float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : colsum_out);
    if (dst == nullptr) continue;  // uniform across the CTA
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      if (c < nvec) {
        const float4 a = pass == 0 ? acc_g[j] : (pass == 1 ? acc_b[j] : acc_c[j]);
        *reinterpret_cast<float4*>(red + warp * D + 4 * c) = a;
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += LN_WARPS * 32) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) s += red[w * D + c];
      atomicAdd(dst + c, s);
    }
  }
}

// cls row of the embedding: x[b, 0, :] = cls + pos[0]   (vit_3d.py:116-118)
__global__ void cls_row_kernel(const float* __restrict__ cls, const float* __restrict__ pos,
                               float* __restrict__ x, int64_t batch_stride, int B, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, c = i % D;
  x[(int64_t)b * batch_stride + c] = cls[c] + pos[c];
}

int ln_grid(int M) {
  const int want = (M + LN_WARPS - 1) / LN_WARPS;
  const int cap = nv_num_sms() * 4;
  return want < cap ? (want > 0 ? want : 1) : cap;
}

}  // namespace

#define NV_LN_DISPATCH(D, ...)                                                            \
  do {                                                                                    \
    if ((D) <= 128) { constexpr int NV = 1; __VA_ARGS__; }                                \
    else if ((D) <= 512) { constexpr int NV = 4; __VA_ARGS__; }                           \
    else if ((D) <= 1024) { constexpr int NV = 8; __VA_ARGS__; }                          \
    else { constexpr int NV = 16; __VA_ARGS__; }                                          \
  } while (0)

int nv_ln_fwd_launch(const float* x, int64_t ld_x, int xg, int xs, int xo, const float* gamma,
                     const float* beta, const float* add, int64_t ld_add, int add_mod, int add_off,
                     void* y, int y_is_bf16, int64_t ld_y, int yg, int ys, int yo, float* mean, float* rstd,
                     int M, int D, float eps, cudaStream_t stream) {
  NV_REQUIRE(M >= 0 && D > 0 && D % 4 == 0 && D <= 2048, "layernorm: D=%d must be a multiple of 4 and <= 2048", D);
  if (M == 0) return NV_OK;
  NV_REQUIRE(ld_x % 4 == 0 && ld_y % 4 == 0, "layernorm: row strides must be multiples of 4 elements");
  RowMap xm{xg, xs, xo}, ym{yg, ys, yo};
  if (add_mod <= 0) add_mod = 1;
  const int grid = ln_grid(M);
  if (y_is_bf16) {
    NV_LN_DISPATCH(D, ln_fwd_kernel<NV, bf16><<<grid, LN_WARPS * 32, 0, stream>>>(
                          x, ld_x, xm, gamma, beta, add, ld_add, add_mod, add_off, (bf16*)y, ld_y, ym, mean,
                          rstd, M, D, eps));
  } else {
    NV_LN_DISPATCH(D, ln_fwd_kernel<NV, float><<<grid, LN_WARPS * 32, 0, stream>>>(
                          x, ld_x, xm, gamma, beta, add, ld_add, add_mod, add_off, (float*)y, ld_y, ym, mean,
                          rstd, M, D, eps));
  }
  NV_LAUNCH_CHECK("ln_fwd_kernel");
  return NV_OK;
}

int nv_ln_bwd_launch(const float* dy, int64_t ld_dy, int dyg, int dys, int dyo, const float* x, int64_t ld_x,
                     int xg, int xs, int xo, const float* mean, const float* rstd, const float* gamma,
                     const float* dres, int64_t ld_dres, float* dx, int64_t ld_dx, int dxg, int dxs, int dxo,
                     bf16* dx_bf16, int64_t ld_dxb, float* dgamma, float* dbeta, float* colsum, int M, int D,
                     cudaStream_t stream) {
  NV_REQUIRE(M >= 0 && D > 0 && D % 4 == 0 && D <= 2048, "layernorm bwd: D=%d must be a multiple of 4 and <= 2048", D);
  if (M == 0) return NV_OK;
  RowMap dym{dyg, dys, dyo}, xm{xg, xs, xo}, dxm{dxg, dxs, dxo};
  int grid = (M + LN_WARPS - 1) / LN_WARPS;
  const int cap = nv_num_sms() * 2;
  if (grid > cap) grid = cap;
  const size_t smem = (size_t)LN_WARPS * D * sizeof(float);
  if (smem > 48 * 1024) {
    NV_LN_DISPATCH(D, NV_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)smem)));
  }
  NV_LN_DISPATCH(D, ln_bwd_kernel<NV><<<grid, LN_WARPS * 32, smem, stream>>>(
                        dy, ld_dy, dym, x, ld_x, xm, mean, rstd, gamma, dres, ld_dres, dx, ld_dx, dxm, dx_bf16,
                        ld_dxb, dgamma, dbeta, colsum, M, D));
  NV_LAUNCH_CHECK("ln_bwd_kernel");
  return NV_OK;
}

int nv_cls_row_launch(const float* cls, const float* pos, float* x, int64_t batch_stride, int B, int D,
                      cudaStream_t stream) {
  if (B <= 0) return NV_OK;
  const int n = B * D;
  cls_row_kernel<<<(n + 255) / 256, 256, 0, stream>>>(cls, pos, x, batch_stride, B, D);
  NV_LAUNCH_CHECK("cls_row_kernel");
  return NV_OK;
}
