// fp32 verification path ("precision_mode = 1"): plain FMA GEMM with arbitrary strides and a two-level
// batch index, plus the materialised softmax forward/backward of the reference attention
// (src/models/vit_3d.py:53-58). tf32/bf16 tensor cores cannot meet the 1e-5 tolerance, so this path
// uses CUDA-core FMAs only. It exists to separate "algorithm wrong" from "bf16 rounding"; it is not
// the fast path. The same kernel serves the tiny GEMMs of the classification head (vit_3d.py:107-110).
#include "nv_common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtParams {
  int M, N, K, Z2;
  const float* A; int64_t sa_m, sa_k, sa_z1, sa_z2;
  const float* B; int64_t sb_n, sb_k, sb_z1, sb_z2;
  float* C; int64_t sc_m, sc_z1, sc_z2;
  const float* bias;
  const float* residual; int64_t ld_res;   // same batch offsets as C
  const float* gelu_u; int64_t ld_u;       // same batch offsets as C
  float* out_pre; int64_t ld_pre;          // same batch offsets as C
  int apply_gelu, accumulate;
  float alpha;
};

__global__ void __launch_bounds__(256) simt_gemm_kernel(const SimtParams p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int z = blockIdx.z;
  const int z1 = z / p.Z2, z2 = z % p.Z2;
  const float* A = p.A + z1 * p.sa_z1 + z2 * p.sa_z2;
  const float* B = p.B + z1 * p.sb_z1 + z2 * p.sb_z2;
  const int64_t c_off = z1 * p.sc_z1 + z2 * p.sc_z2;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: 64 rows x 16 k = 1024 elements, 4 per thread. k is the fast index when the
  // operand is K-major (stride_k == 1), otherwise rows are the fast index.
  const bool a_kfast = (p.sa_k == 1), b_kfast = (p.sb_k == 1);
  for (int k0 = 0; k0 < p.K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + 256 * i;
      int mm, kk;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < p.M && gk < p.K) ? A[gm * p.sa_m + gk * p.sa_k] : 0.f;
      int nn, kb;
      if (b_kfast) { kb = e & 15; nn = e >> 4; } else { nn = e & 63; kb = e >> 6; }
      const int gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < p.N && gkb < p.K) ? B[gn * p.sb_n + gkb * p.sb_k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (p.bias) v += p.bias[gn];
      if (p.gelu_u) v *= gelu_erf_grad(p.gelu_u[c_off + (int64_t)gm * p.ld_u + gn]);
      if (p.apply_gelu) {
        if (p.out_pre) p.out_pre[c_off + (int64_t)gm * p.ld_pre + gn] = v;
        v = gelu_erf(v);
      }
      if (p.residual) v += p.residual[c_off + (int64_t)gm * p.ld_res + gn];
      float* dst = p.C + c_off + (int64_t)gm * p.sc_m + gn;
      if (p.accumulate) v += *dst;
      *dst = v;
    }
  }
}

// in-place row softmax over the last dim (dots already scaled), one warp per row
__global__ void softmax_fwd_kernel(float* __restrict__ s, int64_t rows, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* r = s + row * n;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, r[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) { const float e = expf(r[j] - mx); r[j] = e; sum += e; }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < n; j += 32) r[j] *= inv;
}

// dS = P * (dP - sum_j P*dP), written over dP
__global__ void softmax_bwd_kernel(const float* __restrict__ P, float* __restrict__ dP, int64_t rows, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = P + row * n;
  float* d = dP + row * n;
  float dot = 0.f;
  for (int j = lane; j < n; j += 32) dot += p[j] * d[j];
  dot = warp_sum(dot);
  for (int j = lane; j < n; j += 32) d[j] = p[j] * (d[j] - dot);
}

}  // namespace

int nv_simt_gemm_launch(int M, int N, int K, int Z1, int Z2, const float* A, int64_t sa_m, int64_t sa_k,
                        int64_t sa_z1, int64_t sa_z2, const float* B, int64_t sb_n, int64_t sb_k, int64_t sb_z1,
                        int64_t sb_z2, float* C, int64_t sc_m, int64_t sc_z1, int64_t sc_z2, const float* bias,
                        const float* residual, int64_t ld_res, const float* gelu_u, int64_t ld_u, float* out_pre,
                        int64_t ld_pre, int apply_gelu, int accumulate, float alpha, cudaStream_t stream) {
  NV_REQUIRE(M > 0 && N > 0 && K > 0 && Z1 > 0 && Z2 > 0, "simt gemm: empty problem");
  NV_REQUIRE((int64_t)Z1 * Z2 <= 65535, "simt gemm: batch count %lld exceeds grid.z", (long long)Z1 * Z2);
  SimtParams p;
  p.M = M; p.N = N; p.K = K; p.Z2 = Z2;
  p.A = A; p.sa_m = sa_m; p.sa_k = sa_k; p.sa_z1 = sa_z1; p.sa_z2 = sa_z2;
  p.B = B; p.sb_n = sb_n; p.sb_k = sb_k; p.sb_z1 = sb_z1; p.sb_z2 = sb_z2;
  p.C = C; p.sc_m = sc_m; p.sc_z1 = sc_z1; p.sc_z2 = sc_z2;
  p.bias = bias; p.residual = residual; p.ld_res = ld_res; p.gelu_u = gelu_u; p.ld_u = ld_u;
  p.out_pre = out_pre; p.ld_pre = ld_pre; p.apply_gelu = apply_gelu; p.accumulate = accumulate; p.alpha = alpha;
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, Z1 * Z2);
  simt_gemm_kernel<<<grid, 256, 0, stream>>>(p);
  NV_LAUNCH_CHECK("simt_gemm_kernel");
  return NV_OK;
}

int nv_softmax_fwd_launch(float* s, int64_t rows, int n, cudaStream_t stream) {
  if (rows <= 0) return NV_OK;
  softmax_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(s, rows, n);
  NV_LAUNCH_CHECK("softmax_fwd_kernel");
  return NV_OK;
}

int nv_softmax_bwd_launch(const float* P, float* dP, int64_t rows, int n, cudaStream_t stream) {
  if (rows <= 0) return NV_OK;
  softmax_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(P, dP, rows, n);
  NV_LAUNCH_CHECK("softmax_bwd_kernel");
  return NV_OK;
}
