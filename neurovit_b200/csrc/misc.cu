// Small HBM-bound helpers around the GEMMs: dtype casts (weight caches), column sums (bias
// gradients), batch sums (pos/cls gradients, vit_3d.py:98-99,116-118) and token pooling (:123).
#include "nv_common.cuh"

namespace {

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t n) {
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[i];
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    out[i] = __float2bfloat16(in[i]);
  }
}

// out[r, c] = bf16(in[r, c]) (optional) and outT[c, r] = bf16(in[r, c]); 32x32 smem tile transpose
__global__ void cast_transpose_kernel(const float* __restrict__ in, bf16* __restrict__ out,
                                      bf16* __restrict__ outT, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) {
      v = in[(int64_t)r * C + c];
      if (out) out[(int64_t)r * C + c] = __float2bfloat16(v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) outT[(int64_t)c * R + r] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

// out[c] += sum_r in[r, c]; block = 32 columns x 8 row-lanes, grid.y splits the rows
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ in, int64_t ld, float* __restrict__ out, int M, int N,
                              int rows_per_block) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r_begin = blockIdx.y * rows_per_block;
  const int r_end = min(M, r_begin + rows_per_block);
  float s = 0.f;
  if (c < N)
    for (int r = r_begin + threadIdx.y; r < r_end; r += 8) s += to_f32<T>(in[(int64_t)r * ld + c]);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// out[j] += sum_b in[b * batch_stride + j], j < L   (d_pos = sum over the batch of dx0)
__global__ void batch_sum_kernel(const float* __restrict__ in, int64_t batch_stride, float* __restrict__ out,
                                 int B, int64_t L) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < L; j += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += in[(int64_t)b * batch_stride + j];
    out[j] += s;
  }
}

// pooled[b, :] = mean_t x[b, t, :]   (pool='mean', vit_3d.py:123)
__global__ void mean_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ pooled, int B, int N, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, c = i % D;
  float s = 0.f;
  for (int t = 0; t < N; ++t) s += x[((int64_t)b * N + t) * D + c];
  pooled[i] = s / (float)N;
}
__global__ void mean_pool_bwd_kernel(const float* __restrict__ dpooled, float* __restrict__ dx,
                                     bf16* __restrict__ dx_bf16, int B, int N, int D) {
  const int64_t total = (int64_t)B * N * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % D);
    const int b = (int)(i / ((int64_t)N * D));
    const float v = dpooled[(int64_t)b * D + c] / (float)N;
    dx[i] = v;
    if (dx_bf16) dx_bf16[i] = __float2bfloat16(v);
  }
}

}  // namespace

int nv_cast_f32_bf16_launch(const float* in, bf16* out, int64_t n, cudaStream_t stream) {
  if (n <= 0) return NV_OK;
  NV_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0,
             "cast: pointers must be 16B/8B aligned");
  int64_t blocks = ((n >> 2) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > nv_num_sms() * 8) blocks = nv_num_sms() * 8;
  cast_f32_bf16_kernel<<<(int)blocks, 256, 0, stream>>>(in, out, n);
  NV_LAUNCH_CHECK("cast_f32_bf16_kernel");
  return NV_OK;
}

int nv_cast_transpose_launch(const float* in, bf16* out, bf16* outT, int R, int C, cudaStream_t stream) {
  NV_REQUIRE(R > 0 && C > 0 && outT != nullptr, "cast_transpose: bad arguments");
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  cast_transpose_kernel<<<grid, block, 0, stream>>>(in, out, outT, R, C);
  NV_LAUNCH_CHECK("cast_transpose_kernel");
  return NV_OK;
}

int nv_colsum_launch(const void* in, int in_is_bf16, int64_t ld, float* out, int M, int N, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return NV_OK;
  const int col_blocks = (N + 31) / 32;
  int row_blocks = (nv_num_sms() * 4 + col_blocks - 1) / col_blocks;
  if (row_blocks < 1) row_blocks = 1;
  int rows_per_block = (M + row_blocks - 1) / row_blocks;
  if (rows_per_block < 64) rows_per_block = 64;
  row_blocks = (M + rows_per_block - 1) / rows_per_block;
  dim3 grid(col_blocks, row_blocks), block(32, 8);
  if (in_is_bf16)
    colsum_kernel<bf16><<<grid, block, 0, stream>>>((const bf16*)in, ld, out, M, N, rows_per_block);
  else
    colsum_kernel<float><<<grid, block, 0, stream>>>((const float*)in, ld, out, M, N, rows_per_block);
  NV_LAUNCH_CHECK("colsum_kernel");
  return NV_OK;
}

int nv_batch_sum_launch(const float* in, int64_t batch_stride, float* out, int B, int64_t L, cudaStream_t stream) {
  if (B <= 0 || L <= 0) return NV_OK;
  int64_t blocks = (L + 255) / 256;
  if (blocks > nv_num_sms() * 8) blocks = nv_num_sms() * 8;
  batch_sum_kernel<<<(int)blocks, 256, 0, stream>>>(in, batch_stride, out, B, L);
  NV_LAUNCH_CHECK("batch_sum_kernel");
  return NV_OK;
}

int nv_mean_pool_fwd_launch(const float* x, float* pooled, int B, int N, int D, cudaStream_t stream) {
  if (B <= 0) return NV_OK;
  mean_pool_fwd_kernel<<<(B * D + 255) / 256, 256, 0, stream>>>(x, pooled, B, N, D);
  NV_LAUNCH_CHECK("mean_pool_fwd_kernel");
  return NV_OK;
}

int nv_mean_pool_bwd_launch(const float* dpooled, float* dx, bf16* dx_bf16, int B, int N, int D,
                            cudaStream_t stream) {
  if (B <= 0) return NV_OK;
  int64_t blocks = ((int64_t)B * N * D + 255) / 256;
  if (blocks > nv_num_sms() * 8) blocks = nv_num_sms() * 8;
  mean_pool_bwd_kernel<<<(int)blocks, 256, 0, stream>>>(dpooled, dx, dx_bf16, B, N, D);
  NV_LAUNCH_CHECK("mean_pool_bwd_kernel");
  return NV_OK;
}
