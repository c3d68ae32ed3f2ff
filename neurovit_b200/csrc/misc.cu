// Small HBM-bound helpers around the GEMMs: dtype casts (weight caches), column sums (bias
// gradients), batch sums (pos/cls gradients, vit_3d.py:98-99,116-118) and token pooling (:123).
#include "nv_common.cuh"
#include "nv_rng.cuh"
#include <math.h>

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

// Element-wise dropout with the same (seed, stream, row * N + col) mask the GEMM epilogues draw:
//   v = in * keep * 1/(1-p);  out = v (+ residual);  colsum[c] += sum_r v[r, c]
// Used for the embedding dropout (vit_3d.py:100,119), for masking the branch gradient in backward, and by
// the fp32 verification mode (whose CUDA-core GEMM has no fused dropout). Column-owner threads walk rows.
__global__ void dropout_kernel(const float* __restrict__ in, int64_t ld_in, const float* __restrict__ residual,
                               int64_t ld_res, float* __restrict__ out_f32, int64_t ld_f32, bf16* __restrict__ out_bf16,
                               int64_t ld_bf16, float* __restrict__ colsum, int M, int N, uint32_t thr, float ks,
                               uint64_t seed_host, uint32_t stream_id, const uint64_t* epoch, int row_mul) {
  const uint64_t seed = nv_seed(seed_host, epoch);
  const int n8 = N >> 3;  // one Philox call per thread and row: 8 consecutive columns
  constexpr int R = 4;     // rows in flight per thread (independent 32-byte loads)
  for (int c = threadIdx.x; c < n8; c += blockDim.x) {
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
    for (int rb = blockIdx.x * R; rb < M; rb += gridDim.x * R) {
      float4 v0[R], v1[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int r = min(rb + k, M - 1);
        const float* src = in + (int64_t)r * ld_in + 8 * c;
        v0[k] = *reinterpret_cast<const float4*>(src);
        v1[k] = *reinterpret_cast<const float4*>(src + 4);
      }
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int r = rb + k;
        if (r >= M) break;
        if (thr != 0) {
          const uint32_t b8 = nv_keep_bits8(seed, ((uint64_t)r * row_mul * N + 8 * c) >> 3, stream_id, thr);
          v0[k] = nv_dropout4(v0[k], b8 & 0xFu, ks);
          v1[k] = nv_dropout4(v1[k], b8 >> 4, ks);
        }
        acc0.x += v0[k].x; acc0.y += v0[k].y; acc0.z += v0[k].z; acc0.w += v0[k].w;
        acc1.x += v1[k].x; acc1.y += v1[k].y; acc1.z += v1[k].z; acc1.w += v1[k].w;
        if (residual) {
          const float* rs = residual + (int64_t)r * ld_res + 8 * c;
          const float4 r0 = *reinterpret_cast<const float4*>(rs), r1 = *reinterpret_cast<const float4*>(rs + 4);
          v0[k].x += r0.x; v0[k].y += r0.y; v0[k].z += r0.z; v0[k].w += r0.w;
          v1[k].x += r1.x; v1[k].y += r1.y; v1[k].z += r1.z; v1[k].w += r1.w;
        }
        if (out_f32) {
          float* dst = out_f32 + (int64_t)r * ld_f32 + 8 * c;
          *reinterpret_cast<float4*>(dst) = v0[k];
          *reinterpret_cast<float4*>(dst + 4) = v1[k];
        }
        if (out_bf16)
          *reinterpret_cast<uint4*>(out_bf16 + (int64_t)r * ld_bf16 + 8 * c) =
              make_uint4(pack_bf16x2(v0[k].x, v0[k].y), pack_bf16x2(v0[k].z, v0[k].w), pack_bf16x2(v1[k].x, v1[k].y),
                         pack_bf16x2(v1[k].z, v1[k].w));
      }
    }
    if (colsum) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum + 8 * c), "f"(acc0.x), "f"(acc0.y),
                   "f"(acc0.z), "f"(acc0.w) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum + 8 * c + 4), "f"(acc1.x), "f"(acc1.y),
                   "f"(acc1.z), "f"(acc1.w) : "memory");
    }
  }
}

// ---- classification head on the cls token: LayerNorm(dim) -> Linear(dim, num_classes) (vit_3d.py:105-110,123-126) -----
// One CTA per sample, fp32 throughout. Replaces a 64-row LayerNorm launch plus three [B,1024]x[1024,2]-sized
// GEMM launches of the generic CUDA-core kernel (33 us each: one CTA walking K serially) in each direction.
constexpr int HEAD_T = 256;
__device__ __forceinline__ float head_block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < HEAD_T / 32; ++i) s += red[i];
  return s;
}

__global__ void __launch_bounds__(HEAD_T)
head_fwd_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ y,
                float* __restrict__ mean_out, float* __restrict__ rstd_out, float* __restrict__ logits, int D, int C,
                float eps) {
  __shared__ float red[HEAD_T / 32];
  const int b = blockIdx.x;
  const float* xr = x + (int64_t)b * ld_x;
  float s = 0.f;
  for (int i = threadIdx.x; i < D; i += HEAD_T) s += xr[i];
  const float mean = head_block_sum(s, red) / (float)D;
  float q = 0.f;
  for (int i = threadIdx.x; i < D; i += HEAD_T) { const float d = xr[i] - mean; q += d * d; }
  const float rstd = rsqrtf(head_block_sum(q, red) / (float)D + eps);
  if (threadIdx.x == 0) { mean_out[b] = mean; rstd_out[b] = rstd; }
  for (int i = threadIdx.x; i < D; i += HEAD_T) y[(int64_t)b * D + i] = (xr[i] - mean) * rstd * gamma[i] + beta[i];
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < D; i += HEAD_T)
      acc += ((xr[i] - mean) * rstd * gamma[i] + beta[i]) * W[(int64_t)c * D + i];
    acc = head_block_sum(acc, red);
    if (threadIdx.x == 0) logits[(int64_t)b * C + c] = acc + (bias ? bias[c] : 0.f);
  }
}

// dx row (fp32 + optional bf16) of the cls token; dW, db, dgamma, dbeta accumulated with atomics (zero them first)
__global__ void __launch_bounds__(HEAD_T)
head_bwd_kernel(const float* __restrict__ dl, const float* __restrict__ x, int64_t ld_x, const float* __restrict__ y,
                const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ gamma,
                const float* __restrict__ W, float* __restrict__ dx, int64_t ld_dx, bf16* __restrict__ dx_bf16,
                int64_t ld_dxb, float* __restrict__ dW, float* __restrict__ db, float* __restrict__ dgamma,
                float* __restrict__ dbeta, int D, int C) {
  __shared__ float red[HEAD_T / 32];
  const int b = blockIdx.x;
  const float* xr = x + (int64_t)b * ld_x;
  const float mean = mean_in[b], rstd = rstd_in[b];
  // pass 1: dy_i = sum_c dl[b,c] W[c,i]; LayerNorm backward sums; parameter gradients
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < D; i += HEAD_T) {
    float dy = 0.f;
    for (int c = 0; c < C; ++c) {
      const float g = dl[(int64_t)b * C + c];
      dy = fmaf(g, W[(int64_t)c * D + i], dy);
      atomicAdd(dW + (int64_t)c * D + i, g * y[(int64_t)b * D + i]);
    }
    const float xh = (xr[i] - mean) * rstd;
    atomicAdd(dgamma + i, dy * xh);
    atomicAdd(dbeta + i, dy);
    const float gh = dy * gamma[i];
    s1 += gh;
    s2 += gh * xh;
  }
  if (threadIdx.x < C) atomicAdd(db + threadIdx.x, dl[(int64_t)b * C + threadIdx.x]);
  s1 = head_block_sum(s1, red) / (float)D;
  s2 = head_block_sum(s2, red) / (float)D;
  for (int i = threadIdx.x; i < D; i += HEAD_T) {
    float dy = 0.f;
    for (int c = 0; c < C; ++c) dy = fmaf(dl[(int64_t)b * C + c], W[(int64_t)c * D + i], dy);
    const float xh = (xr[i] - mean) * rstd;
    const float o = rstd * (dy * gamma[i] - s1 - xh * s2);
    dx[(int64_t)b * ld_dx + i] = o;
    if (dx_bf16) dx_bf16[(int64_t)b * ld_dxb + i] = __float2bfloat16(o);
  }
}

// Keep-bit generator: out[g] = the 8 keep bits of Philox group g (= elements [8g, 8g+8) of a dropout site, or
// one byte of the attention mask [B*H, N, 4*ceil(N/32)]). The consumers (GEMM epilogues, LayerNorm-backward
// side-car, attention forward) accept these bytes instead of drawing the bits inline: the Philox arithmetic then
// runs on a side stream beside the block's (HBM-bound) LayerNorm instead of inside a GEMM epilogue or the softmax
// rows. Bits are identical to the inline draw (same seed, stream, epoch, index). Four groups = two Philox calls per
// word (nv_keep_bits32: pair-shared calls, byte transposes, bit-sliced compare).
__global__ void dropout_bits_kernel(uint32_t* __restrict__ out, int64_t n_words, uint32_t thr, uint64_t seed_host,
                                    uint32_t stream_id, const uint64_t* epoch) {
  const uint64_t seed = nv_seed(seed_host, epoch);
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
    out[w] = nv_keep_bits32(seed, (uint64_t)w, stream_id, thr);   // = four nv_keep_bits8 groups, one per byte
  }
}

// Fused AdamW over the trainer's flat buffers (reference: optim.AdamW(lr, weight_decay) at src/Trainer.py:31,75;
// torch semantics: decoupled weight decay, bias-corrected moments, eps added to sqrt(v_hat)). One pass reads
// p, g, m, v and writes p, m, v plus the bf16 copy of p that the next forward's GEMMs read (the weight cache),
// instead of ~80 per-tensor chunks in three multi-tensor launches and 25 cast kernels.
__global__ void adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, bf16* __restrict__ p_bf16, int64_t n4, float lr, float beta1,
                                  float beta2, float eps, float decay, float inv_bc1, float inv_sqrt_bc2,
                                  const float* __restrict__ step_dev) {
  if (step_dev) {  // step count kept on the device (CUDA-graph replays): bias corrections computed here
    const float t = __ldg(step_dev);
    inv_bc1 = 1.0f / (1.0f - powf(beta1, t));
    inv_sqrt_bc2 = rsqrtf(1.0f - powf(beta2, t));
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* pe = &pp.x; const float* ge = &gg.x; float* me = &mm.x; float* ve = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      pe[k] *= decay;  // 1 - lr * weight_decay
      me[k] = beta1 * me[k] + (1.0f - beta1) * ge[k];
      ve[k] = beta2 * ve[k] + (1.0f - beta2) * ge[k] * ge[k];
      const float denom = sqrtf(ve[k]) * inv_sqrt_bc2 + eps;
      pe[k] -= lr * inv_bc1 * me[k] / denom;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (p_bf16)
      reinterpret_cast<uint2*>(p_bf16)[i] = make_uint2(pack_bf16x2(pp.x, pp.y), pack_bf16x2(pp.z, pp.w));
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

int nv_dropout_launch(const float* in, int64_t ld_in, const float* residual, int64_t ld_res, float* out_f32,
                      int64_t ld_f32, bf16* out_bf16, int64_t ld_bf16, float* colsum, int M, int N, float p,
                      uint64_t seed, int stream_id, int row_mul, cudaStream_t stream) {
  NV_REQUIRE(M >= 0 && N > 0 && N % 8 == 0, "dropout: N=%d must be a positive multiple of 8", N);
  NV_REQUIRE(p >= 0.f && p < 1.f, "dropout: p %f out of range [0, 1)", p);
  NV_REQUIRE(in != nullptr && (out_f32 != nullptr || out_bf16 != nullptr || colsum != nullptr), "dropout: null buffers");
  NV_REQUIRE(ld_in % 4 == 0 && ld_res % 4 == 0 && ld_f32 % 4 == 0 && ld_bf16 % 8 == 0,
             "dropout: fp32 row strides must be multiples of 4 elements, bf16 of 8");
  NV_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(out_f32) |
               reinterpret_cast<uintptr_t>(out_bf16)) & 15) == 0, "dropout: buffers must be 16-byte aligned");
  NV_REQUIRE((reinterpret_cast<uintptr_t>(colsum) & 15) == 0, "dropout: colsum must be 16-byte aligned");
  if (M == 0) return NV_OK;
  const uint32_t thr = nv_dropout_threshold(p);
  const int threads = (N / 8) >= 128 ? 128 : ((N / 8 + 31) / 32) * 32;
  int grid = nv_num_sms() * 8;
  if (grid > (M + 3) / 4) grid = (M + 3) / 4;
  dropout_kernel<<<grid, threads, 0, stream>>>(in, ld_in, residual, ld_res, out_f32, ld_f32, out_bf16, ld_bf16, colsum,
                                               M, N, thr, nv_dropout_keep_scale(thr), seed, (uint32_t)stream_id,
                                               thr != 0 ? nv_rng_epoch_dev() : nullptr, row_mul > 0 ? row_mul : 1);
  NV_LAUNCH_CHECK("dropout_kernel");
  return NV_OK;
}

int nv_adamw_flat_launch(float* p, const float* g, float* m, float* v, bf16* p_bf16, int64_t n, float lr, float beta1,
                         float beta2, float eps, float weight_decay, int step, const float* step_dev,
                         cudaStream_t stream) {
  NV_REQUIRE(n >= 0 && n % 4 == 0, "adamw: flat length %lld must be a multiple of 4", (long long)n);
  NV_REQUIRE(step >= 1 || step_dev != nullptr, "adamw: step counts from 1");
  if (step < 1) step = 1;
  NV_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
               reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0,
             "adamw: buffers must be 16-byte aligned");
  if (n == 0) return NV_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const int64_t n4 = n / 4;
  int grid = nv_num_sms() * 8;
  if ((int64_t)grid * 256 > n4) grid = (int)((n4 + 255) / 256);
  adamw_flat_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, p_bf16, n4, lr, beta1, beta2, eps, 1.0f - lr * weight_decay,
                                              (float)(1.0 / bc1), (float)(1.0 / sqrt(bc2)), step_dev);
  NV_LAUNCH_CHECK("adamw_flat_kernel");
  return NV_OK;
}

int nv_dropout_bits_launch(uint32_t* out, int64_t n_groups, float p, uint64_t seed, int stream_id, cudaStream_t stream) {
  NV_REQUIRE(n_groups >= 0 && n_groups % 4 == 0, "dropout_bits: the number of 8-element groups (%lld) must be a multiple of 4",
             (long long)n_groups);
  NV_REQUIRE(p > 0.f && p < 1.f, "dropout_bits: p %f out of range (0, 1)", p);
  NV_REQUIRE(out != nullptr && (reinterpret_cast<uintptr_t>(out) & 3) == 0, "dropout_bits: output must be 4-byte aligned");
  if (n_groups == 0) return NV_OK;
  const int64_t n_words = n_groups / 4;
  int grid = nv_num_sms() * 8;
  if ((int64_t)grid * 128 > n_words) grid = (int)((n_words + 127) / 128);
  dropout_bits_kernel<<<grid, 128, 0, stream>>>(out, n_words, nv_dropout_threshold(p), seed, (uint32_t)stream_id,
                                                nv_rng_epoch_dev());
  NV_LAUNCH_CHECK("dropout_bits_kernel");
  return NV_OK;
}

int nv_head_fwd_launch(const float* x, int64_t ld_x, const float* gamma, const float* beta, const float* W,
                       const float* bias, float* y, float* mean, float* rstd, float* logits, int B, int D, int C,
                       float eps, cudaStream_t stream) {
  NV_REQUIRE(B >= 0 && D > 0 && C > 0 && C <= HEAD_T, "head: bad sizes B=%d D=%d C=%d", B, D, C);
  if (B == 0) return NV_OK;
  head_fwd_kernel<<<B, HEAD_T, 0, stream>>>(x, ld_x, gamma, beta, W, bias, y, mean, rstd, logits, D, C, eps);
  NV_LAUNCH_CHECK("head_fwd_kernel");
  return NV_OK;
}

int nv_head_bwd_launch(const float* dl, const float* x, int64_t ld_x, const float* y, const float* mean,
                       const float* rstd, const float* gamma, const float* W, float* dx, int64_t ld_dx, bf16* dx_bf16,
                       int64_t ld_dxb, float* dW, float* db, float* dgamma, float* dbeta, int B, int D, int C,
                       cudaStream_t stream) {
  NV_REQUIRE(B >= 0 && D > 0 && C > 0 && C <= HEAD_T, "head: bad sizes B=%d D=%d C=%d", B, D, C);
  if (B == 0) return NV_OK;
  head_bwd_kernel<<<B, HEAD_T, 0, stream>>>(dl, x, ld_x, y, mean, rstd, gamma, W, dx, ld_dx, dx_bf16, ld_dxb, dW, db,
                                            dgamma, dbeta, D, C);
  NV_LAUNCH_CHECK("head_bwd_kernel");
  return NV_OK;
}
