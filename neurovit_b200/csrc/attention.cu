// Flash-style multi-head softmax attention, forward and backward, head_dim = 64, bf16 operands with
// fp32 softmax statistics kept in registers and reduced with warp shuffles.
// Reference: src/models/vit_3d.py:51-59 (q,k,v split; dots = q k^T * scale; softmax(dim=-1); attn v;
// 'b h n d -> b n (h d)'). No mask, not causal; the score matrix [B,h,N,N] is never materialised.
//
//   fwd : O = softmax(Q K^T * scale) V,   LSE_i = log sum_j exp(s_ij)            (saved for backward)
//   bwd : D_i = sum_d dO_id O_id;  P = exp(S - LSE);  dV = P^T dO;  dP = dO V^T;
//         dS = P o (dP - D);  dQ = dS K * scale;  dK = dS^T Q * scale            
//
// Each CTA = 4 warps, 64 query (or key) rows, 16 rows per warp; tiles are 64x64 bf16 in 128B-swizzled
// shared memory fed by cp.async; tensor-core work is mma.sync m16n8k16 (attention is ~6% of the
// model's FLOPs at 385 tokens; the linear layers run on tcgen05, see gemm_tc.cu).
// Tensors are addressed as base + b*batch_stride + row*row_stride + h*64 (elements), so q/k/v are read
// in place from the [B,N,3*h*64] QKV GEMM output and O is written directly in [B,N,h*64].
#include "nv_common.cuh"

namespace {

constexpr int HD = 64;        // head dim
constexpr int BR = 64;        // rows per CTA tile
constexpr int TILE_BYTES = BR * HD * 2;  // 8 KB
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnTensor {
  const bf16* ptr;
  int64_t batch_stride, row_stride;  // elements; head offset is h*64
};
struct AttnTensorOut {
  bf16* ptr;
  int64_t batch_stride, row_stride;
};

__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {  // byte offset of a 16B chunk
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// async copy of a [64 x 64] bf16 tile (rows row0.. of a [N x 64] strided matrix) into swizzled smem;
// rows >= N are zero-filled. 128 threads, 4 chunks each.
__device__ __forceinline__ void load_tile_async(uint8_t* smem_tile, const bf16* g, int64_t row_stride, int row0,
                                                int N) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = threadIdx.x + 128 * i;
    const int r = e >> 3, ch = e & 7;
    const uint32_t dst = smem_u32(smem_tile) + tile_off(r, ch);
    const int gr = row0 + r;
    const bf16* src = g + (int64_t)(gr < N ? gr : 0) * row_stride + ch * 8;
    const int bytes = gr < N ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory");
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// A fragment: rows row_base..+15, k = 16*kstep..+15 of a row-major [rows x 64] tile
__device__ __forceinline__ void ldsm_a(uint32_t (&r)[4], const uint8_t* tile, int row_base, int kstep, int lane) {
  ldsm_x4(r, smem_u32(tile) + tile_off(row_base + (lane & 15), 2 * kstep + (lane >> 4)));
}
// B fragments for two n-tiles (n = n_base..+15) from a tile stored [n][k] (k contiguous):
// r[0],r[1] -> n-tile 0 (b0,b1); r[2],r[3] -> n-tile 1
__device__ __forceinline__ void ldsm_b_nk(uint32_t (&r)[4], const uint8_t* tile, int n_base, int kstep, int lane) {
  const int mi = lane >> 3;
  ldsm_x4(r, smem_u32(tile) + tile_off(n_base + ((mi >> 1) << 3) + (lane & 7), 2 * kstep + (mi & 1)));
}
// B fragments for two n-tiles (n chunks nc0, nc0+1) from a tile stored [k][n] (n contiguous), k = k_base..+15
__device__ __forceinline__ void ldsm_b_kn(uint32_t (&r)[4], const uint8_t* tile, int k_base, int nc0, int lane) {
  const int mi = lane >> 3;
  ldsm_x4_t(r, smem_u32(tile) + tile_off(k_base + ((mi & 1) << 3) + (lane & 7), nc0 + (mi >> 1)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// C[16 x 64] += A[16 x 64(k)] * Bt where the B tile is stored [n][k]
__device__ __forceinline__ void mm_a_bnk(float (&c)[8][4], const uint32_t (&a)[4][4], const uint8_t* btile, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t r[4];
      ldsm_b_nk(r, btile, 16 * np, ks, lane);
      mma_bf16(c[2 * np], a[ks], r[0], r[1]);
      mma_bf16(c[2 * np + 1], a[ks], r[2], r[3]);
    }
}
// C[16 x 64] += A[16 x 64(k)] * B where the B tile is stored [k][n]
__device__ __forceinline__ void mm_a_bkn(float (&c)[8][4], const uint32_t (&a)[4][4], const uint8_t* btile, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t r[4];
      ldsm_b_kn(r, btile, 16 * kk, 2 * dp, lane);
      mma_bf16(c[2 * dp], a[kk], r[0], r[1]);
      mma_bf16(c[2 * dp + 1], a[kk], r[2], r[3]);
    }
}
// accumulator tile (16 x 64 fp32, C layout) -> A fragments (bf16) for a following MMA over its columns
__device__ __forceinline__ void acc_to_afrag(uint32_t (&a)[4][4], const float (&c)[8][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = pack_bf16x2(c[2 * kk][0], c[2 * kk][1]);
    a[kk][1] = pack_bf16x2(c[2 * kk][2], c[2 * kk][3]);
    a[kk][2] = pack_bf16x2(c[2 * kk + 1][0], c[2 * kk + 1][1]);
    a[kk][3] = pack_bf16x2(c[2 * kk + 1][2], c[2 * kk + 1][3]);
  }
}
__device__ __forceinline__ void zero_acc(float (&c)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_fwd_kernel(AttnTensor q, AttnTensor k, AttnTensor v, AttnTensorOut o, float* __restrict__ lse, int N, int H,
                float scale) {
  __shared__ __align__(128) uint8_t sQ[TILE_BYTES];
  __shared__ __align__(128) uint8_t sK[2][TILE_BYTES];
  __shared__ __align__(128) uint8_t sV[2][TILE_BYTES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BR;
  const bf16* qp = q.ptr + (int64_t)b * q.batch_stride + h * HD;
  const bf16* kp = k.ptr + (int64_t)b * k.batch_stride + h * HD;
  const bf16* vp = v.ptr + (int64_t)b * v.batch_stride + h * HD;
  const int nkv = (N + BR - 1) / BR;

  load_tile_async(sQ, qp, q.row_stride, q0, N);
  load_tile_async(sK[0], kp, k.row_stride, 0, N);
  load_tile_async(sV[0], vp, v.row_stride, 0, N);
  cp_async_commit();

  uint32_t qf[4][4];
  float oacc[8][4];
  zero_acc(oacc);
  float m_run[2] = {-INFINITY, -INFINITY};  // running max (log2 domain) for rows g, g+8
  float l_run[2] = {0.f, 0.f};
  const float c = scale * LOG2E;

  for (int j = 0; j < nkv; ++j) {
    cp_async_wait<0>();
    __syncthreads();
    if (j + 1 < nkv) {
      load_tile_async(sK[(j + 1) & 1], kp, k.row_stride, (j + 1) * BR, N);
      load_tile_async(sV[(j + 1) & 1], vp, v.row_stride, (j + 1) * BR, N);
      cp_async_commit();
    }
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldsm_a(qf[ks], sQ, warp * 16, ks, lane);
    }
    float s[8][4];
    zero_acc(s);
    mm_a_bnk(s, qf, sK[j & 1], lane);
    // scale to log2 domain, mask keys >= N
    const int key0 = j * BR;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = key0 + nt * 8 + 2 * t + (e & 1);
        const float val = key < N ? s[nt][e] * c : -INFINITY;
        s[nt][e] = val;
        mx[e >> 1] = fmaxf(mx[e >> 1], val);
      }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);  // finite: every KV block has >= 1 valid key
      alpha[r] = exp2f(m_run[r] - m_new);
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(s[nt][e] - m_run[e >> 1]);
        s[nt][e] = p;
        rs[e >> 1] += p;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];  // per-thread partial sums
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      oacc[nt][0] *= alpha[0]; oacc[nt][1] *= alpha[0];
      oacc[nt][2] *= alpha[1]; oacc[nt][3] *= alpha[1];
    }
    uint32_t pf[4][4];
    acc_to_afrag(pf, s);
    mm_a_bkn(oacc, pf, sV[j & 1], lane);
  }
  // finalise: reduce row sums across the 4 lanes of a quad, normalise, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  bf16* op = o.ptr + (int64_t)b * o.batch_stride + h * HD;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + 8 * r;
    if (row < N) {
      const float inv = 1.0f / l_run[r];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<uint32_t*>(op + (int64_t)row * o.row_stride + nt * 8 + 2 * t) =
            pack_bf16x2(oacc[nt][2 * r] * inv, oacc[nt][2 * r + 1] * inv);
      if (t == 0) lse[((int64_t)b * H + h) * N + row] = (m_run[r] + log2f(l_run[r])) * LN2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward pre-pass: delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]
// ------------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(AttnTensor dO, AttnTensor O, float* __restrict__ delta, int B, int N, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)B * H * N) return;
  const int i = (int)(w % N);
  const int h = (int)((w / N) % H);
  const int b = (int)(w / ((int64_t)N * H));
  const uint32_t a = *reinterpret_cast<const uint32_t*>(dO.ptr + (int64_t)b * dO.batch_stride + (int64_t)i * dO.row_stride + h * HD + 2 * lane);
  const uint32_t c = *reinterpret_cast<const uint32_t*>(O.ptr + (int64_t)b * O.batch_stride + (int64_t)i * O.row_stride + h * HD + 2 * lane);
  const float2 fa = unpack_bf16x2(a), fc = unpack_bf16x2(c);
  const float s = warp_sum(fa.x * fc.x + fa.y * fc.y);
  if (lane == 0) delta[w] = s;
}

// ------------------------------------------------------------------------------------------------
// backward, dK / dV: one CTA per 64-key block, loops over query blocks; each warp owns 16 keys and
// works on the transposed score tile S^T[key, q] so dK/dV accumulate in registers.
// ------------------------------------------------------------------------------------------------
struct BwdSmem {
  uint8_t k[TILE_BYTES];
  uint8_t v[TILE_BYTES];
  uint8_t q[2][TILE_BYTES];
  uint8_t d[2][TILE_BYTES];
  float lse2[2][BR];
  float delta[2][BR];
};

__device__ __forceinline__ void load_rowstats(float* s_lse2, float* s_delta, const float* lse, const float* delta,
                                              int row0, int N) {
  if (threadIdx.x < BR) {
    const int r = row0 + threadIdx.x;
    s_lse2[threadIdx.x] = r < N ? lse[r] * LOG2E : INFINITY;  // +inf => P = 0 for padded rows
    s_delta[threadIdx.x] = r < N ? delta[r] : 0.f;
  }
}

__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(AttnTensor q, AttnTensor k, AttnTensor v, AttnTensor dO, const float* __restrict__ lse,
                    const float* __restrict__ delta, AttnTensorOut dk, AttnTensorOut dv, int N, int H, float scale) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * BR;
  const bf16* qp = q.ptr + (int64_t)b * q.batch_stride + h * HD;
  const bf16* kp = k.ptr + (int64_t)b * k.batch_stride + h * HD;
  const bf16* vp = v.ptr + (int64_t)b * v.batch_stride + h * HD;
  const bf16* dop = dO.ptr + (int64_t)b * dO.batch_stride + h * HD;
  const float* lse_bh = lse + ((int64_t)b * H + h) * N;
  const float* delta_bh = delta + ((int64_t)b * H + h) * N;
  const int nq = (N + BR - 1) / BR;
  const float c = scale * LOG2E;

  load_tile_async(sm.k, kp, k.row_stride, k0, N);
  load_tile_async(sm.v, vp, v.row_stride, k0, N);
  load_tile_async(sm.q[0], qp, q.row_stride, 0, N);
  load_tile_async(sm.d[0], dop, dO.row_stride, 0, N);
  cp_async_commit();
  load_rowstats(sm.lse2[0], sm.delta[0], lse_bh, delta_bh, 0, N);

  uint32_t kf[4][4], vf[4][4];
  float dk_acc[8][4], dv_acc[8][4];
  zero_acc(dk_acc);
  zero_acc(dv_acc);
  const int key_lo = k0 + warp * 16 + g;  // rows g and g+8 of this warp's slice

  for (int j = 0; j < nq; ++j) {
    cp_async_wait<0>();
    __syncthreads();
    if (j + 1 < nq) {
      load_tile_async(sm.q[(j + 1) & 1], qp, q.row_stride, (j + 1) * BR, N);
      load_tile_async(sm.d[(j + 1) & 1], dop, dO.row_stride, (j + 1) * BR, N);
      cp_async_commit();
      load_rowstats(sm.lse2[(j + 1) & 1], sm.delta[(j + 1) & 1], lse_bh, delta_bh, (j + 1) * BR, N);
    }
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        ldsm_a(kf[ks], sm.k, warp * 16, ks, lane);
        ldsm_a(vf[ks], sm.v, warp * 16, ks, lane);
      }
    }
    const uint8_t* sq = sm.q[j & 1];
    const uint8_t* sd = sm.d[j & 1];
    const float* s_lse2 = sm.lse2[j & 1];
    const float* s_delta = sm.delta[j & 1];
    float st[8][4];  // S^T tile: 16 keys x 64 queries
    zero_acc(st);
    mm_a_bnk(st, kf, sq, lane);
    float dpt[8][4];
    zero_acc(dpt);
    mm_a_bnk(dpt, vf, sd, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = nt * 8 + 2 * t + (e & 1);
        const int key = key_lo + 8 * (e >> 1);
        const float p = key < N ? exp2f(st[nt][e] * c - s_lse2[qi]) : 0.f;
        st[nt][e] = p;
        dpt[nt][e] = p * (dpt[nt][e] - s_delta[qi]);
      }
    uint32_t pf[4][4];
    acc_to_afrag(pf, st);
    mm_a_bkn(dv_acc, pf, sd, lane);  // dV += P^T dO
    acc_to_afrag(pf, dpt);
    mm_a_bkn(dk_acc, pf, sq, lane);  // dK += dS^T Q
  }
  bf16* dkp = dk.ptr + (int64_t)b * dk.batch_stride + h * HD;
  bf16* dvp = dv.ptr + (int64_t)b * dv.batch_stride + h * HD;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = key_lo + 8 * r;
    if (key < N) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<uint32_t*>(dkp + (int64_t)key * dk.row_stride + nt * 8 + 2 * t) =
            pack_bf16x2(dk_acc[nt][2 * r] * scale, dk_acc[nt][2 * r + 1] * scale);
        *reinterpret_cast<uint32_t*>(dvp + (int64_t)key * dv.row_stride + nt * 8 + 2 * t) =
            pack_bf16x2(dv_acc[nt][2 * r], dv_acc[nt][2 * r + 1]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, dQ: one CTA per 64-query block, loops over key blocks
// ------------------------------------------------------------------------------------------------
struct BwdQSmem {
  uint8_t q[TILE_BYTES];
  uint8_t d[TILE_BYTES];
  uint8_t k[2][TILE_BYTES];
  uint8_t v[2][TILE_BYTES];
};

__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(AttnTensor q, AttnTensor k, AttnTensor v, AttnTensor dO, const float* __restrict__ lse,
                   const float* __restrict__ delta, AttnTensorOut dq, int N, int H, float scale) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  BwdQSmem& sm = *reinterpret_cast<BwdQSmem*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BR;
  const bf16* qp = q.ptr + (int64_t)b * q.batch_stride + h * HD;
  const bf16* kp = k.ptr + (int64_t)b * k.batch_stride + h * HD;
  const bf16* vp = v.ptr + (int64_t)b * v.batch_stride + h * HD;
  const bf16* dop = dO.ptr + (int64_t)b * dO.batch_stride + h * HD;
  const int nkv = (N + BR - 1) / BR;
  const float c = scale * LOG2E;

  load_tile_async(sm.q, qp, q.row_stride, q0, N);
  load_tile_async(sm.d, dop, dO.row_stride, q0, N);
  load_tile_async(sm.k[0], kp, k.row_stride, 0, N);
  load_tile_async(sm.v[0], vp, v.row_stride, 0, N);
  cp_async_commit();

  float lse2[2], dl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + 8 * r;
    lse2[r] = row < N ? lse[((int64_t)b * H + h) * N + row] * LOG2E : INFINITY;
    dl[r] = row < N ? delta[((int64_t)b * H + h) * N + row] : 0.f;
  }
  uint32_t qf[4][4], df[4][4];
  float dq_acc[8][4];
  zero_acc(dq_acc);

  for (int j = 0; j < nkv; ++j) {
    cp_async_wait<0>();
    __syncthreads();
    if (j + 1 < nkv) {
      load_tile_async(sm.k[(j + 1) & 1], kp, k.row_stride, (j + 1) * BR, N);
      load_tile_async(sm.v[(j + 1) & 1], vp, v.row_stride, (j + 1) * BR, N);
      cp_async_commit();
    }
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        ldsm_a(qf[ks], sm.q, warp * 16, ks, lane);
        ldsm_a(df[ks], sm.d, warp * 16, ks, lane);
      }
    }
    const uint8_t* sk = sm.k[j & 1];
    const uint8_t* sv = sm.v[j & 1];
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mm_a_bnk(s, qf, sk, lane);   // S = Q K^T
    mm_a_bnk(dp, df, sv, lane);  // dP = dO V^T
    const int key0 = j * BR;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = key0 + nt * 8 + 2 * t + (e & 1);
        const float p = key < N ? exp2f(s[nt][e] * c - lse2[e >> 1]) : 0.f;
        s[nt][e] = p * (dp[nt][e] - dl[e >> 1]);
      }
    uint32_t dsf[4][4];
    acc_to_afrag(dsf, s);
    mm_a_bkn(dq_acc, dsf, sk, lane);  // dQ += dS K
  }
  bf16* dqp = dq.ptr + (int64_t)b * dq.batch_stride + h * HD;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + 8 * r;
    if (row < N) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<uint32_t*>(dqp + (int64_t)row * dq.row_stride + nt * 8 + 2 * t) =
            pack_bf16x2(dq_acc[nt][2 * r] * scale, dq_acc[nt][2 * r + 1] * scale);
    }
  }
}

int check_tensor(const void* p, int64_t bs, int64_t rs, const char* name) {
  NV_REQUIRE(p != nullptr, "attention: %s is null", name);
  NV_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0 && bs % 8 == 0 && rs % 8 == 0,
             "attention: %s must be 16-byte aligned with strides that are multiples of 8 elements", name);
  return NV_OK;
}

}  // namespace

int nv_attn_fwd_launch(const bf16* q, const bf16* k, const bf16* v, int64_t qkv_batch_stride, int64_t qkv_row_stride,
                       bf16* o, int64_t o_batch_stride, int64_t o_row_stride, float* lse, int B, int N, int H,
                       int head_dim, float scale, cudaStream_t stream) {
  NV_REQUIRE(head_dim == HD, "attention: head_dim %d unsupported by the bf16 flash kernel (needs 64)", head_dim);
  NV_REQUIRE(B >= 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "attention: bad sizes B=%d N=%d H=%d", B, N, H);
  if (B == 0) return NV_OK;
  int s;
  if ((s = check_tensor(q, qkv_batch_stride, qkv_row_stride, "q")) != NV_OK) return s;
  if ((s = check_tensor(k, qkv_batch_stride, qkv_row_stride, "k")) != NV_OK) return s;
  if ((s = check_tensor(v, qkv_batch_stride, qkv_row_stride, "v")) != NV_OK) return s;
  if ((s = check_tensor(o, o_batch_stride, o_row_stride, "o")) != NV_OK) return s;
  AttnTensor tq{q, qkv_batch_stride, qkv_row_stride}, tk{k, qkv_batch_stride, qkv_row_stride},
      tv{v, qkv_batch_stride, qkv_row_stride};
  AttnTensorOut to{o, o_batch_stride, o_row_stride};
  dim3 grid((N + BR - 1) / BR, H, B);
  attn_fwd_kernel<<<grid, 128, 0, stream>>>(tq, tk, tv, to, lse, N, H, scale);
  NV_LAUNCH_CHECK("attn_fwd_kernel");
  return NV_OK;
}

int nv_attn_bwd_launch(const bf16* q, const bf16* k, const bf16* v, int64_t qkv_batch_stride, int64_t qkv_row_stride,
                       const bf16* o, const bf16* dO, int64_t o_batch_stride, int64_t o_row_stride, const float* lse,
                       float* delta_ws, bf16* dq, bf16* dk, bf16* dv, int64_t dqkv_batch_stride,
                       int64_t dqkv_row_stride, int B, int N, int H, int head_dim, float scale, cudaStream_t stream) {
  NV_REQUIRE(head_dim == HD, "attention: head_dim %d unsupported by the bf16 flash kernel (needs 64)", head_dim);
  NV_REQUIRE(B >= 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "attention: bad sizes B=%d N=%d H=%d", B, N, H);
  if (B == 0) return NV_OK;
  int s;
  if ((s = check_tensor(q, qkv_batch_stride, qkv_row_stride, "q")) != NV_OK) return s;
  if ((s = check_tensor(k, qkv_batch_stride, qkv_row_stride, "k")) != NV_OK) return s;
  if ((s = check_tensor(v, qkv_batch_stride, qkv_row_stride, "v")) != NV_OK) return s;
  if ((s = check_tensor(o, o_batch_stride, o_row_stride, "o")) != NV_OK) return s;
  if ((s = check_tensor(dO, o_batch_stride, o_row_stride, "dO")) != NV_OK) return s;
  if ((s = check_tensor(dq, dqkv_batch_stride, dqkv_row_stride, "dq")) != NV_OK) return s;
  if ((s = check_tensor(dk, dqkv_batch_stride, dqkv_row_stride, "dk")) != NV_OK) return s;
  if ((s = check_tensor(dv, dqkv_batch_stride, dqkv_row_stride, "dv")) != NV_OK) return s;
  AttnTensor tq{q, qkv_batch_stride, qkv_row_stride}, tk{k, qkv_batch_stride, qkv_row_stride},
      tv{v, qkv_batch_stride, qkv_row_stride}, tO{o, o_batch_stride, o_row_stride},
      tdO{dO, o_batch_stride, o_row_stride};
  AttnTensorOut tdq{dq, dqkv_batch_stride, dqkv_row_stride}, tdk{dk, dqkv_batch_stride, dqkv_row_stride},
      tdv{dv, dqkv_batch_stride, dqkv_row_stride};
  const int64_t rows = (int64_t)B * H * N;
  attn_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(tdO, tO, delta_ws, B, N, H);
  NV_LAUNCH_CHECK("attn_delta_kernel");
  dim3 grid((N + BR - 1) / BR, H, B);
  static bool attr_set = false;
  if (!attr_set) {
    NV_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
    NV_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdQSmem)));
    attr_set = true;
  }
  attn_bwd_dkv_kernel<<<grid, 128, sizeof(BwdSmem), stream>>>(tq, tk, tv, tdO, lse, delta_ws, tdk, tdv, N, H, scale);
  NV_LAUNCH_CHECK("attn_bwd_dkv_kernel");
  attn_bwd_dq_kernel<<<grid, 128, sizeof(BwdQSmem), stream>>>(tq, tk, tv, tdO, lse, delta_ws, tdq, N, H, scale);
  NV_LAUNCH_CHECK("attn_bwd_dq_kernel");
  return NV_OK;
}
