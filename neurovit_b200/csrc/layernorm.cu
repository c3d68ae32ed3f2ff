// LayerNorm forward/backward and the embedding finalisation (HBM-bound kernels).
// Reference call sites: nn.LayerNorm at src/models/vit_3d.py:18,37,93,95,108; cls/pos add at :116-118.
// Math: y=(x-mu)*rstd*g+b, biased variance, eps inside the sqrt.
//   dx = rstd*(gh - mean(gh) - xh*mean(gh*xh)),  gh = dy*g,  dg = sum_rows dy*xh,  db = sum_rows dy.
// One warp owns one row; a lane owns the float4 chunks {lane + 32*j}, so per-column partial sums for
// dg/db stay in registers across the grid-stride row loop and are reduced once per CTA.
#include "nv_common.cuh"
#include "nv_rng.cuh"
#include <stdlib.h>

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
// optional colsum_out[D] += sum_rows dx (bias gradient of the linear that produced x's residual branch).
// Column-owner layout: thread t owns the float4 column chunk t (blockDim = D/4 rounded up to a warp), the
// CTA walks the rows four at a time (12 independent 16-byte loads in flight per thread, ~60 registers, so
// several CTAs stay resident per SM), the two per-row reductions go warp-shuffle -> shared -> broadcast,
// and the per-column dgamma/dbeta/colsum partials stay in 12 registers for the whole grid-stride loop.
constexpr int LNB_ROWS = 4;
constexpr int LNB_MAX_WARPS = 16;

// dy is fp32 (autograd tensors) or bf16 (the dgrad GEMM's output in bf16 mode: half the read traffic)
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4<bf16>(const bf16* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename DyT>
__global__ void __launch_bounds__(LNB_MAX_WARPS * 32)
ln_bwd_kernel(const DyT* __restrict__ dy, int64_t ld_dy, RowMap dymap, const float* __restrict__ x,
              int64_t ld_x, RowMap xmap, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const float* __restrict__ gamma, const float* __restrict__ dres, int64_t ld_dres,
              float* __restrict__ dx, int64_t ld_dx, RowMap dxmap, bf16* __restrict__ dx_bf16,
              int64_t ld_dxb, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ colsum_out, int M, int D) {
  __shared__ float red[2][LNB_MAX_WARPS][2 * LNB_ROWS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int c = threadIdx.x;  // float4 column chunk owned by this thread
  const bool active = c < (D >> 2);
  const float inv_d = 1.0f / (float)D;
  float4 acc_g = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = acc_g, acc_c = acc_g;
  const float4 g = active ? *reinterpret_cast<const float4*>(gamma + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  int buf = 0;
  for (int r0 = blockIdx.x * LNB_ROWS; r0 < M; r0 += gridDim.x * LNB_ROWS) {
    float4 xv[LNB_ROWS], dv[LNB_ROWS], rv[LNB_ROWS];
    float mean[LNB_ROWS], rstd[LNB_ROWS];
#pragma unroll
    for (int k = 0; k < LNB_ROWS; ++k) {
      const int r = r0 + k;
      const bool ok = active && r < M;
      xv[k] = ok ? *reinterpret_cast<const float4*>(x + xmap(r) * ld_x + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      dv[k] = ok ? ld4<DyT>(dy + dymap(r) * ld_dy + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      rv[k] = (ok && dres) ? *reinterpret_cast<const float4*>(dres + (int64_t)r * ld_dres + 4 * c)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      mean[k] = r < M ? mean_in[r] : 0.f;
      rstd[k] = r < M ? rstd_in[r] : 0.f;
    }
    float4 xh[LNB_ROWS], gh[LNB_ROWS];
    float part[2 * LNB_ROWS];
#pragma unroll
    for (int k = 0; k < LNB_ROWS; ++k) {
      const bool ok = active && (r0 + k) < M;
      xh[k].x = ok ? (xv[k].x - mean[k]) * rstd[k] : 0.f;
      xh[k].y = ok ? (xv[k].y - mean[k]) * rstd[k] : 0.f;
      xh[k].z = ok ? (xv[k].z - mean[k]) * rstd[k] : 0.f;
      xh[k].w = ok ? (xv[k].w - mean[k]) * rstd[k] : 0.f;
      acc_g.x += dv[k].x * xh[k].x; acc_g.y += dv[k].y * xh[k].y;
      acc_g.z += dv[k].z * xh[k].z; acc_g.w += dv[k].w * xh[k].w;
      acc_b.x += dv[k].x; acc_b.y += dv[k].y; acc_b.z += dv[k].z; acc_b.w += dv[k].w;
      gh[k].x = dv[k].x * g.x; gh[k].y = dv[k].y * g.y; gh[k].z = dv[k].z * g.z; gh[k].w = dv[k].w * g.w;
      part[2 * k] = (gh[k].x + gh[k].y) + (gh[k].z + gh[k].w);
      part[2 * k + 1] = (gh[k].x * xh[k].x + gh[k].y * xh[k].y) + (gh[k].z * xh[k].z + gh[k].w * xh[k].w);
    }
#pragma unroll
    for (int i = 0; i < 2 * LNB_ROWS; ++i) part[i] = warp_sum(part[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 2 * LNB_ROWS; ++i) red[buf][warp][i] = part[i];
    }
    __syncthreads();  // one barrier per 4 rows; `red` is double-buffered so the next iteration may overwrite
#pragma unroll
    for (int i = 0; i < 2 * LNB_ROWS; ++i) {
      float s = 0.f;
      for (int w = 0; w < nwarps; ++w) s += red[buf][w][i];
      part[i] = s * inv_d;
    }
    buf ^= 1;
#pragma unroll
    for (int k = 0; k < LNB_ROWS; ++k) {
      const int r = r0 + k;
      if (!(active && r < M)) continue;
      const float s1 = part[2 * k], s2 = part[2 * k + 1];
      float4 o;
      o.x = rstd[k] * (gh[k].x - s1 - xh[k].x * s2) + rv[k].x;
      o.y = rstd[k] * (gh[k].y - s1 - xh[k].y * s2) + rv[k].y;
      o.z = rstd[k] * (gh[k].z - s1 - xh[k].z * s2) + rv[k].z;
      o.w = rstd[k] * (gh[k].w - s1 - xh[k].w * s2) + rv[k].w;
      acc_c.x += o.x; acc_c.y += o.y; acc_c.z += o.z; acc_c.w += o.w;
      const int64_t orow = dxmap(r);
      if (dx) *reinterpret_cast<float4*>(dx + orow * ld_dx + 4 * c) = o;
      if (dx_bf16) OutStore<bf16>::st(dx_bf16 + orow * ld_dxb + 4 * c, o);
    }
  }
  if (active) {
    if (dgamma)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dgamma + 4 * c), "f"(acc_g.x), "f"(acc_g.y),
                   "f"(acc_g.z), "f"(acc_g.w) : "memory");
    if (dbeta)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbeta + 4 * c), "f"(acc_b.x), "f"(acc_b.y),
                   "f"(acc_b.z), "f"(acc_b.w) : "memory");
    if (colsum_out)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum_out + 4 * c), "f"(acc_c.x),
                   "f"(acc_c.y), "f"(acc_c.z), "f"(acc_c.w) : "memory");
  }
}

// ---- LayerNorm backward, bulk-copy pipelined variant (the one the transformer blocks use) --------------
// Same column-owner arithmetic as ln_bwd_kernel, but the three input streams (dy, x, dres) are staged
// through shared memory by the TMA engine. Warp specialised: the LAST warp is the producer (one lane issues
// 1-D bulk copies, one per row and tensor, 2-4 KB each, into a ring of `stages` LNP_ROWS-row stages tracked
// by full/empty mbarriers); the other warps own one float4 column chunk per thread and synchronise among
// themselves with a named barrier, so the copy-issue latency never sits on the compute warps' critical
// path. Sized for two CTAs per SM (~160 KB of reads in flight per SM regardless of register pressure, and
// one CTA's reduction latency hides behind the other's). The register-staged kernel above keeps only 16
// warps per SM resident and tops out near 3 TB/s.
constexpr int LNP_ROWS = 2;
constexpr int LNP_MAX_STAGES = 8;

// NT = compile-time bound on the compute threads (256 covers D <= 1024 with two CTAs per SM and up to ~112
// registers per thread, so the shuffle / shared-memory reduction chains of the rows stay independent;
// 512 covers D <= 2048)
template <typename DyT, int NT>
__global__ void __launch_bounds__(NT + 32, NT <= 256 ? 2 : 1)
ln_bwd_pipe_kernel(const DyT* __restrict__ dy, int64_t ld_dy, RowMap dymap, const float* __restrict__ x,
                   int64_t ld_x, RowMap xmap, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                   const float* __restrict__ gamma, const float* __restrict__ dres, int64_t ld_dres,
                   float* __restrict__ dx, int64_t ld_dx, RowMap dxmap, bf16* __restrict__ dx_bf16,
                   int64_t ld_dxb, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   float* __restrict__ colsum_out, int M, int D, int stages, uint32_t side_thr, float side_ks,
                   uint64_t side_seed_host, uint32_t side_stream, const uint64_t* side_epoch,
                   const uint8_t* __restrict__ side_bits) {
  extern __shared__ __align__(128) uint8_t lnp_smem[];
  __shared__ __align__(16) float red[2][2 * LNP_ROWS][LNB_MAX_WARPS];
  __shared__ __align__(8) uint64_t full[LNP_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty[LNP_MAX_STAGES];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int ncompute = blockDim.x - 32;  // compute threads (a multiple of 32)
  const bool producer = (int)threadIdx.x >= ncompute;
  const int c = threadIdx.x;
  const bool active = c < (D >> 2);
  const float inv_d = 1.0f / (float)D;
  const uint32_t dy_row_bytes = (uint32_t)D * sizeof(DyT), f_row_bytes = (uint32_t)D * 4u;
  const uint32_t row_bytes = dy_row_bytes + f_row_bytes + (dres ? f_row_bytes : 0u);
  const uint32_t stage_bytes = row_bytes * LNP_ROWS;
  // stage layout: [dy rows | x rows | dres rows]
  auto s_dy = [&](int s, int k) { return reinterpret_cast<const DyT*>(lnp_smem + s * stage_bytes + k * dy_row_bytes); };
  auto s_x = [&](int s, int k) {
    return reinterpret_cast<const float*>(lnp_smem + s * stage_bytes + LNP_ROWS * dy_row_bytes + k * f_row_bytes);
  };
  auto s_res = [&](int s, int k) {
    return reinterpret_cast<const float*>(lnp_smem + s * stage_bytes + LNP_ROWS * (dy_row_bytes + f_row_bytes) + k * f_row_bytes);
  };
  const int iters = (M + gridDim.x * LNP_ROWS - 1) / (gridDim.x * LNP_ROWS);
  for (int i = threadIdx.x; i < 2 * 2 * LNP_ROWS * LNB_MAX_WARPS; i += blockDim.x) (&red[0][0][0])[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ncompute >> 5); }
    fence_mbar_init();
  }
  __syncthreads();

  if (producer) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int r0 = (it * gridDim.x + blockIdx.x) * LNP_ROWS;
        if (r0 >= M) break;
        const int s = it % stages;
        if (it >= stages) mbar_wait(&empty[s], ((it / stages) - 1) & 1);
        const int nr = min(LNP_ROWS, M - r0);
        mbar_arrive_expect_tx(&full[s], nr * row_bytes);
        for (int k = 0; k < nr; ++k) {
          const int r = r0 + k;
          bulk_load_1d(const_cast<DyT*>(s_dy(s, k)), dy + dymap(r) * ld_dy, dy_row_bytes, &full[s]);
          bulk_load_1d(const_cast<float*>(s_x(s, k)), x + xmap(r) * ld_x, f_row_bytes, &full[s]);
          if (dres) bulk_load_1d(const_cast<float*>(s_res(s, k)), dres + (int64_t)r * ld_dres, f_row_bytes, &full[s]);
        }
      }
    }
    return;
  }

  float4 acc_g = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = acc_g, acc_c = acc_g;
  const float4 g = active ? *reinterpret_cast<const float4*>(gamma + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  int buf = 0;
  float mean_n[LNP_ROWS], rstd_n[LNP_ROWS];  // statistics of the NEXT row block, fetched one iteration ahead
  uint32_t sbits_n = 0;  // pre-drawn side-car keep byte of the NEXT row block (this lane's row of the pair)
  auto fetch_stats = [&](int it) {
    const int r0 = (it * gridDim.x + blockIdx.x) * LNP_ROWS;
#pragma unroll
    for (int k = 0; k < LNP_ROWS; ++k) {
      const int r = min(r0 + k, M - 1);
      mean_n[k] = __ldg(mean_in + r);
      rstd_n[k] = __ldg(rstd_in + r);
    }
    if (side_bits != nullptr && active) {
      const int rmine = min(r0 + (c & 1), M - 1);
      sbits_n = __ldg(side_bits + (((uint64_t)dxmap(rmine) * D + 4 * (c & ~1)) >> 3));
    }
  };
  fetch_stats(0);
  // effective seed of the inline draw, read ONCE: inside the loop the multiply on the freshly loaded epoch stalled every
  // iteration on a global load (ncu: IMAD ... 0x7f4a7c15 on long_sb) — even when the bits are pre-drawn and no seed is used
  const uint64_t side_seed = (side_thr != 0 && side_bits == nullptr) ? nv_seed(side_seed_host, side_epoch) : 0ull;
  for (int it = 0; it < iters; ++it) {
    const int r0 = (it * gridDim.x + blockIdx.x) * LNP_ROWS;
    if (r0 >= M) break;  // uniform across the CTA
    const int s = it % stages;
    float mean[LNP_ROWS], rstd[LNP_ROWS];
#pragma unroll
    for (int k = 0; k < LNP_ROWS; ++k) { mean[k] = mean_n[k]; rstd[k] = rstd_n[k]; }
    const uint32_t sbits = sbits_n;
    if (it + 1 < iters) fetch_stats(it + 1);
    mbar_wait(&full[s], (it / stages) & 1);
    float4 xh[LNP_ROWS], gh[LNP_ROWS], rv[LNP_ROWS];
    float part[2 * LNP_ROWS];
#pragma unroll
    for (int k = 0; k < LNP_ROWS; ++k) {
      const bool ok = active && (r0 + k) < M;
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), dv = xv;
      rv[k] = xv;
      if (ok) {
        xv = *reinterpret_cast<const float4*>(s_x(s, k) + 4 * c);
        dv = ld4<DyT>(s_dy(s, k) + 4 * c);
        if (dres) rv[k] = *reinterpret_cast<const float4*>(s_res(s, k) + 4 * c);
      }
      xh[k].x = ok ? (xv.x - mean[k]) * rstd[k] : 0.f;
      xh[k].y = ok ? (xv.y - mean[k]) * rstd[k] : 0.f;
      xh[k].z = ok ? (xv.z - mean[k]) * rstd[k] : 0.f;
      xh[k].w = ok ? (xv.w - mean[k]) * rstd[k] : 0.f;
      acc_g.x += dv.x * xh[k].x; acc_g.y += dv.y * xh[k].y;
      acc_g.z += dv.z * xh[k].z; acc_g.w += dv.w * xh[k].w;
      acc_b.x += dv.x; acc_b.y += dv.y; acc_b.z += dv.z; acc_b.w += dv.w;
      gh[k].x = dv.x * g.x; gh[k].y = dv.y * g.y; gh[k].z = dv.z * g.z; gh[k].w = dv.w * g.w;
      part[2 * k] = (gh[k].x + gh[k].y) + (gh[k].z + gh[k].w);
      part[2 * k + 1] = (gh[k].x * xh[k].x + gh[k].y * xh[k].y) + (gh[k].z * xh[k].z + gh[k].w * xh[k].w);
    }
    // this warp is done reading the stage: hand it back to the producer
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
#pragma unroll
    for (int i = 0; i < 2 * LNP_ROWS; ++i) part[i] = warp_sum(part[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 2 * LNP_ROWS; ++i) red[buf][i][warp] = part[i];
    }
    asm volatile("bar.sync 1, %0;" ::"r"(ncompute) : "memory");  // compute warps only
#pragma unroll
    for (int i = 0; i < 2 * LNP_ROWS; ++i) {  // entries of warps that do not exist stay zero
      float sacc = 0.f;
#pragma unroll
      for (int w4 = 0; w4 < LNB_MAX_WARPS / 4; ++w4) {
        const float4 t = *reinterpret_cast<const float4*>(&red[buf][i][w4 * 4]);
        sacc += (t.x + t.y) + (t.z + t.w);
      }
      part[i] = sacc * inv_d;
    }
    buf ^= 1;
    // Side-car dropout: dx feeds (a) the residual stream, unmasked fp32, and (b) the backward of the previous
    // block's last linear, whose output went through nn.Dropout — (b)'s operand (the bf16 copy) and its bias
    // column sums take that layer's forward mask here, so no separate masking pass is needed. One Philox call
    // covers 8 columns = lanes (c, c ^ 1): each lane draws the bits of one of the two rows and swaps.
    uint32_t keep4[LNP_ROWS];
#pragma unroll
    for (int k = 0; k < LNP_ROWS; ++k) keep4[k] = 0xFu;
    if (side_thr != 0) {
      static_assert(LNP_ROWS == 2, "pair exchange below assumes two rows per iteration");
      const int rmine = min(r0 + (c & 1), M - 1);
      const uint64_t grp = ((uint64_t)dxmap(rmine) * D + 4 * (c & ~1)) >> 3;
      const uint32_t mine = side_bits ? sbits  // drawn ahead (nv_dropout_bits), fetched one iteration early
                                      : nv_keep_bits8(side_seed, grp, side_stream, side_thr);
      const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
#pragma unroll
      for (int k = 0; k < LNP_ROWS; ++k) keep4[k] = ((((c & 1) == k) ? mine : other) >> ((c & 1) * 4)) & 0xFu;
    }
#pragma unroll
    for (int k = 0; k < LNP_ROWS; ++k) {
      const int r = r0 + k;
      if (!(active && r < M)) continue;
      const float s1 = part[2 * k], s2 = part[2 * k + 1];
      float4 o;
      o.x = rstd[k] * (gh[k].x - s1 - xh[k].x * s2) + rv[k].x;
      o.y = rstd[k] * (gh[k].y - s1 - xh[k].y * s2) + rv[k].y;
      o.z = rstd[k] * (gh[k].z - s1 - xh[k].z * s2) + rv[k].z;
      o.w = rstd[k] * (gh[k].w - s1 - xh[k].w * s2) + rv[k].w;
      const int64_t orow = dxmap(r);
      if (dx) *reinterpret_cast<float4*>(dx + orow * ld_dx + 4 * c) = o;
      if (side_thr != 0) o = nv_dropout4(o, keep4[k], side_ks);
      acc_c.x += o.x; acc_c.y += o.y; acc_c.z += o.z; acc_c.w += o.w;
      if (dx_bf16) OutStore<bf16>::st(dx_bf16 + orow * ld_dxb + 4 * c, o);
    }
  }
  if (active) {
    if (dgamma)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dgamma + 4 * c), "f"(acc_g.x), "f"(acc_g.y),
                   "f"(acc_g.z), "f"(acc_g.w) : "memory");
    if (dbeta)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbeta + 4 * c), "f"(acc_b.x), "f"(acc_b.y),
                   "f"(acc_b.z), "f"(acc_b.w) : "memory");
    if (colsum_out)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum_out + 4 * c), "f"(acc_c.x),
                   "f"(acc_c.y), "f"(acc_c.z), "f"(acc_c.w) : "memory");
  }
}

template <typename DyT>
int launch_ln_bwd_pipe(const DyT* dy, int64_t ld_dy, RowMap dym, const float* x, int64_t ld_x, RowMap xm,
                       const float* mean, const float* rstd, const float* gamma, const float* dres, int64_t ld_dres,
                       float* dx, int64_t ld_dx, RowMap dxm, bf16* dx_bf16, int64_t ld_dxb, float* dgamma, float* dbeta,
                       float* colsum, int M, int D, int threads, uint32_t side_thr, uint64_t side_seed, int side_stream,
                       const uint8_t* side_bits, cudaStream_t stream) {
  const int stage_bytes = (D * (int)sizeof(DyT) + D * 4 + (dres ? D * 4 : 0)) * LNP_ROWS;
  int stages = (108 * 1024) / stage_bytes;  // two CTAs per SM
  if (stages > LNP_MAX_STAGES) stages = LNP_MAX_STAGES;
  if (stages < 3) stages = 3;
  const int smem = stage_bytes * stages;
  auto kern = threads <= 256 ? ln_bwd_pipe_kernel<DyT, 256> : ln_bwd_pipe_kernel<DyT, 512>;
  static int smem_set[64][2] = {};  // largest opt-in so far, per device (function attributes are per device)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  int& set = smem_set[dev][threads <= 256 ? 0 : 1];
  if (smem > set) {
    NV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    set = smem;
  }
  int grid = (M + LNP_ROWS - 1) / LNP_ROWS;
  int cap = nv_num_sms() * ((smem <= 108 * 1024 && threads <= 256) ? 2 : 1);
  if (grid > cap) grid = cap;
  kern<<<grid, threads + 32, smem, stream>>>(dy, ld_dy, dym, x, ld_x, xm, mean, rstd, gamma, dres, ld_dres, dx, ld_dx,
                                             dxm, dx_bf16, ld_dxb, dgamma, dbeta, colsum, M, D, stages, side_thr,
                                             nv_dropout_keep_scale(side_thr), side_seed, (uint32_t)side_stream,
                                             side_thr != 0 ? nv_rng_epoch_dev() : nullptr, side_thr != 0 ? side_bits : nullptr);
  return NV_OK;
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

int nv_ln_bwd_launch(const void* dy, int dy_is_bf16, int64_t ld_dy, int dyg, int dys, int dyo, const float* x, int64_t ld_x,
                     int xg, int xs, int xo, const float* mean, const float* rstd, const float* gamma,
                     const float* dres, int64_t ld_dres, float* dx, int64_t ld_dx, int dxg, int dxs, int dxo,
                     bf16* dx_bf16, int64_t ld_dxb, float* dgamma, float* dbeta, float* colsum, int M, int D,
                     float side_drop_p, uint64_t side_drop_seed, int side_drop_stream, const uint8_t* side_drop_bits,
                     cudaStream_t stream) {
  NV_REQUIRE(M >= 0 && D > 0 && D % 4 == 0 && D <= 2048, "layernorm bwd: D=%d must be a multiple of 4 and <= 2048", D);
  if (M == 0) return NV_OK;
  for (const float* p : {dgamma, dbeta, colsum})
    NV_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "layernorm bwd: dgamma/dbeta/colsum must be 16-byte aligned");
  RowMap dym{dyg, dys, dyo}, xm{xg, xs, xo}, dxm{dxg, dxs, dxo};
  const int threads = ((D / 4 + 31) / 32) * 32;
  // bulk-copy pipelined kernel when every row is a 16-byte-aligned multiple of 16 bytes and the problem is
  // big enough to be bandwidth-bound; the register-staged kernel covers the rest (tiny M, odd D / strides)
  const int dy_es = dy_is_bf16 ? 2 : 4;
  const bool pipe_ok = M >= 64 && (D * dy_es) % 16 == 0 && (ld_dy * dy_es) % 16 == 0 && ld_x % 4 == 0 &&
                       ld_dres % 4 == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dres) & 15) == 0 &&
                       (int64_t)D * (dy_es + 8) * LNP_ROWS * 3 <= 200 * 1024;
  NV_REQUIRE(side_drop_p >= 0.f && side_drop_p < 1.f, "layernorm bwd: side_drop_p %f out of range [0, 1)", side_drop_p);
  const uint32_t side_thr = nv_dropout_threshold(side_drop_p);
  NV_REQUIRE(side_thr == 0 || (pipe_ok && D % 8 == 0),
             "layernorm bwd: the side-car dropout needs the pipelined kernel (M >= 64, D %% 8 == 0, 16-byte aligned rows)");
  if (pipe_ok) {
    int s;
    if (dy_is_bf16)
      s = launch_ln_bwd_pipe<bf16>((const bf16*)dy, ld_dy, dym, x, ld_x, xm, mean, rstd, gamma, dres, ld_dres, dx, ld_dx,
                                   dxm, dx_bf16, ld_dxb, dgamma, dbeta, colsum, M, D, threads, side_thr, side_drop_seed,
                                   side_drop_stream, side_drop_bits, stream);
    else
      s = launch_ln_bwd_pipe<float>((const float*)dy, ld_dy, dym, x, ld_x, xm, mean, rstd, gamma, dres, ld_dres, dx,
                                    ld_dx, dxm, dx_bf16, ld_dxb, dgamma, dbeta, colsum, M, D, threads, side_thr,
                                    side_drop_seed, side_drop_stream, side_drop_bits, stream);
    if (s != NV_OK) return s;
    NV_LAUNCH_CHECK("ln_bwd_pipe_kernel");
    return NV_OK;
  }
  int ctas_per_sm = 1;  // size the persistent grid to exactly one resident wave
  if (dy_is_bf16) NV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, ln_bwd_kernel<bf16>, threads, 0));
  else NV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, ln_bwd_kernel<float>, threads, 0));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  int grid = (M + LNB_ROWS - 1) / LNB_ROWS;
  const int cap = nv_num_sms() * ctas_per_sm;
  if (grid > cap) grid = cap;
  if (dy_is_bf16) {
    NV_REQUIRE(ld_dy % 4 == 0 && (reinterpret_cast<uintptr_t>(dy) & 7) == 0, "layernorm bwd: bf16 dy must be 8-byte aligned");
    ln_bwd_kernel<bf16><<<grid, threads, 0, stream>>>((const bf16*)dy, ld_dy, dym, x, ld_x, xm, mean, rstd, gamma, dres,
                                                      ld_dres, dx, ld_dx, dxm, dx_bf16, ld_dxb, dgamma, dbeta, colsum, M, D);
  } else {
    ln_bwd_kernel<float><<<grid, threads, 0, stream>>>((const float*)dy, ld_dy, dym, x, ld_x, xm, mean, rstd, gamma, dres,
                                                       ld_dres, dx, ld_dx, dxm, dx_bf16, ld_dxb, dgamma, dbeta, colsum, M, D);
  }
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
