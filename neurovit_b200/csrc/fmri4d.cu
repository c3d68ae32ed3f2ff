// 4D input pipeline of the NeuroEncoder path: the time axis leaves the innermost position.
// Reference: src/models/NeuroEncoder.py:54-56 — fmri.permute(0, 4, 1, 2, 3) followed by reshape(B*T, H, W, D), a
// strided copy of the whole sample (110 MB at 64x64x48x140) that ATen runs as a generic gather; and
// src/data/DatasetADNI_4D.py:84-86 — the per-sample z-score (x - mean) / (std + 1e-8) over all H*W*D*T values that
// the dataset applies just before. Here both are ONE pass at HBM rate:
//   y[b, t, s] = (x[b, s, t] - mean_b) * inv_b        s = flattened (H, W, D), inv_b = 1 / (std_b + eps)
// (plain de-interleave, bit-exact with permute().reshape(), when no statistics are given).
//
// A CTA owns FM_S consecutive s of one sample for ALL t of a chunk: its source is one contiguous run of FM_S*T floats
// (fully coalesced), staged in shared memory with an odd row pitch, and written as T rows of FM_S
// consecutive floats (256-byte segments). Algorithmic traffic: 4 B read + 4 B written per element.
// Statistics: per-sample sum and sum of squares accumulated in fp64 (numpy's mean/std of the dataset run in fp64 on
// the fp64 array nibabel delivers), one atomicAdd pair per CTA.
#include "nv_common.cuh"

namespace {

constexpr int FM_S = 64;        // s values per CTA
constexpr int FM_TC = 256;      // t values per chunk (T <= FM_TC: one chunk, the source run is contiguous)
constexpr int FM_THREADS = 256;

// stats[b] = {sum, sumsq} (fp64), already complete when the de-interleave kernel runs
__global__ void __launch_bounds__(FM_THREADS)
fmri_deinterleave_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t S, int T, const double* stats,
                         double count, double eps) {
  extern __shared__ float tile[];  // [FM_S][tc + 1]
  const int b = blockIdx.y;
  const int64_t s0 = (int64_t)blockIdx.x * FM_S;
  const int ns = (int)min((int64_t)FM_S, S - s0);
  const float* xb = x + (int64_t)b * S * T;
  float* yb = y + (int64_t)b * S * T;
  double mean = 0.0, inv = 1.0;
  const bool zs = stats != nullptr;
  if (zs) {
    const double sum = stats[2 * b], sq = stats[2 * b + 1];
    mean = sum / count;
    const double var = fmax(sq / count - mean * mean, 0.0);
    inv = 1.0 / (sqrt(var) + eps);
  }
  for (int t0 = 0; t0 < T; t0 += FM_TC) {
    const int tc = min(FM_TC, T - t0);
    const int pitch = tc | 1;  // odd: the transposed reads below hit 32 different banks
    if (tc == T && (T & 3) == 0 && ns == FM_S) {
      // the CTA's whole source run [s0 T, (s0 + FM_S) T) is contiguous and 16-byte aligned (s0 is a multiple of 64):
      // 16-byte loads, (row, column) advanced incrementally instead of divided out per element
      const float4* src = reinterpret_cast<const float4*>(xb + s0 * T);
      const int nvec = FM_S * T / 4;
      int e = 4 * threadIdx.x, r = e / T, c = e - r * T;   // T % 4 == 0: a vector never straddles two rows
      for (int i = threadIdx.x; i < nvec; i += FM_THREADS) {
        const float4 v = __ldg(src + i);
        float* d = tile + r * pitch + c;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        c += 4 * FM_THREADS;
        while (c >= T) { c -= T; ++r; }
      }
    } else {
      // ragged strip / T not a multiple of 4 / chunked T: element loads, still coalesced along t
      const int nel = ns * tc;
      int r = threadIdx.x / tc, c = threadIdx.x - r * tc;
      for (int i = threadIdx.x; i < nel; i += FM_THREADS) {
        tile[r * pitch + c] = __ldg(xb + (s0 + r) * T + t0 + c);
        c += FM_THREADS;
        while (c >= tc) { c -= tc; ++r; }
      }
    }
    __syncthreads();
    const int sl = threadIdx.x & (FM_S - 1);
    if (sl < ns) {
      const float* tp = tile + sl * pitch;
      float* yp = yb + (int64_t)(t0 + threadIdx.x / FM_S) * S + s0 + sl;
      const int64_t ystep = (int64_t)(FM_THREADS / FM_S) * S;
      for (int t = threadIdx.x / FM_S; t < tc; t += FM_THREADS / FM_S, yp += ystep) {
        float v = tp[t];
        if (zs) v = (float)(((double)v - mean) * inv);
        *yp = v;
      }
    }
    __syncthreads();
  }
}

// per-sample {sum, sumsq} in fp64; grid (chunks, B)
__global__ void __launch_bounds__(256)
fmri_moments_kernel(const float* __restrict__ x, int64_t per_sample, double* stats) {
  const int b = blockIdx.y;
  const float* xb = x + (int64_t)b * per_sample;
  double s = 0.0, q = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if ((per_sample & 3) == 0) {
    const float4* xv = reinterpret_cast<const float4*>(xb);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample / 4; i += stride) {
      const float4 v = __ldg(xv + i);
      const double a = v.x, c = v.y, d = v.z, e = v.w;
      s += (a + c) + (d + e);
      q += (a * a + c * c) + (d * d + e * e);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += stride) {
      const double a = __ldg(xb + i);
      s += a;
      q += a * a;
    }
  }
  __shared__ double rs[8], rq[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { s += rs[i]; q += rq[i]; }
    atomicAdd(stats + 2 * b, s);
    atomicAdd(stats + 2 * b + 1, q);
  }
}

}  // namespace

// x [B, S, T] fp32 contiguous -> y [B, T, S]; stats_ws: NULL (plain de-interleave) or B*2 doubles of workspace
// (zeroed here) for the per-sample z-score with `eps` added to the standard deviation.
int nv_fmri_deinterleave_launch(const float* x, float* y, int B, int64_t S, int T, double* stats_ws, double eps,
                                cudaStream_t stream) {
  NV_REQUIRE(x != nullptr && y != nullptr && x != y, "fmri_deinterleave: null or aliased buffers");
  NV_REQUIRE(B >= 1 && S >= 1 && T >= 1 && B <= 65535, "fmri_deinterleave: bad sizes B=%d S=%lld T=%d", B, (long long)S, T);
  if (stats_ws != nullptr) {
    NV_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * B, stream));
    const int64_t per = S * T;
    const int chunks = (int)min((int64_t)(2 * nv_num_sms()), (per / 4 + 255) / 256 + 1);
    fmri_moments_kernel<<<dim3(chunks, B), 256, 0, stream>>>(x, per, stats_ws);
    NV_LAUNCH_CHECK("fmri_moments_kernel");
  }
  const int tc = T < FM_TC ? T : FM_TC;
  const size_t smem = sizeof(float) * FM_S * (size_t)(tc | 1);
  static uint64_t attr_flags = 0;
  if (nv_first_on_device(&attr_flags))
    NV_CUDA(cudaFuncSetAttribute(fmri_deinterleave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(float) * FM_S * (FM_TC | 1))));
  const int64_t tiles = (S + FM_S - 1) / FM_S;
  NV_REQUIRE(tiles <= 2147483647LL, "fmri_deinterleave: volume too large");
  fmri_deinterleave_kernel<<<dim3((unsigned)tiles, B), FM_THREADS, smem, stream>>>(x, y, S, T, stats_ws, (double)S * T, eps);
  NV_LAUNCH_CHECK("fmri_deinterleave_kernel");
  return NV_OK;
}
