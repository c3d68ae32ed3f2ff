// Counter-based random bits for the dropout sites of the hot path (reference: nn.Dropout at
// src/models/vit_3d.py:21,23,39,45,100 — 25 sites per forward).
//
// torch's Philox stream cannot be reproduced bit-for-bit by a fused kernel, so the contract
// here is statistical: every element is dropped independently with probability p_eff = thr / 65536
// (thr = round(p * 65536)), survivors are scaled by 1 / (1 - p_eff), and the SAME bits are seen by forward
// and backward because they are a pure function of (seed, stream, element index) — nothing is stored
// except where a kernel's transposed walk would make regeneration expensive (attention, which saves a
// 1-bit-per-score mask instead).
//
// Generator: Philox4x32 with 7 rounds (Salmon et al., SC'11: the smallest round count that passes
// BigCrush); one call yields eight 16-bit uniforms.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NV_RNG_HD __host__ __device__ __forceinline__
#else
#define NV_RNG_HD inline
#endif

// drop when a 16-bit uniform is < threshold; 0 disables dropout
NV_RNG_HD uint32_t nv_dropout_threshold(float p) {
  if (!(p > 0.f)) return 0u;
  uint32_t t = (uint32_t)(p * 65536.0f + 0.5f);
  return t > 65535u ? 65535u : t;
}
NV_RNG_HD float nv_dropout_keep_scale(uint32_t thr) { return thr == 0 ? 1.0f : 65536.0f / (float)(65536u - thr); }

#ifdef __CUDACC__
// key = seed (64 bit), counter = (idx: 64 bit, stream: 32 bit, 0)
__device__ __forceinline__ uint4 philox4x32_7(uint64_t seed, uint64_t idx, uint32_t stream) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = stream, c3 = 0x2B992DDFu;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-bits (bit i set = element i survives) of the 8 consecutive elements [8*group, 8*group + 8)
__device__ __forceinline__ uint32_t nv_keep_bits8(uint64_t seed, uint64_t group, uint32_t stream, uint32_t thr) {
  const uint4 r = philox4x32_7(seed, group, stream);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m |= ((w[i] & 0xFFFFu) >= thr ? 1u : 0u) << (2 * i);
    m |= ((w[i] >> 16) >= thr ? 1u : 0u) << (2 * i + 1);
  }
  return m;
}
// Effective seed of a launch: the host-drawn seed plus a device-resident epoch counter (nv_rng_epoch_*), so a
// CUDA graph that bakes the host seeds in still draws fresh masks on every replay — the captured step ends
// with nv_rng_epoch_advance. Forward and backward of one step see the same epoch.
__device__ __forceinline__ uint64_t nv_seed(uint64_t seed, const uint64_t* epoch) {
  return epoch ? seed + __ldg(epoch) * 0x9E3779B97F4A7C15ull : seed;
}
// keep-bits of the 4 consecutive elements [idx, idx + 4), idx a multiple of 4 (element-indexed dropout sites:
// GEMM epilogues and the element-wise kernel share this so forward and backward regenerate the same mask)
__device__ __forceinline__ uint32_t nv_keep_bits4(uint64_t seed, uint64_t idx, uint32_t stream, uint32_t thr) {
  return (nv_keep_bits8(seed, idx >> 3, stream, thr) >> ((idx & 4) ? 4 : 0)) & 0xFu;
}
__device__ __forceinline__ float4 nv_dropout4(float4 v, uint32_t bits, float ks) {
  v.x = (bits & 1u) ? v.x * ks : 0.f;
  v.y = (bits & 2u) ? v.y * ks : 0.f;
  v.z = (bits & 4u) ? v.z * ks : 0.f;
  v.w = (bits & 8u) ? v.w * ks : 0.f;
  return v;
}
#endif
