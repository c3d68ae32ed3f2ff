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
// BigCrush); one call (128 bits) decides the SIXTEEN elements of a pair of 8-element groups (2c, 2c + 1). The call's
// 16 output bytes are bit-planes of eight lanes (lane j = bit j of a byte): bytes 0..7 form the half H0, bytes 8..15
// the half H1, each eight planes = one byte value per lane. Group 2c + s takes H_s as the HIGH byte of its lanes'
// 16-bit uniforms and the other half, planes in reversed order (the byte value bit-reversed), as the LOW byte:
//     u_j(group 2c + s) = 256 * H_s[j] + bitrev8(H_{1-s}[j]).
// (H0[j], H1[j]) is uniform on 256 x 256 and the map is a bijection, so every u_j is exactly uniform on [0, 65536) and
// p_eff = thr / 65536 holds exactly. Lane j of the two groups of a pair (elements 8 apart) reuse each other's high byte
// as low byte, so their keep decisions differ from independent ones only through the event "high byte == high byte of
// thr" (probability 2^-8): correlation 4.3e-4 at p = 0.1, <= 7e-4 for 0.05 <= p <= 0.95, <= 4e-3 for 0.01 <= p <= 0.99
// (enumerated over all 65536 byte pairs; for p < 2^-8 the pair's rare drops tend to coincide — marginals stay exact);
// all other pairs of elements are independent. This halves the Philox calls per element — the generator's
// 32 x 32 -> 64 bit multiplies are its bottleneck — at full 16-bit resolution of p.
// "u_j >= thr" is evaluated bit-sliced, for all lanes of a register at once, by the least-significant-first recurrence
//     lt <- t_q ? (lt | ~B_q) : (lt & ~B_q)        (t_q = bit q of thr, B_q = plane q; one LOP3 per plane)
// after which bit j of lt says u_j < thr. Unpacking halfwords and comparing each (ISETP + SEL + shift + OR per element)
// was 55 % of the mask generator's instructions, all on the half-rate integer pipe; nv_keep_bits32 packs the two calls
// of four groups into the four bytes of a register (byte transposes) and runs the sixteen planes once for 32 elements.
// oracle/rng_oracle.py restates this contract in numpy; tests pin the kernels to it bit for bit.
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
// threshold bit q as an all-ones / all-zeros select mask
__device__ __forceinline__ uint32_t nv_thr_plane(uint32_t thr, int q) { return 0u - ((thr >> q) & 1u); }
// one step of the bit-sliced "u < thr" recurrence over plane B (least significant plane first)
// = majority(lt, ~B, T); written as ONE lop3 (immLut 0xB2 = f(0xF0, 0xCC, 0xAA)): nvcc otherwise emits three
__device__ __forceinline__ uint32_t nv_lt_step(uint32_t lt, uint32_t B, uint32_t T) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xB2;" : "=r"(r) : "r"(lt), "r"(B), "r"(T));
  return r;
}
// keep-bits (bit i set = element i survives) of the 8 consecutive elements [8*group, 8*group + 8)
__device__ __forceinline__ uint32_t nv_keep_bits8(uint64_t seed, uint64_t group, uint32_t stream, uint32_t thr) {
  const uint4 r = philox4x32_7(seed, group >> 1, stream);
  const bool s = group & 1;
  const uint32_t hi[2] = {s ? r.z : r.x, s ? r.w : r.y}, lo[2] = {s ? r.x : r.z, s ? r.y : r.w};
  uint32_t lt = 0;   // only the low byte is meaningful: the planes' neighbours ride along in the upper bits
#pragma unroll
  for (int q = 0; q < 8; ++q) {   // low byte of u: the other half's planes, last first
    const int b = 7 - q;
    lt = nv_lt_step(lt, lo[b >> 2] >> (8 * (b & 3)), nv_thr_plane(thr, q));
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) lt = nv_lt_step(lt, hi[q >> 2] >> (8 * (q & 3)), nv_thr_plane(thr, 8 + q));
  return ~lt & 0xFFu;
}
// keep-bits of the 32 consecutive elements [32*word, 32*word + 32): byte k = nv_keep_bits8(4*word + k), from the two
// calls 2*word (groups 0, 1) and 2*word + 1 (groups 2, 3). R_q = [c0.byte(q), c0.byte(8+q), c1.byte(q), c1.byte(8+q)]
// is high plane 8 + q of the four groups; the same register with the bytes of each pair swapped is low plane 7 - q, so
// the low planes run on the R_q as they are and the comparison state is pair-swapped once in between.
__device__ __forceinline__ uint32_t nv_keep_bits32(uint64_t seed, uint64_t word, uint32_t stream, uint32_t thr) {
  const uint4 c0 = philox4x32_7(seed, 2 * word, stream), c1 = philox4x32_7(seed, 2 * word + 1, stream);
  uint32_t R[8];
#pragma unroll
  for (int i = 0; i < 2; ++i) {   // bytes 4i .. 4i+3 of H0 (a, c) and of H1 (b, d): 4 x 4 byte transpose
    const uint32_t a = i == 0 ? c0.x : c0.y, b = i == 0 ? c0.z : c0.w, c = i == 0 ? c1.x : c1.y, d = i == 0 ? c1.z : c1.w;
    const uint32_t ab_lo = __byte_perm(a, b, 0x5140), ab_hi = __byte_perm(a, b, 0x7362);   // [a0 b0 a1 b1], [a2 b2 a3 b3]
    const uint32_t cd_lo = __byte_perm(c, d, 0x5140), cd_hi = __byte_perm(c, d, 0x7362);
    R[4 * i + 0] = __byte_perm(ab_lo, cd_lo, 0x5410);   // [a_q b_q c_q d_q]
    R[4 * i + 1] = __byte_perm(ab_lo, cd_lo, 0x7632);
    R[4 * i + 2] = __byte_perm(ab_hi, cd_hi, 0x5410);
    R[4 * i + 3] = __byte_perm(ab_hi, cd_hi, 0x7632);
  }
  uint32_t lt = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) lt = nv_lt_step(lt, R[7 - q], nv_thr_plane(thr, q));   // bytes in order [g1 g0 g3 g2]
  lt = __byte_perm(lt, 0, 0x2301);                                                   // -> [g0 g1 g2 g3]
#pragma unroll
  for (int q = 0; q < 8; ++q) lt = nv_lt_step(lt, R[q], nv_thr_plane(thr, 8 + q));
  return ~lt;
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
