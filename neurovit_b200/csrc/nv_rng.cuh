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
// BigCrush); one call (128 bits) decides the eight elements of one group. The 128 bits are read as SIXTEEN
// BIT-PLANES OF EIGHT LANES: byte q of the output (q = 0 the low byte of the first word) holds bit q of the eight
// 16-bit uniforms u_0..u_7, lane j = bit j of the byte. "u_j >= thr" is then evaluated bit-sliced, for all lanes of a
// register at once, by the least-significant-first recurrence
//     lt <- t_q ? (lt | ~B_q) : (lt & ~B_q)        (t_q = bit q of thr; one LOP3 per plane)
// after which bit j of lt says u_j < thr. Unpacking eight halfwords and comparing each (ISETP + SEL + shift + OR per
// element) was 55 % of the mask generator's instructions, all on the half-rate integer pipe; nv_keep_bits32 packs four
// calls into the four bytes of a register (byte transposes) and runs the sixteen planes once for 32 elements.
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
  const uint4 r = philox4x32_7(seed, group, stream);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t lt = 0;   // only the low byte is meaningful: the planes' neighbours ride along in the upper bits
#pragma unroll
  for (int q = 0; q < 16; ++q) lt = nv_lt_step(lt, w[q >> 2] >> (8 * (q & 3)), nv_thr_plane(thr, q));
  return ~lt & 0xFFu;
}
// keep-bits of the 32 consecutive elements [32*word, 32*word + 32): byte k = nv_keep_bits8(4*word + k). The four calls'
// outputs are byte-transposed so that byte k of every plane register belongs to call k, and the sixteen planes run once.
__device__ __forceinline__ uint32_t nv_keep_bits32(uint64_t seed, uint64_t word, uint32_t stream, uint32_t thr) {
  uint4 r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) r[k] = philox4x32_7(seed, 4 * word + k, stream);
  uint32_t lt = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t a = i == 0 ? r[0].x : i == 1 ? r[0].y : i == 2 ? r[0].z : r[0].w;
    const uint32_t b = i == 0 ? r[1].x : i == 1 ? r[1].y : i == 2 ? r[1].z : r[1].w;
    const uint32_t c = i == 0 ? r[2].x : i == 1 ? r[2].y : i == 2 ? r[2].z : r[2].w;
    const uint32_t d = i == 0 ? r[3].x : i == 1 ? r[3].y : i == 2 ? r[3].z : r[3].w;
    const uint32_t ab_lo = __byte_perm(a, b, 0x5140), ab_hi = __byte_perm(a, b, 0x7362);   // [a0 b0 a1 b1], [a2 b2 a3 b3]
    const uint32_t cd_lo = __byte_perm(c, d, 0x5140), cd_hi = __byte_perm(c, d, 0x7362);
    const uint32_t t[4] = {__byte_perm(ab_lo, cd_lo, 0x5410), __byte_perm(ab_lo, cd_lo, 0x7632),   // [a_b b_b c_b d_b]
                           __byte_perm(ab_hi, cd_hi, 0x5410), __byte_perm(ab_hi, cd_hi, 0x7632)};
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) lt = nv_lt_step(lt, t[bb], nv_thr_plane(thr, 4 * i + bb));
  }
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
