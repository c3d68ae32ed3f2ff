"""TEST INFRASTRUCTURE — numpy restatement of the dropout keep-bit contract of neurovit_b200/csrc/nv_rng.cuh.

The reference draws its dropout masks from torch's generator (nn.Dropout, src/models/vit_3d.py:21,23,39,45,100); a
fused kernel cannot replay that stream, so the library defines its own counter-based contract and this
file states it in the plainest possible form, with no bit-slicing:

    effective seed  s = seed + epoch * 0x9E3779B97F4A7C15              (mod 2^64; nv_seed)
    one PAIR of groups (2c, 2c + 1), 16 consecutive elements  <-  Philox4x32-7(key = s, counter = (c, stream, 0x2B992DDF))
    the call's 16 output bytes (little-endian over x, y, z, w) are bit-planes of 8 lanes (lane j = bit j of a byte):
    H0 = bytes 0..7, H1 = bytes 8..15, each one byte VALUE per lane (plane q of a half = bit q of that value);
        u_j(group 2c + s) = 256 * H_s[j] + bitrev8(H_{1-s}[j])          j = 0..7   (a 16-bit uniform per element)
    element 8 g + j is KEPT iff u_j >= thr,  thr = round(p * 65536);  survivors are scaled by 65536 / (65536 - thr)
(marginals exactly uniform; lane j of the two groups of a pair reuse each other's high byte as low byte — see the
dependence note in nv_rng.cuh and test_keep_rate_and_independence_of_lanes.)

The CUDA side evaluates "u_j >= thr" bit-sliced (one LOP3 per plane for all lanes of a register); here the uniforms
are assembled explicitly. Pinned two ways: tests/test_oracle.py checks the generator against Random123's published known-answer vectors
(philox4x32, 7 and 10 rounds) and the vectorised form against the scalar one; tests/test_gpu_kernels.py checks
nv_dropout_bits and the inline draws of nv_dropout against it bit for bit. Only tests/ import this module.
"""
from __future__ import annotations

import numpy as np

M32 = np.uint64(0xFFFFFFFF)
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
COUNTER_HI = 0x2B992DDF
EPOCH_MUL = 0x9E3779B97F4A7C15
ROUNDS = 7


def effective_seed(seed: int, epoch: int = 0) -> int:
    """nv_seed (nv_rng.cuh): seed + epoch * c mod 2^64."""
    return (int(seed) + int(epoch) * EPOCH_MUL) & 0xFFFFFFFFFFFFFFFF


def philox4x32(seed: int, idx, stream: int, rounds: int = ROUNDS):
    """Philox4x32 (Salmon et al., SC'11) with `rounds` rounds; key = the 64-bit seed, counter = (idx lo, idx hi,
    stream, COUNTER_HI). idx: uint64 array. Returns four uint32 arrays (philox4x32_7 of nv_rng.cuh)."""
    idx = np.asarray(idx, dtype=np.uint64)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    c0 = idx & M32
    c1 = idx >> np.uint64(32)
    c2 = np.full_like(idx, np.uint64(stream & 0xFFFFFFFF))
    c3 = np.full_like(idx, np.uint64(COUNTER_HI))
    for _ in range(rounds):
        p0 = PHILOX_M0 * c0          # 32 x 32 -> 64 bit products (operands < 2^32: no overflow in uint64)
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(PHILOX_W0)) & M32
        k1 = (k1 + np.uint64(PHILOX_W1)) & M32
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def threshold(p: float) -> int:
    """nv_dropout_threshold."""
    if not p > 0:
        return 0
    return min(int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)), 65535)


def keep_scale(thr: int) -> float:
    return 1.0 if thr == 0 else 65536.0 / (65536 - thr)


def uniforms16(seed: int, groups, stream: int) -> np.ndarray:
    """[len(groups), 8] uint32: the eight 16-bit uniforms of each group (byte values assembled explicitly)."""
    groups = np.asarray(groups, dtype=np.uint64)
    words = philox4x32(seed, groups >> np.uint64(1), stream)
    byte = np.stack([(words[q >> 2] >> np.uint32(8 * (q & 3))) & np.uint32(0xFF) for q in range(16)], axis=1)  # [n, 16]
    lanes = np.arange(8, dtype=np.uint32)

    def lane_values(planes, reverse):   # planes [n, 8] (byte q = plane q) -> per-lane byte value [n, 8 lanes]
        v = np.zeros((planes.shape[0], 8), dtype=np.uint32)
        for q in range(8):
            bit = (planes[:, q, None] >> lanes[None, :]) & np.uint32(1)
            v |= bit << np.uint32(7 - q if reverse else q)
        return v

    odd = (groups & np.uint64(1)).astype(bool)[:, None]
    h0, h1 = byte[:, :8], byte[:, 8:]
    hi = np.where(odd, h1, h0)
    lo = np.where(odd, h0, h1)
    return (lane_values(hi, False) << np.uint32(8)) | lane_values(lo, True)


def keep_bits8(seed: int, groups, stream: int, thr: int) -> np.ndarray:
    """uint8 per group: bit j set = element 8 g + j survives (nv_keep_bits8)."""
    groups = np.asarray(groups, dtype=np.uint64)
    keep = uniforms16(seed, groups, stream) >= np.uint32(thr)
    return (keep.astype(np.uint32) << np.arange(8, dtype=np.uint32)[None, :]).sum(axis=1).astype(np.uint8)


def dropout_bits(n_groups: int, p: float, seed: int, stream: int, epoch: int = 0) -> np.ndarray:
    """What nv_dropout_bits writes: byte g = keep bits of elements [8 g, 8 g + 8) of the site (seed, stream)."""
    return keep_bits8(effective_seed(seed, epoch), np.arange(n_groups, dtype=np.uint64), stream, threshold(p))


def keep_mask(M: int, N: int, p: float, seed: int, stream: int, epoch: int = 0, row_mul: int = 1) -> np.ndarray:
    """[M, N] bool keep mask of an element-indexed site (GEMM epilogues, nv_dropout): element index = row * row_mul * N
    + col; N % 8 == 0."""
    assert N % 8 == 0
    rows = np.arange(M, dtype=np.uint64)[:, None] * np.uint64(row_mul * N)
    groups = ((rows + np.arange(0, N, 8, dtype=np.uint64)[None, :]) >> np.uint64(3)).reshape(-1)
    bits = keep_bits8(effective_seed(seed, epoch), groups, stream, threshold(p))
    return ((bits[:, None] >> np.arange(8, dtype=np.uint8)[None, :]) & 1).astype(bool).reshape(M, N)


def philox4x32_raw(counter, key, rounds: int):
    """Scalar Philox4x32 on Python integers with an arbitrary counter / key: the form Random123's known-answer vectors
    (kat_vectors: philox4x32 7 / 10) are stated in. tests/test_oracle.py checks those vectors."""
    c = [int(v) & 0xFFFFFFFF for v in counter]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(rounds):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k1, p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + PHILOX_W0) & 0xFFFFFFFF, (k1 + PHILOX_W1) & 0xFFFFFFFF
    return c


def philox_scalar(seed: int, idx: int, stream: int, rounds: int = ROUNDS):
    """The library's counter / key layout through the scalar form (cross-checks the vectorised philox4x32)."""
    return philox4x32_raw([idx & 0xFFFFFFFF, (idx >> 32) & 0xFFFFFFFF, stream, COUNTER_HI],
                          [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], rounds)


def bitsliced_keep_bits8(words, thr: int, odd: bool = False) -> int:
    """The CUDA evaluation order (nv_keep_bits8): least-significant plane first, lt <- t_q ? (lt | ~B) : (lt & ~B);
    low planes = the other half's bytes last first, high planes = the group's own half. words: the call's four 32-bit
    outputs. The returned byte must equal the explicit compare — tests assert that."""
    hi = words[2:] if odd else words[:2]
    lo = words[:2] if odd else words[2:]
    lt = 0
    for q in range(16):
        if q < 8:
            b = 7 - q
            B = (lo[b >> 2] >> (8 * (b & 3))) & 0xFFFFFFFF
        else:
            b = q - 8
            B = (hi[b >> 2] >> (8 * (b & 3))) & 0xFFFFFFFF
        lt = (lt | ~B) if (thr >> q) & 1 else (lt & ~B)
        lt &= 0xFFFFFFFF
    return ~lt & 0xFF
