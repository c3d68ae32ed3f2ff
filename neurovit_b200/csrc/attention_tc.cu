// tcgen05 / TMEM flash attention, forward and backward, head_dim = 64, bf16 operands, fp32 statistics.
// Reference: src/models/vit_3d.py:51-59 (dots = q k^T * scale; softmax(dim=-1); dropout; attn v;
// 'b h n d -> b n (h d)'), SURVEY 8a row A8. No mask, not causal; [B,h,N,N] is never materialised.
//
// Token 0 (the cls token, vit_3d.py:116) is kept OUT of the tensor-core tiles: every ViT3D geometry has
// N = n + 1 tokens with n a product of grid sizes (384, 1000, 1728), so tiling N itself left a 1-row fourth query
// tile and a 1-key last block at N = 385 (a quarter of the CTAs nearly empty, a fifth of the key blocks one
// column wide). The tiles cover tokens 1..n; token 0 is
//   - as a KEY: one extra score per row, computed by the row's own thread from its Q (dO) row in shared memory
//     (64 FMAs) and folded into the row's softmax / dQ in registers (fwd, dQ); its dK / dV are a reduction over
//     all queries done by one SIMT CTA per (batch, head) at the end of the dK/dV grid;
//   - as a QUERY: one SIMT CTA per (batch, head) at the end of the forward and dQ grids (385 x 64 dot products),
//     and a rank-1 update added by each key's thread in the dK/dV epilogue.
// Dropout mask layout [B*H, N, ceil(N/32)] words: row = query token, bit position = key token - 1 for the tiled
// keys and N - 1 for key token 0 (so 32-key chunks of the tiles stay word-aligned).
//
// All three tile kernels share one skeleton: a CTA owns 128 rows (= the 128 TMEM lanes) of one (batch, head),
//   warps 0-3  "row" warps: thread i owns TMEM lane i, i.e. one query (fwd, dQ) or one key (dK/dV) row.
//              tcgen05.ld gives the thread its whole score row, so the softmax statistics (row max, row
//              sum, LSE, delta) live in that thread's registers with no cross-thread reduction at all;
//              probabilities go back to shared memory as a bf16 K-major 128B-swizzled MMA operand.
//   warp  4    TMA producer: 64-col bf16 boxes of q / k / v / dO read IN PLACE from the QKV GEMM
//              output through 3-D tensor maps (col, token, batch); rows past N are zero-filled by TMA.
//   warp  5    TMEM allocator + tcgen05.mma issuer.
// TMEM budget is 256 columns and shared memory <= 113 KB per CTA, so two CTAs are co-resident per SM:
// one CTA's exponentials (MUFU-bound at head_dim 64) overlap the other CTA's MMAs and TMA waits.
// Ragged n: row warps whose 32 rows are all >= n skip their work, the last column block shrinks to a multiple
// of 16.
//
//   fwd : per key block j (96 keys): S_j = Q K_j^T -> TMEM (double-buffered); the row thread pulls its 96 scores
//         into registers in one TMEM pass, p = 2^(S*c - m) with a lazily raised reference m (only when the block
//         maximum exceeds it by 2^8), P -> smem, O += P V_j accumulates in TMEM across blocks (rescaled in place
//         on the rare raises).
//   dQ  : prologue: delta_i = dO_i . O_i (written for the dK/dV kernel), the extra key's dS; per key block j
//         (64 keys): S = Q K_j^T (double-buffered, issued two blocks ahead through a 3-stage K/V ring), dP = dO V_j^T
//         (one block ahead) -> TMEM; the rows take all of the block's exponentials first, then dP:
//         dS = P o (dP - delta) -> smem; dQ += dS K_j in TMEM.
//   dKV : per query block i (64 queries): S^T = K Q_i^T, dP^T = V dO_i^T -> TMEM (lane = key), issued BEFORE the
//         previous block's accumulation; P^T, dS^T -> smem; dV += P^T dO_i, dK += dS^T Q_i accumulate in TMEM; epilogue
//         adds the cls query's rank-1 term.
//   cls-only forward (nv_attention_cls_fwd): the SIMT CTAs alone — the last block under a cls-pooled head needs the
//         attention output of token 0 only.
#include "nv_common.cuh"
#include "nv_rng.cuh"
#include <cstdlib>
#include <cstring>

// Optional in-kernel phase clocks (build with NV_PROFILE=1): selected threads accumulate clock64() deltas per
// phase into nv_attn_dbg, read back through nv_debug_read (not part of the public ABI).
#ifdef NV_PROFILE
__device__ long long nv_attn_dbg[512];
#define PROF_DECL long long pt_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long pt0_ = clock64(), pt1_;
#define PROF_MARK(i) do { pt1_ = clock64(); pt_[i] += pt1_ - pt0_; pt0_ = pt1_; } while (0)
#define PROF_RESET() do { pt0_ = clock64(); } while (0)
#define PROF_DUMP(slot) do { for (int i_ = 0; i_ < 12; ++i_) nv_attn_dbg[(slot) * 12 + i_] = pt_[i_]; } while (0)
extern "C" int nv_debug_read(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, nv_attn_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : 3;
}
#else
#define PROF_DECL
#define PROF_MARK(i)
#define PROF_RESET()
#define PROF_DUMP(slot)
#endif

// Optional progress markers for debugging a hang (build with NV_DEBUG_PROGRESS=1): threads store a code into a
// host-mapped buffer (nv_debug_set_progress_buffer) that the host can read while the kernel is still running.
#ifdef NV_DEBUG_PROGRESS
__device__ volatile int* nv_dbg_progress = nullptr;
extern "C" int nv_debug_set_progress_buffer(int* host_mapped) {
  return cudaMemcpyToSymbol(nv_dbg_progress, &host_mapped, sizeof(int*)) == cudaSuccess ? 0 : 3;
}
#define DBG_MARK(slot, code) do { if (nv_dbg_progress) { nv_dbg_progress[(slot)] = (code); __threadfence_system(); } } while (0)
#else
#define DBG_MARK(slot, code)
#endif

namespace {

constexpr int HD = 64;
constexpr int BQ = 128;               // rows per CTA (TMEM lanes)
constexpr int SLAB = BQ * 128;        // 16 KB: [128 rows x 64 bf16], K-major, 128B swizzle
constexpr int BOX = 64 * 128;         // 8 KB: one TMA box (64 rows x 64 bf16)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int NTHREADS = 192;
constexpr int TMA_WARP = 4, MMA_WARP = 5;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 32 consecutive bf16 of row `row` of a K-major SW128 operand tile, columns [32*chunk, 32*chunk+32)
__device__ __forceinline__ void store_operand_chunk(uint32_t tile_base, int row, int chunk, const uint32_t (&w)[16]) {
  const uint32_t slab = tile_base + (uint32_t)(chunk >> 1) * SLAB + (uint32_t)row * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = ((chunk & 1) << 2) + i;
    st_shared_v4(slab + (uint32_t)((ch ^ (row & 7)) << 4), w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  }
}
// rows [row0, row0 + 64*nbox) x 64 columns at column `col` of batch `b` -> consecutive 8 KB boxes
__device__ __forceinline__ void load_rows(uint8_t* dst, const CUtensorMap* m, uint64_t* bar, int col, int row0, int b,
                                          int nbox) {
  for (int i = 0; i < nbox; ++i) tma_load_3d(dst + i * BOX, m, bar, col, row0 + 64 * i, b);
}
__device__ __forceinline__ uint64_t kmajor_desc(const uint8_t* tile) { return umma_smem_desc_sw128(smem_u32(tile), 16, 1024); }
__device__ __forceinline__ uint64_t mnmajor_desc(const uint8_t* tile) { return umma_smem_desc_sw128(smem_u32(tile), SLAB, 1024); }
// D[128 x n] (+)= A[128 x 64 (one slab, K-major)] * B[n x 64 (K-major)]^T : 4 k-steps of 16
__device__ __forceinline__ void mma_k64(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0);
}
// D[128 x 64] (+)= A[128 x 16*ksteps (K-major slabs of 64)] * B[16*ksteps x 64 (MN-major: rows = k)]
__device__ __forceinline__ void mma_rows(uint32_t d_tmem, const uint8_t* a_tile, const uint8_t* b_tile, uint32_t idesc,
                                         int ksteps, bool accumulate) {
  const uint64_t a0 = kmajor_desc(a_tile), b0 = mnmajor_desc(b_tile);
  for (int k = 0; k < ksteps; ++k)
    umma_f16_ss(d_tmem, a0 + (uint64_t)((k >> 2) * (SLAB >> 4) + (k & 3) * 2), b0 + (uint64_t)(k * (2048 >> 4)), idesc,
                (accumulate || k > 0) ? 1u : 0u);
}

struct Common {
  int N, H, B;        // N tokens per sample INCLUDING token 0; the tiles cover tokens 1 .. N-1
  float scale;        // dim_head^-0.5
  // dropout on the attention probabilities (vit_3d.py:56). Forward draws the keep bits (Philox, keyed by
  // seed and (batch*head, query, position / 8)) and saves them, one bit per score, as mask[B*H, N, mask_words];
  // both backward kernels read the saved bits (dK/dV walks the score matrix transposed). Bit position of key
  // token t: t - 1 for t >= 1, N - 1 for t = 0 (key_pos).
  uint32_t drop_thr;  // 0 = off
  float keep_scale;   // 1 / (1 - p_eff)
  uint64_t seed;
  const uint64_t* epoch;  // device epoch counter added to the seed (CUDA-graph replays)
  uint32_t* mask;
  int mask_words;     // ceil(N / 32)
  int mask_ready;     // forward: the mask was drawn ahead of time (nv_dropout_bits): read it instead of drawing
  int ntile_ctas;     // B * H * ceil((N-1)/128): CTAs [0, ntile_ctas) run tiles, the rest token 0 (cta_role)
  int simt_groups;    // (batch, head) pairs served by one token-0 SIMT CTA (1, 2 or 3)
  // raw operands, for the SIMT handling of token 0 (the tiles go through the tensor maps)
  const bf16 *q, *k, *v;
  int64_t qkv_bs, qkv_rs;
};

__device__ __forceinline__ int key_pos(int token, int N) { return token == 0 ? N - 1 : token - 1; }

// ---- SIMT helpers for token 0 -------------------------------------------------------------------------
// 64 bf16 (one head slice of one token) held as eight 16-byte registers; loaded from one global address by every
// thread of the CTA (a broadcast)
struct Vec64 { uint4 c[8]; };
__device__ __forceinline__ void load_vec64(Vec64& v, const bf16* g) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v.c[i] = __ldg(reinterpret_cast<const uint4*>(g) + i);
}
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b, float acc) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 x = unpack_bf16x2(aw[j]), y = unpack_bf16x2(bw[j]);
    acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
  }
  return acc;
}
// 16-byte piece c of row `row` of a [rows x 64 bf16] K-major SW128 tile (as TMA wrote it)
__device__ __forceinline__ uint4 lds_row16(uint32_t tile_u32, int row, int c) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(tile_u32 + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4)));
  return v;
}
// this thread's row of a swizzled tile . a 64-vector in registers
__device__ __forceinline__ float row_dot(uint32_t tile_u32, int row, const Vec64& v) {
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    a0 = dot8(lds_row16(tile_u32, row, c), v.c[c], a0);
    a1 = dot8(lds_row16(tile_u32, row, c + 1), v.c[c + 1], a1);
  }
  return a0 + a1;
}
// this thread's row of tile A . the same row of tile B (both [rows x 64 bf16] SW128 tiles)
__device__ __forceinline__ float row_dot2(uint32_t a_u32, uint32_t b_u32, int row) {
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    a0 = dot8(lds_row16(a_u32, row, c), lds_row16(b_u32, row, c), a0);
    a1 = dot8(lds_row16(a_u32, row, c + 1), lds_row16(b_u32, row, c + 1), a1);
  }
  return a0 + a1;
}
// acc[i] += w * v[32*half + i], i < 32 (TMEM accumulator columns of one half of the head)
__device__ __forceinline__ void axpy_half(uint32_t (&acc)[32], float w, const Vec64& v, int half) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 x = v.c[half * 4 + c];
    const uint32_t xw[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(xw[j]);
      acc[8 * c + 2 * j] = __float_as_uint(fmaf(w, f.x, __uint_as_float(acc[8 * c + 2 * j])));
      acc[8 * c + 2 * j + 1] = __float_as_uint(fmaf(w, f.y, __uint_as_float(acc[8 * c + 2 * j + 1])));
    }
  }
}
// eight bf16 -> fp32 (a 16-bit shift each)
__device__ __forceinline__ void unpack8(const uint4& x, float (&f)[8]) {
  const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
// A SIMT CTA serves SIMT_GROUPS (batch, head) pairs at once, NTHREADS / SIMT_GROUPS threads each. These CTAs are
// instruction-bound (bf16 unpacking and the dot-product shuffles, not the loads), and what they cost is the slot-time
// they hold — a slot is half an SM's shared memory and registers — at the END of the grid, where they fill the slots the
// last partial wave of tile CTAs leaves idle. Two lanes per token (32 dims each: one shuffle per dot product, 64 FMAs per
// 64 unpacks). Pairs per CTA, measured (B=64, N=385, dropout 0.1, fwd / bwd us per layer; N=1729, B=16 in brackets):
// 3 pairs 84.0 / 248.9 (211 / 695), 2 pairs 81.9 / 244.7 (198 / 658), 1 pair 88.1 / 255.0 (186 / 619): with three the
// SIMT CTAs outlast the tile CTAs they run beside (25-50 k cycles against 16-27 k), with one there are more of them than
// idle slots at N=385. simt_groups_for(N) picks two up to 1000 tokens and one above.
constexpr int SIMT_LPT = 2;                             // lanes per token
constexpr int SIMT_DPL = HD / SIMT_LPT;                 // dims per lane (32 = four 16-byte loads)
constexpr int SIMT_WARPS = NTHREADS / 32;
// pairs per SIMT CTA, chosen per launch (Common::simt_groups): the measurements above — two at ViT3D's short sequences,
// one from ~1000 tokens on, where a pair's work is long enough to want the whole CTA
__host__ __device__ inline int simt_groups_for(int N) { return N > 1000 ? 1 : 2; }
struct SimtWho { int g, tl, half, slot, b, h; bool valid; int gthreads, slots, gwarps; };
__device__ __forceinline__ SimtWho simt_who(int cta_j, int B, int H, int groups) {
  SimtWho w;
  w.gthreads = NTHREADS / groups;          // threads per pair (192 / 96 / 64: whole warps)
  w.slots = w.gthreads / SIMT_LPT;         // tokens processed per pass by one pair's threads
  w.gwarps = w.gthreads / 32;
  w.g = threadIdx.x / w.gthreads;
  w.tl = threadIdx.x % w.gthreads;
  w.half = w.tl & 1;
  w.slot = w.tl >> 1;
  const int pair = cta_j * groups + w.g;
  w.valid = pair < B * H;
  const int pc = w.valid ? pair : B * H - 1;   // idle groups shadow the last pair (loads only) to keep barriers uniform
  w.h = pc % H;
  w.b = pc / H;
  return w;
}
struct Row32 { uint4 c[4]; };   // this lane's 32 dims of one token's head slice
__device__ __forceinline__ void load_row32(Row32& r, const bf16* g) {
#pragma unroll
  for (int i = 0; i < 4; ++i) r.c[i] = __ldg(reinterpret_cast<const uint4*>(g) + i);
}
__device__ __forceinline__ void unpack32(const Row32& r, float (&f)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t w[4] = {r.c[i].x, r.c[i].y, r.c[i].z, r.c[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[8 * i + 2 * j] = __uint_as_float(w[j] << 16);            // bf16 -> fp32 is a 16-bit shift
      f[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
    }
  }
}
__device__ __forceinline__ float dot32(const float (&a)[32], const float (&b)[32]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int d = 0; d < 32; d += 4) {
    s0 = fmaf(a[d], b[d], s0); s1 = fmaf(a[d + 1], b[d + 1], s1);
    s2 = fmaf(a[d + 2], b[d + 2], s2); s3 = fmaf(a[d + 3], b[d + 3], s3);
  }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float sum_pair(float s) { return s + __shfl_xor_sync(0xffffffffu, s, 1); }  // the token's two lanes
// fold per-lane partial sums acc[32] (dims 32 half .. 32 half + 31, one token slot per lane pair) over the warp and
// leave them in red[warp][0..63]; the caller adds the pair's warps after a barrier
__device__ __forceinline__ void warp_fold64(float (&acc)[32], float (*red)[HD + 4], int t, int half) {
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    float r = acc[d];
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 16);
    if ((t & 31) < 2) red[t >> 5][32 * half + d] = r;
  }
}

// 1-D grid: the tile CTAs first ((tile, head, batch), tile fastest), then the SIMT CTAs for token 0.
// CTAs are dispatched in index order, so the short SIMT CTAs fill the slots the last partial wave of tile CTAs
// leaves idle instead of taking a tile slot each.
struct CtaRole { int tile, h, b; bool simt; };   // simt: tile = index among the SIMT CTAs
__device__ __forceinline__ CtaRole cta_role(int ntiles, int H, int ntile_ctas) {
  CtaRole r;
  const int id = blockIdx.x;
  r.simt = id >= ntile_ctas;
  if (!r.simt) { r.tile = id % ntiles; r.h = (id / ntiles) % H; r.b = id / (ntiles * H); }
  else { r.tile = id - ntile_ctas; r.h = 0; r.b = 0; }
  return r;
}

// =====================================================================================================
// forward
// =====================================================================================================
struct FwdParams {
  Common c;
  bf16* o;
  int64_t o_bs, o_rs;
  float* lse;
};
constexpr int FWD_KB = 96;                    // keys per block: 2 S buffers (2 x 96) + O (64) = 256 TMEM columns
constexpr int FWD_KV_BYTES = FWD_KB * 128;    // 12 KB per K or V stage
// P operand tile of one block: two 64-key slabs (rows are 128 B apart in both; the second is half used)
constexpr int FWD_SMEM_TILES = SLAB /*Q*/ + 4 * FWD_KV_BYTES /*K x 2, V x 2*/ + 2 * SLAB /*P*/;
constexpr int FWD_SMEM = FWD_SMEM_TILES + 128;
constexpr int FWD_NCH = FWD_KB / 32;
// Lazy rescale: the exponentials use a reference m_used that is only raised when the block maximum exceeds
// it by more than 2^FWD_TAU (log2 domain). Probabilities then stay <= 2^FWD_TAU (exact in fp32 / bf16), O and
// the row sum are rescaled only on those rare raises, and the final O / l is mathematically unchanged.
constexpr float FWD_TAU = 8.0f;

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// One 32-key chunk of one query row: raw scores (registers) -> probabilities, packed as 16 bf16x2 words, and
// the chunk's row-sum. FULL = all 32 keys valid (branch-free); otherwise keys >= nvalid (chunk-local) give 0.
template <bool FULL, bool DROPOUT, bool READY>
__device__ __forceinline__ float fwd_chunk_probs(const uint32_t (&sv)[32], float cs, float m_used, int nvalid,
                                                 const Common& c, uint64_t mrow, int key0, bool row_ok,
                                                 uint32_t (&w)[16], uint32_t km_ready = 0) {
  float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
  uint32_t km32 = 0;
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) {
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      e[i] = ex2(fmaf(__uint_as_float(sv[g8 * 8 + i]), cs, -m_used));
      if (!FULL) e[i] = (g8 * 8 + i) < nvalid ? e[i] : 0.f;
    }
    rs0 += e[0] + e[4]; rs1 += e[1] + e[5]; rs2 += e[2] + e[6]; rs3 += e[3] + e[7];
    if (DROPOUT) {
      uint32_t km;
      if (READY) {
        km = (km_ready >> (8 * g8)) & 0xFFu;  // drawn ahead (nv_dropout_bits), fetched before the barrier waits
      } else {
        km = nv_keep_bits8(nv_seed(c.seed, c.epoch), mrow * (uint64_t)(c.mask_words * 4) + (uint64_t)((key0 >> 3) + g8), 0u,
                           c.drop_thr);
        km32 |= km << (8 * g8);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) e[i] = (km >> i) & 1u ? e[i] : 0.f;   // 1 / (1 - p) is applied once, with 1 / l
    }
    w[g8 * 4 + 0] = pack_bf16x2(e[0], e[1]);
    w[g8 * 4 + 1] = pack_bf16x2(e[2], e[3]);
    w[g8 * 4 + 2] = pack_bf16x2(e[4], e[5]);
    w[g8 * 4 + 3] = pack_bf16x2(e[6], e[7]);
  }
  if (DROPOUT && !READY && row_ok) c.mask[mrow * c.mask_words + (key0 >> 5)] = km32;  // one word per 32-key chunk
  return (rs0 + rs1) + (rs2 + rs3);
}
__device__ __forceinline__ float max32(const uint32_t (&v)[32]) {
  float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    a0 = max3(a0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    a1 = max3(a1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    a2 = max3(a2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
    a3 = max3(a3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
  }
  return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
}
__device__ __forceinline__ float max32_masked(const uint32_t (&v)[32], int nvalid) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) mx = i < nvalid ? fmaxf(mx, __uint_as_float(v[i])) : mx;
  return mx;
}
// 16 packed words = 32 probabilities of row `row`, keys [32*chunk, 32*chunk+32) of the block -> operand tile
__device__ __forceinline__ void store_p_chunk(uint32_t sP_u32, int row, int chunk, const uint32_t (&w)[16]) {
  const uint32_t prow = sP_u32 + (uint32_t)(chunk >> 1) * SLAB + (uint32_t)row * 128;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int ch = ((chunk & 1) << 2) + g;
    st_shared_v4(prow + (uint32_t)((ch ^ (row & 7)) << 4), w[g * 4], w[g * 4 + 1], w[g * 4 + 2], w[g * 4 + 3]);
  }
}

// Forward of query token 0: all N keys, SIMT. Two lanes share a key (32 dims each); one pass with a running
// (max, sum, o) per lane pair, merged over the pair-group's two warps at the end.
template <bool DROPOUT>
__device__ __forceinline__ void fwd_cls_query(const FwdParams& p, float* sm, int cta_j) {
  const Common& c = p.c;
  const SimtWho w = simt_who(cta_j, c.B, c.H, c.simt_groups);
  const int t = threadIdx.x, b = w.b, h = w.h;
  const int N = c.N;
  const int64_t bh = (int64_t)b * c.H + h;
  float (*red)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(sm);
  uint32_t* mw_s = reinterpret_cast<uint32_t*>(red + SIMT_WARPS) + w.g * c.mask_words;  // keep words of mask row (bh, 0)
  const float cs = c.scale * LOG2E;
  const bf16* kb = c.k + (int64_t)b * c.qkv_bs + h * HD + SIMT_DPL * w.half;
  const bf16* vb = c.v + (int64_t)b * c.qkv_bs + h * HD + SIMT_DPL * w.half;
  float qc[32];
  {
    Row32 qr;
    load_row32(qr, c.q + (int64_t)b * c.qkv_bs + h * HD + SIMT_DPL * w.half);
    unpack32(qr, qc);
#pragma unroll
    for (int d = 0; d < 32; ++d) qc[d] *= cs;   // scores come out in the log2 domain
  }
  if (DROPOUT) {
    const int64_t mrow = bh * N;  // query token 0
    for (int wi = w.tl; wi < c.mask_words; wi += w.gthreads) {
      uint32_t word;
      if (c.mask_ready) {
        word = __ldg(c.mask + mrow * c.mask_words + wi);
      } else {
        const uint64_t seed = nv_seed(c.seed, c.epoch);
        word = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          word |= nv_keep_bits8(seed, (uint64_t)mrow * (uint64_t)(c.mask_words * 4) + (uint64_t)(4 * wi + g), 0u, c.drop_thr) << (8 * g);
        if (w.valid) c.mask[mrow * c.mask_words + wi] = word;
      }
      mw_s[wi] = word;
    }
    __syncthreads();
  }
  float m_g = -INFINITY, l_g = 0.f, oacc[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) oacc[d] = 0.f;
  // software-pipelined: the next key's rows are in flight while this key is processed (these warps are alone with
  // their load latency: nothing else hides it)
  Row32 kk, vv;
  {
    const int64_t roff = (int64_t)min(w.slot, N - 1) * c.qkv_rs;
    load_row32(kk, kb + roff);
    load_row32(vv, vb + roff);
  }
  for (int key0 = 0; key0 < N; key0 += w.slots) {   // uniform trip count: the shuffles need whole warps
    const int key = key0 + w.slot;
    Row32 kn, vn;
    {
      const int64_t roff = (int64_t)min(key + w.slots, N - 1) * c.qkv_rs;
      load_row32(kn, kb + roff);
      load_row32(vn, vb + roff);
    }
    float sj;
    {
      float kf[32];
      unpack32(kk, kf);
      const float sred = sum_pair(dot32(qc, kf));   // the shuffle must run in every lane: never under the `live` predicate
      sj = key < N ? sred : -INFINITY;
    }
    const float m_new = fmaxf(m_g, sj);
    const float m_safe = m_new == -INFINITY ? 0.f : m_new;   // a lane pair that has seen no key yet
    const float alpha = ex2(m_g - m_safe);                   // 0 on the first key
    const float pj = ex2(sj - m_safe);                       // 0 for keys past N
    l_g = fmaf(l_g, alpha, pj);
    float pd = pj;
    if (DROPOUT) {
      const int pos = key_pos(key < N ? key : 0, N);
      pd = (mw_s[pos >> 5] >> (pos & 31)) & 1u ? pj * c.keep_scale : 0.f;
    }
    {
      float vf[32];
      unpack32(vv, vf);
#pragma unroll
      for (int d = 0; d < 32; ++d) oacc[d] = fmaf(oacc[d], alpha, pd * vf[d]);
    }
    m_g = m_new;
    kk = kn; vv = vn;
  }
  // merge the pair-group's lane pairs: common reference M, then plain sums
  const int w0 = w.g * w.gwarps;  // first warp of this pair-group
  float mx = warp_max(m_g);
  if ((t & 31) == 0) red[t >> 5][HD + 1] = mx;
  __syncthreads();
  float M = red[w0][HD + 1];
#pragma unroll
  for (int i = 1; i < w.gwarps; ++i) M = fmaxf(M, red[w0 + i][HD + 1]);
  const float sc_g = ex2(m_g - M);  // 0 for lane pairs without keys
#pragma unroll
  for (int d = 0; d < 32; ++d) oacc[d] *= sc_g;
  warp_fold64(oacc, red, t, w.half);
  const float l = warp_sum(w.half == 0 ? l_g * sc_g : 0.f);
  if ((t & 31) == 0) red[t >> 5][HD] = l;
  __syncthreads();
  if (w.tl < HD && w.valid) {
    float r = 0.f, L = 0.f;
#pragma unroll
    for (int i = 0; i < w.gwarps; ++i) { r += red[w0 + i][w.tl]; L += red[w0 + i][HD]; }
    p.o[(int64_t)b * p.o_bs + h * HD + w.tl] = __float2bfloat16(r / L);
    if (w.tl == 0) p.lse[bh * N] = (M + log2f(L)) * LN2;
  }
}

// READY: the keep bits were drawn ahead of time (nv_dropout_bits, the training step's path); the inline Philox draw is
// then not compiled into the kernel at all — the hot loop of the two-path kernel missed the instruction cache in
// 13 % of its stall samples (profiles/r02_summary.md).
template <bool DROPOUT, bool READY>
__global__ void __launch_bounds__(NTHREADS, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + SLAB;                       // stage s at s * FWD_KV_BYTES
  uint8_t* sV = smem + SLAB + 2 * FWD_KV_BYTES;
  uint8_t* sP = smem + SLAB + 4 * FWD_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FWD_SMEM_TILES);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // 2
  uint64_t* k_empty = bars + 3;     // 2
  uint64_t* v_full = bars + 5;      // 2
  uint64_t* v_empty = bars + 7;     // 2
  uint64_t* s_full = bars + 9;      // 2
  uint64_t* sp_ready = bars + 11;   // 1
  uint64_t* o_full = bars + 12;     // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.c.N;
  const int n = N - 1;  // tokens 1 .. n are tiled; tile-local index r <-> token r + 1
  const CtaRole role = cta_role((n + BQ - 1) / BQ, p.c.H, p.c.ntile_ctas);
  const int b = role.b, h = role.h, q0 = role.tile * BQ;
  if (role.simt) {  // query token 0, SIMT (no barriers / TMEM in use yet)
    PROF_DECL
    fwd_cls_query<DROPOUT>(p, reinterpret_cast<float*>(smem), role.tile);
    PROF_MARK(0);
#ifdef NV_PROFILE
    if (role.tile == 14 && threadIdx.x == 0) PROF_DUMP(12);
#endif
    return;
  }
  const int nblk = (n + FWD_KB - 1) / FWD_KB;
  const int nactive = (min(BQ, n - q0) + 31) >> 5;  // row warps with at least one valid query

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("attn_tc_fwd: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(sp_ready, 32 * nactive);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 2 * FWD_KB;  // S buffers at +0, +96; O at +192

  if (warp == TMA_WARP) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, SLAB);
      tma_load_3d(sQ, &tq, q_full, h * HD, 1 + q0, b);  // one 128-row box
    }
    __syncwarp();
    for (int j = 0; j < nblk; ++j) {
      const int s = j & 1;
      const uint32_t ph = ((j >> 1) & 1) ^ 1;
      mbar_wait(&k_empty[s], ph);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[s], FWD_KV_BYTES);
        tma_load_3d(sK + s * FWD_KV_BYTES, &tk, &k_full[s], h * HD, 1 + j * FWD_KB, b);
      }
      __syncwarp();
      mbar_wait(&v_empty[s], ph);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[s], FWD_KV_BYTES);
        tma_load_3d(sV + s * FWD_KV_BYTES, &tv, &v_full[s], h * HD, 1 + j * FWD_KB, b);
      }
      __syncwarp();
    }
  } else if (warp == MMA_WARP) {
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD, 0, 1);
    const uint64_t q_desc = kmajor_desc(sQ);
    auto cols16 = [&](int j) { return (min(FWD_KB, n - j * FWD_KB) + 15) & ~15; };
    auto issue_s = [&](int j) {  // S_j = Q K_j^T into S buffer j & 1; frees the K stage when done
      const int s = j & 1;
      mma_k64(tS + (uint32_t)(s * FWD_KB), q_desc, kmajor_desc(sK + s * FWD_KV_BYTES), umma_idesc_bf16(128, cols16(j), 0, 0));
      umma_commit(&s_full[s]);
      umma_commit(&k_empty[s]);
    };
    mbar_wait(q_full, 0);
    for (int j = 0; j < 2 && j < nblk; ++j) {
      mbar_wait(&k_full[j], 0);
      tc_fence_after();
      if (elect_one()) issue_s(j);
      __syncwarp();
    }
    for (int j = 0; j < nblk; ++j) {
      const int s = j & 1;
      mbar_wait(sp_ready, j & 1);            // rows: S_j consumed, P_j in smem, O rescaled if it had to be
      mbar_wait(&v_full[s], (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        mma_rows(tO, sP, sV + s * FWD_KV_BYTES, idesc_pv, cols16(j) >> 4, j > 0);  // O += P_j V_j
        umma_commit(o_full);
        umma_commit(&v_empty[s]);
      }
      __syncwarp();
      if (j + 2 < nblk) {
        mbar_wait(&k_full[s], ((j + 2) >> 1) & 1);
        tc_fence_after();
        if (elect_one()) issue_s(j + 2);
        __syncwarp();
      }
    }
  } else if (warp < nactive) {
    const int row = warp * 32 + lane;  // row within the tile == TMEM lane
    const int qrow = q0 + row;         // tile-local query index; token = qrow + 1
    const int token = min(qrow + 1, N - 1);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t sP_u32 = smem_u32(sP);
    const float cs = p.c.scale * LOG2E;
    const int bh = b * p.c.H + h;
    const uint64_t mrow = (uint64_t)bh * N + token;
    // key token 0: its raw score from this row's Q in shared memory; it also seeds the softmax reference, so the
    // probability computed at the end (with the final reference) cannot overflow
    PROF_DECL
    float s_x;
    {
      Vec64 k0;
      load_vec64(k0, p.c.k + (int64_t)b * p.c.qkv_bs + h * HD);   // in flight while the Q tile lands
      mbar_wait(q_full, 0);
      s_x = row_dot(smem_u32(sQ), row, k0);
    }
    float m_used = s_x * cs, l_run = 0.f;
    PROF_MARK(8);

    for (int j = 0; j < nblk; ++j) {
      const int key0 = j * FWD_KB;     // tile-local key index == mask bit position
      const int nvalid = min(FWD_KB, n - key0);
      const bool full = nvalid == FWD_KB;
      const uint32_t tSj = tS + lane_base + (uint32_t)((j & 1) * FWD_KB);
      uint32_t kmw[3] = {0u, 0u, 0u};  // pre-drawn dropout keep words of this row's block
      if (DROPOUT && READY) {
        const uint32_t* mp = p.c.mask + mrow * p.c.mask_words + (key0 >> 5);
        kmw[0] = __ldg(mp);
        if (nvalid > 32) kmw[1] = __ldg(mp + 1);
        if (nvalid > 64) kmw[2] = __ldg(mp + 2);
      }
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      PROF_MARK(j == 0 ? 0 : 1);
      // the whole block of raw scores of this row -> registers (one TMEM pass)
      uint32_t c0[32], c1[32], c2[32];
      tmem_ld_32x32(tSj, c0);
      if (nvalid > 32) tmem_ld_32x32(tSj + 32, c1);
      if (nvalid > 64) tmem_ld_32x32(tSj + 64, c2);
      tmem_ld_wait();
      float mx;
      if (full) {
        mx = max3(max32(c0), max32(c1), max32(c2));
      } else {
        mx = max32_masked(c0, nvalid);
        if (nvalid > 32) mx = fmaxf(mx, max32_masked(c1, nvalid - 32));
        if (nvalid > 64) mx = fmaxf(mx, max32_masked(c2, nvalid - 64));
      }
      const float m_blk = mx * cs;  // finite: every block holds >= 1 valid key
      const bool raise = m_blk > m_used + FWD_TAU;  // always true for the first block (m_used = -inf)
      const float m_new = raise ? m_blk : m_used;
      const float alpha = raise ? ex2(m_used - m_new) : 1.0f;
      m_used = m_new;
      PROF_MARK(2);
      // probabilities -> K-major SW128 operand rows; the P V MMA reads ceil16(nvalid) key columns.
      // The first chunk's exponentials are computed before waiting for the previous P V (which still reads
      // P's buffer), so that wait is normally already satisfied.
      const bool row_ok = qrow < n;
      uint32_t w[16];
      float rs;
      if (full) rs = fwd_chunk_probs<true, DROPOUT, READY>(c0, cs, m_used, 32, p.c, mrow, key0, row_ok, w, kmw[0]);
      else      rs = fwd_chunk_probs<false, DROPOUT, READY>(c0, cs, m_used, nvalid, p.c, mrow, key0, row_ok, w, kmw[0]);
      PROF_MARK(3);
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
      }
      PROF_MARK(4);
      store_p_chunk(sP_u32, row, 0, w);
      if (full) {
        rs += fwd_chunk_probs<true, DROPOUT, READY>(c1, cs, m_used, 32, p.c, mrow, key0 + 32, row_ok, w, kmw[1]);
        store_p_chunk(sP_u32, row, 1, w);
        rs += fwd_chunk_probs<true, DROPOUT, READY>(c2, cs, m_used, 32, p.c, mrow, key0 + 64, row_ok, w, kmw[2]);
        store_p_chunk(sP_u32, row, 2, w);
      } else {
        if (nvalid > 32) {
          rs += fwd_chunk_probs<false, DROPOUT, READY>(c1, cs, m_used, nvalid - 32, p.c, mrow, key0 + 32, row_ok, w, kmw[1]);
          store_p_chunk(sP_u32, row, 1, w);
        }
        if (nvalid > 64) {
          rs += fwd_chunk_probs<false, DROPOUT, READY>(c2, cs, m_used, nvalid - 64, p.c, mrow, key0 + 64, row_ok, w, kmw[2]);
          store_p_chunk(sP_u32, row, 2, w);
        }
      }
      l_run = fmaf(l_run, alpha, rs);
      if (j > 0 && __any_sync(0xffffffffu, raise)) {  // rare: bring the TMEM accumulator to the new reference
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tmem_ld_32x32(tO + lane_base + hh * 32, c0);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) c0[i] = __float_as_uint(__uint_as_float(c0[i]) * alpha);
          tmem_st_32x32(tO + lane_base + hh * 32, c0);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(sp_ready);
      PROF_MARK(5);
    }
    Vec64 v0;
    load_vec64(v0, p.c.v + (int64_t)b * p.c.qkv_bs + h * HD);   // in flight under the last P V
    mbar_wait(o_full, (nblk - 1) & 1);  // last block's P V
    tc_fence_after();
    PROF_MARK(6);
    {
      // key token 0 joins here: p_x = 2^(s_x c - m) with the final reference, O += dropout(p_x) v_0
      const float p_x = ex2(fmaf(s_x, cs, -m_used));
      l_run += p_x;
      float p_xd = p_x;
      if (DROPOUT) {
        const int pos = N - 1;
        uint32_t keep;
        if (READY) {
          keep = (__ldg(p.c.mask + mrow * p.c.mask_words + (pos >> 5)) >> (pos & 31)) & 1u;
        } else {
          const uint32_t byte = nv_keep_bits8(nv_seed(p.c.seed, p.c.epoch),
                                              mrow * (uint64_t)(p.c.mask_words * 4) + (uint64_t)(pos >> 3), 0u, p.c.drop_thr);
          keep = (byte >> (pos & 7)) & 1u;
          // the word of position N-1 is written by the last key chunk unless that chunk ended on a word boundary
          if ((pos & 31) == 0 && qrow < n) p.c.mask[mrow * p.c.mask_words + (pos >> 5)] = byte;
        }
        p_xd = keep ? p_x : 0.f;
      }
      // dropped probabilities went into P unscaled (one select per score instead of a multiply and a select): the
      // survivors' 1 / (1 - p) rides on the normalisation
      const float inv = (DROPOUT ? p.c.keep_scale : 1.0f) / l_run;
      bf16* op = p.o + (int64_t)b * p.o_bs + (int64_t)token * p.o_rs + h * HD;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {  // tcgen05.ld is warp-collective: every lane loads, valid rows store
        uint32_t v[32];
        tmem_ld_32x32(tO + lane_base + hh * 32, v);
        tmem_ld_wait();
        axpy_half(v, p_xd, v0, hh);
        if (qrow < n) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 t;
            t.x = pack_bf16x2(__uint_as_float(v[8 * i]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
            t.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
            t.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
            t.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
            *reinterpret_cast<uint4*>(op + hh * 32 + 8 * i) = t;
          }
        }
      }
      if (qrow < n) p.lse[((int64_t)b * p.c.H + h) * N + token] = (m_used + log2f(l_run)) * LN2;
    }
    PROF_MARK(7);
#ifdef NV_PROFILE
    if (role.tile == 1 && h == 3 && (b == 5 || b == 40) && (threadIdx.x == 0 || threadIdx.x == 70))
      PROF_DUMP((b == 40 ? 2 : 0) + (threadIdx.x == 70 ? 1 : 0));
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// =====================================================================================================
// backward, dQ: CTA = 128 queries; loop over 64-key blocks
// =====================================================================================================
struct BwdParams {
  Common c;
  const float* lse;    // [B,H,N] natural log
  float* delta;        // [B,H,N]: written by the dQ kernel (delta_i = dO_i . O_i), read by the dK/dV kernel
  const bf16 *o, *dO;  // [B,N,H*64] (o_bs, o_rs)
  int64_t o_bs, o_rs;
  bf16* dq; bf16* dk; bf16* dv;
  int64_t d_bs, d_rs;
};

// dQ of query token 0 (and its delta): all N keys, SIMT, two lanes per key.
__device__ __forceinline__ void bwd_cls_query_dq(const BwdParams& p, float* sm, int cta_j) {
  const Common& c = p.c;
  const SimtWho w = simt_who(cta_j, c.B, c.H, c.simt_groups);
  const int t = threadIdx.x, b = w.b, h = w.h;
  const int N = c.N;
  const int64_t bh = (int64_t)b * c.H + h;
  float (*red)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(sm);
  const float cs = c.scale * LOG2E;
  const bool dropout = c.drop_thr != 0;
  float qc[32], doc[32];
  float delta;
  {
    Row32 r;
    float oc[32];
    load_row32(r, c.q + (int64_t)b * c.qkv_bs + h * HD + SIMT_DPL * w.half);
    unpack32(r, qc);
    load_row32(r, p.dO + (int64_t)b * p.o_bs + h * HD + SIMT_DPL * w.half);
    unpack32(r, doc);
    load_row32(r, p.o + (int64_t)b * p.o_bs + h * HD + SIMT_DPL * w.half);
    unpack32(r, oc);
    delta = sum_pair(dot32(doc, oc));
  }
  if (w.tl == 0 && w.valid) p.delta[bh * N] = delta;
  const float lse2 = __ldg(p.lse + bh * N) * LOG2E;
  const bf16* kb = c.k + (int64_t)b * c.qkv_bs + h * HD + SIMT_DPL * w.half;
  const bf16* vb = c.v + (int64_t)b * c.qkv_bs + h * HD + SIMT_DPL * w.half;
  const uint32_t* mrow0 = c.mask + bh * N * c.mask_words;   // mask row of query token 0
  float acc[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) acc[d] = 0.f;
  Row32 kk, vv;   // software-pipelined like the forward
  {
    const int64_t roff = (int64_t)min(w.slot, N - 1) * c.qkv_rs;
    load_row32(kk, kb + roff);
    load_row32(vv, vb + roff);
  }
  for (int key0 = 0; key0 < N; key0 += w.slots) {
    const int key = key0 + w.slot;
    const bool live = key < N;
    const int kc = live ? key : 0;
    Row32 kn, vn;
    {
      const int64_t roff = (int64_t)min(key + w.slots, N - 1) * c.qkv_rs;
      load_row32(kn, kb + roff);
      load_row32(vn, vb + roff);
    }
    const int pos = key_pos(kc, N);
    const uint32_t mw = dropout ? __ldg(mrow0 + (pos >> 5)) : 0xFFFFFFFFu;
    float kf[32], dp;
    {
      float vf[32];
      unpack32(vv, vf);
      dp = sum_pair(dot32(doc, vf));
    }
    unpack32(kk, kf);
    const float sacc = sum_pair(dot32(qc, kf));
    const float pj = ex2(fmaf(sacc, cs, -lse2));
    const bool keep = (mw >> (pos & 31)) & 1u;
    const float ds = live ? pj * ((keep ? dp * c.keep_scale : 0.f) - delta) : 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) acc[d] = fmaf(ds, kf[d], acc[d]);
    kk = kn; vv = vn;
  }
  warp_fold64(acc, red, t, w.half);
  __syncthreads();
  if (w.tl < HD && w.valid) {
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < w.gwarps; ++i) r += red[w.g * w.gwarps + i][w.tl];
    p.dq[(int64_t)b * p.d_bs + h * HD + w.tl] = __float2bfloat16(r * c.scale);
  }
}

// dK and dV of key token 0: a reduction over all N queries, SIMT, two lanes per query.
// Reads delta (all queries), so it runs in the dK/dV grid, after the dQ grid has written it.
__device__ __forceinline__ void bwd_extra_key_dkv(const BwdParams& p, float* sm, int cta_j) {
  const Common& c = p.c;
  const SimtWho w = simt_who(cta_j, c.B, c.H, c.simt_groups);
  const int t = threadIdx.x, b = w.b, h = w.h;
  const int N = c.N;
  const int64_t bh = (int64_t)b * c.H + h;
  float (*red)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(sm);
  float (*red2)[HD + 4] = red + SIMT_WARPS;
  const float cs = c.scale * LOG2E;
  const bool dropout = c.drop_thr != 0;
  const int pos = N - 1;  // mask bit position of key token 0
  // k and v of token 0 are re-read (L1 hits) piece by piece for every query: the registers go to the two accumulators
  const uint4* k0p = reinterpret_cast<const uint4*>(c.k + (int64_t)b * c.qkv_bs + h * HD + SIMT_DPL * w.half);
  const uint4* v0p = reinterpret_cast<const uint4*>(c.v + (int64_t)b * c.qkv_bs + h * HD + SIMT_DPL * w.half);
  const bf16* qb = c.q + (int64_t)b * c.qkv_bs + h * HD + SIMT_DPL * w.half;
  const bf16* dob = p.dO + (int64_t)b * p.o_bs + h * HD + SIMT_DPL * w.half;
  float dk_acc[32], dv_acc[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) dk_acc[d] = dv_acc[d] = 0.f;
  Row32 qq, dd;   // software-pipelined like the forward
  float l2, dl;
  uint32_t mw = 0xFFFFFFFFu;
  {
    const int ic = min(w.slot, N - 1);
    load_row32(qq, qb + (int64_t)ic * c.qkv_rs);
    load_row32(dd, dob + (int64_t)ic * p.o_rs);
    l2 = __ldg(p.lse + bh * N + ic) * LOG2E;
    dl = p.delta[bh * N + ic];   // written by the dQ grid: plain load
    if (dropout) mw = __ldg(c.mask + (bh * N + ic) * c.mask_words + (pos >> 5));
  }
  for (int i0 = 0; i0 < N; i0 += w.slots) {
    const int i = i0 + w.slot;
    const bool live = i < N;
    Row32 qn, dn;
    float l2n, dln;
    uint32_t mwn = 0xFFFFFFFFu;
    {
      const int ic = min(i + w.slots, N - 1);
      load_row32(qn, qb + (int64_t)ic * c.qkv_rs);
      load_row32(dn, dob + (int64_t)ic * p.o_rs);
      l2n = __ldg(p.lse + bh * N + ic) * LOG2E;
      dln = p.delta[bh * N + ic];
      if (dropout) mwn = __ldg(c.mask + (bh * N + ic) * c.mask_words + (pos >> 5));
    }
    // 16-byte piece at a time: the rows stay packed in registers (the two accumulators need the space)
    float s0 = 0.f, s1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a[8], bb[8];
      unpack8(qq.c[i], a); unpack8(__ldg(k0p + i), bb);
#pragma unroll
      for (int j = 0; j < 8; j += 2) { s0 = fmaf(a[j], bb[j], s0); s1 = fmaf(a[j + 1], bb[j + 1], s1); }
      unpack8(dd.c[i], a); unpack8(__ldg(v0p + i), bb);
#pragma unroll
      for (int j = 0; j < 8; j += 2) { d0 = fmaf(a[j], bb[j], d0); d1 = fmaf(a[j + 1], bb[j + 1], d1); }
    }
    const float sacc = sum_pair(s0 + s1), dp = sum_pair(d0 + d1);
    const float pj = ex2(fmaf(sacc, cs, -l2));
    const bool keep = (mw >> (pos & 31)) & 1u;
    const float pm = (live && keep) ? pj * c.keep_scale : 0.f;
    const float ds = live ? pj * ((keep ? dp * c.keep_scale : 0.f) - dl) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a[8];
      unpack8(qq.c[i], a);
#pragma unroll
      for (int j = 0; j < 8; ++j) dk_acc[8 * i + j] = fmaf(ds, a[j], dk_acc[8 * i + j]);
      unpack8(dd.c[i], a);
#pragma unroll
      for (int j = 0; j < 8; ++j) dv_acc[8 * i + j] = fmaf(pm, a[j], dv_acc[8 * i + j]);
    }
    qq = qn; dd = dn; l2 = l2n; dl = dln; mw = mwn;
  }
  warp_fold64(dk_acc, red, t, w.half);
  warp_fold64(dv_acc, red2, t, w.half);
  __syncthreads();
  if (w.tl < HD && w.valid) {
    float rk = 0.f, rv = 0.f;
#pragma unroll
    for (int i = 0; i < w.gwarps; ++i) { rk += red[w.g * w.gwarps + i][w.tl]; rv += red2[w.g * w.gwarps + i][w.tl]; }
    p.dk[(int64_t)b * p.d_bs + h * HD + w.tl] = __float2bfloat16(rk * c.scale);
    p.dv[(int64_t)b * p.d_bs + h * HD + w.tl] = __float2bfloat16(rv);
  }
}
constexpr int BWD_CB = 64;  // columns per block (keys in dQ, queries in dKV)
// S is double-buffered in TMEM (2 x 64 columns + dP 64 + dQ 64 = the CTA's 256) and issued two key blocks ahead through
// a three-stage K/V ring; dP is issued one block ahead. With a single S buffer (round 1 / early round 2) the row warps
// waited 1.1-1.4 k cycles per 64-key block for the round trip rows -> MMA warp -> S, dP -> commit -> rows — as long as
// their own arithmetic; now S is already there and the dP round trip runs under the block's exponentials
// (118.8 -> 108.5 us per layer at B=64, N=385, dropout 0.1; profiles/r02_summary.md §4).
constexpr int DQ_STAGES = 3;
constexpr int DQ_SMEM_TILES = 2 * SLAB /*Q, dO*/ + DQ_STAGES * 2 * BOX /*K,V ring*/ + SLAB /*dS*/;
constexpr int DQ_SMEM = DQ_SMEM_TILES + 256 + 1024;

__global__ void __launch_bounds__(NTHREADS, 2)
attn_tc_bwd_dq_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                      const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                      const __grid_constant__ CUtensorMap to, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + SLAB;
  uint8_t* sKV = smem + 2 * SLAB;            // stage s (of DQ_STAGES): K at s*2*BOX, V at s*2*BOX + BOX
  uint8_t* sdS = smem + 2 * SLAB + DQ_STAGES * 2 * BOX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_SMEM_TILES);
  uint64_t* qd_full = bars;         // 1
  uint64_t* kv_full = bars + 1;     // 3
  uint64_t* kv_empty = bars + 4;    // 3
  uint64_t* s_full = bars + 7;      // 2  S is double-buffered in TMEM
  uint64_t* dp_full = bars + 9;     // 1
  uint64_t* ds_ready = bars + 10;   // 1
  uint64_t* ds_free = bars + 11;    // 1  (also "dQ accumulated" after the last block)
  uint64_t* o_full = bars + 12;     // 1  the O tile, parked in the dS slab until the rows have taken delta from it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.c.N;
  const int n = N - 1;  // tokens 1 .. n are tiled; tile-local index r <-> token r + 1
  const CtaRole role = cta_role((n + BQ - 1) / BQ, p.c.H, p.c.ntile_ctas);
  const int b = role.b, h = role.h, q0 = role.tile * BQ;
  if (role.simt) {  // query token 0 (delta, dQ), SIMT
    PROF_DECL
    bwd_cls_query_dq(p, reinterpret_cast<float*>(smem), role.tile);
    PROF_MARK(0);
#ifdef NV_PROFILE
    if (role.tile == 14 && threadIdx.x == 0) PROF_DUMP(13);
#endif
    return;
  }
  const int nblk = (n + BWD_CB - 1) / BWD_CB;
  const int nactive = (min(BQ, n - q0) + 31) >> 5;

  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo); tma_prefetch_desc(&to);
    mbar_init(qd_full, 1);
    for (int i = 0; i < DQ_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
    mbar_init(dp_full, 1);
    mbar_init(ds_ready, 32 * nactive);
    mbar_init(ds_free, 1);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base /* two buffers: +0, +64 */, tdP = tmem_base + 128, tdQ = tmem_base + 192;

  if (warp == TMA_WARP) {
    if (elect_one()) {
      mbar_arrive_expect_tx(qd_full, 2 * SLAB);
      load_rows(sQ, &tq, qd_full, h * HD, 1 + q0, b, 2);
      load_rows(sdO, &tdo, qd_full, h * HD, 1 + q0, b, 2);
      // O rows of the tile -> the dS slab (free until the first dS is written): each row thread reads delta = dO . O
      // from its own row and only then overwrites that row with dS, so no other synchronisation is needed
      mbar_arrive_expect_tx(o_full, SLAB);
      load_rows(sdS, &to, o_full, h * HD, 1 + q0, b, 2);
    }
    __syncwarp();
    for (int j = 0; j < nblk; ++j) {
      const int s = j % DQ_STAGES;
      mbar_wait(&kv_empty[s], ((j / DQ_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[s], 2 * BOX);
        load_rows(sKV + s * 2 * BOX, &tk, &kv_full[s], h * HD, 1 + j * BWD_CB, b, 1);
        load_rows(sKV + s * 2 * BOX + BOX, &tv, &kv_full[s], h * HD, 1 + j * BWD_CB, b, 1);
      }
      __syncwarp();
    }
  } else if (warp == MMA_WARP) {
    constexpr uint32_t idesc_dq = umma_idesc_bf16(128, HD, 0, 1);
    const uint64_t q_desc = kmajor_desc(sQ), do_desc = kmajor_desc(sdO);
    auto cols16 = [&](int j) { return (min(BWD_CB, n - j * BWD_CB) + 15) & ~15; };
    auto kv_stage = [&](int j) { return sKV + (j % DQ_STAGES) * 2 * BOX; };
    auto kv_wait = [&](int j) { mbar_wait(&kv_full[j % DQ_STAGES], (j / DQ_STAGES) & 1); };
    auto issue_s = [&](int j) {    // S_j = Q K_j^T into S buffer j & 1
      mma_k64(tS + (uint32_t)((j & 1) * BWD_CB), q_desc, kmajor_desc(kv_stage(j)), umma_idesc_bf16(128, cols16(j), 0, 0));
      umma_commit(&s_full[j & 1]);
    };
    auto issue_dp = [&](int j) {   // dP_j = dO V_j^T (single buffer)
      mma_k64(tdP, do_desc, kmajor_desc(kv_stage(j) + BOX), umma_idesc_bf16(128, cols16(j), 0, 0));
      umma_commit(dp_full);
    };
    // Schedule: S runs TWO blocks ahead of the rows (its buffer pair frees one when the rows have read it), dP one block
    // ahead, right behind the rows' arrival; the rows find S_j waiting and spend the dP round trip on their exponentials.
    mbar_wait(qd_full, 0);
    kv_wait(0);
    tc_fence_after();
    if (elect_one()) { issue_s(0); issue_dp(0); }
    __syncwarp();
    if (nblk > 1) {
      kv_wait(1);
      tc_fence_after();
      if (elect_one()) issue_s(1);
      __syncwarp();
    }
    for (int j = 0; j < nblk; ++j) {
      if (j + 2 < nblk) kv_wait(j + 2);      // K/V of the block whose scores are issued below (stage freed by dQ_{j-1})
      mbar_wait(ds_ready, j & 1);            // rows: S_j and dP_j consumed, dS_j in shared memory
      tc_fence_after();
      if (elect_one()) {
        if (j + 1 < nblk) issue_dp(j + 1);   // its K/V stage was waited for when S_{j+1} was issued
        mma_rows(tdQ, sdS, kv_stage(j), idesc_dq, cols16(j) >> 4, j > 0);  // dQ += dS K_j
        umma_commit(&kv_empty[j % DQ_STAGES]);
        umma_commit(ds_free);
        if (j + 2 < nblk) issue_s(j + 2);
      }
      __syncwarp();
    }
  } else if (warp < nactive) {
    const int row = warp * 32 + lane;
    const int qrow = q0 + row;               // tile-local query index; token = qrow + 1
    const int token = min(qrow + 1, N - 1);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t sdS_u32 = smem_u32(sdS);
    const float cs = p.c.scale * LOG2E;
    const bool dropout = p.c.drop_thr != 0;
    const int64_t stat = ((int64_t)b * p.c.H + h) * N + token;
    const uint64_t mrow = (uint64_t)stat;
    const float lse2 = p.lse[stat] * LOG2E;
    // prologue on this row's Q / dO in shared memory: delta_i = dO_i . O_i (also published for the dK/dV kernel)
    // and the dS of key token 0, which stays outside the tiles
    PROF_DECL
    float dl, ds_x;
    {
      Vec64 v0, k0;   // in flight while the Q / dO / O tiles land
      load_vec64(v0, p.c.v + (int64_t)b * p.c.qkv_bs + h * HD);
      load_vec64(k0, p.c.k + (int64_t)b * p.c.qkv_bs + h * HD);
      uint32_t kw = 0xFFFFFFFFu;
      if (dropout) kw = __ldg(p.c.mask + mrow * p.c.mask_words + ((N - 1) >> 5));
      mbar_wait(qd_full, 0);
      mbar_wait(o_full, 0);
      dl = row_dot2(smem_u32(sdO), sdS_u32, row);
      if (qrow < n) p.delta[stat] = dl;
      const float dp_x = row_dot(smem_u32(sdO), row, v0);
      const float p_x = ex2(fmaf(row_dot(smem_u32(sQ), row, k0), cs, -lse2));
      const bool keep = (kw >> ((N - 1) & 31)) & 1u;
      ds_x = p_x * ((keep ? dp_x * p.c.keep_scale : 0.f) - dl);
    }
    PROF_MARK(8);
    for (int j = 0; j < nblk; ++j) {
      const int key0 = j * BWD_CB;           // tile-local key index == mask bit position
      const int nch = (min(BWD_CB, n - key0) + 31) >> 5;
      uint32_t kmw[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};  // dropout keep words of the block, fetched before the waits
      if (dropout) {
        kmw[0] = __ldg(p.c.mask + mrow * p.c.mask_words + (key0 >> 5));
        if (nch > 1) kmw[1] = __ldg(p.c.mask + mrow * p.c.mask_words + (key0 >> 5) + 1);
      }
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);   // issued two blocks ago: normally complete
      tc_fence_after();
      PROF_MARK(j == 0 ? 0 : 1);
      // probabilities of the whole block first (they need S only): the dP MMA issued at this row's last arrival
      // completes underneath
      float pe[2 * 32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c < nch) {
          uint32_t sv[32];
          tmem_ld_32x32(tS + lane_base + (uint32_t)((j & 1) * BWD_CB + c * 32), sv);
          tmem_ld_wait();
          // keys >= N: K rows are zero-filled, so whatever finite dS lands there multiplies zeros in dS K
#pragma unroll
          for (int i = 0; i < 32; ++i) pe[c * 32 + i] = ex2(fmaf(__uint_as_float(sv[i]), cs, -lse2));
        }
      }
      PROF_MARK(2);
      mbar_wait(dp_full, j & 1);
      tc_fence_after();
      PROF_MARK(6);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c < nch) {
          uint32_t dv[32];
          tmem_ld_32x32(tdP + lane_base + c * 32, dv);
          tmem_ld_wait();
          float ds[32];
          if (dropout) {
            const uint32_t km = kmw[c];
#pragma unroll
            for (int i = 0; i < 32; ++i)   // dS = P (dP m - delta), m = keep / (1 - p): select, FMA, multiply
              ds[i] = pe[c * 32 + i] * fmaf(__uint_as_float(dv[i]), (km >> i) & 1u ? p.c.keep_scale : 0.f, -dl);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) ds[i] = pe[c * 32 + i] * (__uint_as_float(dv[i]) - dl);
          }
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(ds[2 * i], ds[2 * i + 1]);
          if (c == 0 && j > 0) {   // dS tile consumed by the previous dQ MMA (issued a whole block ago)
            mbar_wait(ds_free, (j - 1) & 1);
            tc_fence_after();
          }
          store_operand_chunk(sdS_u32, row, c, w);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(ds_ready);
      PROF_MARK(3);
    }
    Vec64 k0;
    load_vec64(k0, p.c.k + (int64_t)b * p.c.qkv_bs + h * HD);   // in flight under the last dQ MMA
    mbar_wait(ds_free, (nblk - 1) & 1);  // last dQ MMA retired
    tc_fence_after();
    PROF_MARK(4);
    {  // tcgen05.ld is warp-collective (.sync.aligned): every lane loads, only valid rows store
      bf16* dst = p.dq + (int64_t)b * p.d_bs + (int64_t)token * p.d_rs + h * HD;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tdQ + lane_base + c * 32, v);
        tmem_ld_wait();
        axpy_half(v, ds_x, k0, c);   // + dS_{i,0} k_0
        if (qrow < n) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 t;
            t.x = pack_bf16x2(__uint_as_float(v[8 * i]) * p.c.scale, __uint_as_float(v[8 * i + 1]) * p.c.scale);
            t.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * p.c.scale, __uint_as_float(v[8 * i + 3]) * p.c.scale);
            t.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * p.c.scale, __uint_as_float(v[8 * i + 5]) * p.c.scale);
            t.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * p.c.scale, __uint_as_float(v[8 * i + 7]) * p.c.scale);
            *reinterpret_cast<uint4*>(dst + c * 32 + 8 * i) = t;
          }
        }
      }
    }
    PROF_MARK(5);
#ifdef NV_PROFILE
    if (role.tile == 1 && h == 3 && (b == 5 || b == 40) && (threadIdx.x == 0 || threadIdx.x == 70))
      PROF_DUMP(4 + (b == 40 ? 2 : 0) + (threadIdx.x == 70 ? 1 : 0));
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// =====================================================================================================
// backward, dK / dV: CTA = 128 keys (TMEM lane = key); loop over 64-query blocks
// =====================================================================================================
constexpr int DKV_SMEM_TILES = 2 * SLAB /*K, V*/ + 2 * 2 * BOX /*Q_i, dO_i x 2 stages*/ + 2 * SLAB /*P^T, dS^T*/;
constexpr int DKV_SMEM = DKV_SMEM_TILES + 128 + 4 * 768 /*per-warp lse/delta/mask staging*/ + 4 * 256 /*per-warp q_0 / dO_0*/ + 1024;

// When a block's P^T / dS^T are ready the MMA warp issues the NEXT block's S^T / dP^T before this block's dV / dK
// accumulation (the rows wait for 8 MMAs instead of 16), and the rows take the "operand tiles free" barrier only right
// before their first store.
__global__ void __launch_bounds__(NTHREADS, 2)
attn_tc_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                       const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                       const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + SLAB;
  uint8_t* sQD = smem + 2 * SLAB;            // stage s: Q_i at s*2*BOX, dO_i at s*2*BOX + BOX
  uint8_t* sPT = smem + 2 * SLAB + 4 * BOX;
  uint8_t* sdST = sPT + SLAB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DKV_SMEM_TILES);
  uint64_t* kv_full = bars;         // 1
  uint64_t* qd_full = bars + 1;     // 2
  uint64_t* qd_empty = bars + 3;    // 2
  uint64_t* s_full = bars + 5;      // 1
  uint64_t* pds_ready = bars + 6;   // 1
  uint64_t* pds_free = bars + 7;    // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.c.N;
  const int n = N - 1;  // tokens 1 .. n are tiled; tile-local index r <-> token r + 1
  const CtaRole role = cta_role((n + BQ - 1) / BQ, p.c.H, p.c.ntile_ctas);
  const int b = role.b, h = role.h, k0 = role.tile * BQ;
  if (role.simt) {  // dK / dV of key token 0, SIMT
    PROF_DECL
    bwd_extra_key_dkv(p, reinterpret_cast<float*>(smem), role.tile);
    PROF_MARK(0);
#ifdef NV_PROFILE
    if (role.tile == 14 && threadIdx.x == 0) PROF_DUMP(14);
#endif
    return;
  }
  const int nblk = (n + BWD_CB - 1) / BWD_CB;
  const int nactive = (min(BQ, n - k0) + 31) >> 5;

  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&qd_full[i], 1); mbar_init(&qd_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(pds_ready, 32 * nactive);
    mbar_init(pds_free, 1);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tST = tmem_base, tdPT = tmem_base + 64, tdV = tmem_base + 128, tdK = tmem_base + 192;

  if (warp == TMA_WARP) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * SLAB);
      load_rows(sK, &tk, kv_full, h * HD, 1 + k0, b, 2);
      load_rows(sV, &tv, kv_full, h * HD, 1 + k0, b, 2);
    }
    __syncwarp();
    for (int i = 0; i < nblk; ++i) {
      const int s = i & 1;
      mbar_wait(&qd_empty[s], ((i >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&qd_full[s], 2 * BOX);
        load_rows(sQD + s * 2 * BOX, &tq, &qd_full[s], h * HD, 1 + i * BWD_CB, b, 1);
        load_rows(sQD + s * 2 * BOX + BOX, &tdo, &qd_full[s], h * HD, 1 + i * BWD_CB, b, 1);
      }
      __syncwarp();
    }
  } else if (warp == MMA_WARP) {
    constexpr uint32_t idesc_acc = umma_idesc_bf16(128, HD, 0, 1);
    const uint64_t k_desc = kmajor_desc(sK), v_desc = kmajor_desc(sV);
    auto cols16 = [&](int i) { return (min(BWD_CB, n - i * BWD_CB) + 15) & ~15; };
    auto issue_scores = [&](int i) {
      const uint8_t* qt = sQD + (i & 1) * 2 * BOX;
      const uint32_t idesc = umma_idesc_bf16(128, cols16(i), 0, 0);
      mma_k64(tST, k_desc, kmajor_desc(qt), idesc);         // S^T  = K Q_i^T
      mma_k64(tdPT, v_desc, kmajor_desc(qt + BOX), idesc);   // dP^T = V dO_i^T
      umma_commit(s_full);
    };
    mbar_wait(kv_full, 0);
    mbar_wait(&qd_full[0], 0);
    tc_fence_after();
    if (elect_one()) issue_scores(0);
    __syncwarp();
    for (int i = 0; i < nblk; ++i) {
      const int s = i & 1;
      if (i + 1 < nblk) mbar_wait(&qd_full[(i + 1) & 1], ((i + 1) >> 1) & 1);
      mbar_wait(pds_ready, i & 1);
      tc_fence_after();
      if (elect_one()) {
        if (i + 1 < nblk) issue_scores(i + 1);
        const uint8_t* qt = sQD + s * 2 * BOX;
        mma_rows(tdV, sPT, qt + BOX, idesc_acc, cols16(i) >> 4, i > 0);  // dV += P^T  dO_i
        mma_rows(tdK, sdST, qt, idesc_acc, cols16(i) >> 4, i > 0);       // dK += dS^T Q_i
        umma_commit(&qd_empty[s]);
        umma_commit(pds_free);
      }
      __syncwarp();
    }
  } else if (warp < nactive) {
    const int row = warp * 32 + lane;
    const int krow = k0 + row;               // tile-local key index (== mask bit position); token = krow + 1
    const int ktoken = min(krow + 1, N - 1);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t sPT_u32 = smem_u32(sPT), sdST_u32 = smem_u32(sdST);
    const float cs = p.c.scale * LOG2E;
    const float* lse_bh = p.lse + ((int64_t)b * p.c.H + h) * N;
    // per-warp staging of the block's 64 per-query statistics: [lse*log2e | delta], read back as broadcasts
    float* stats = reinterpret_cast<float*>(smem + DKV_SMEM_TILES + 128) + warp * 192;
    uint32_t* mwords = reinterpret_cast<uint32_t*>(stats + 128);  // dropout: the 64 queries' mask words of this warp's keys
    const bool dropout = p.c.drop_thr != 0;
    const uint64_t mbase = ((uint64_t)b * p.c.H + h) * N;
    PROF_DECL
    // per-query statistics of a block: fetched one block ahead into registers (the global-load latency hides
    // behind the previous block's exponentials), then staged in shared memory and read back as broadcasts
    // The loaded values are NOT touched here (no scaling, no select): any arithmetic on them would park the thread on
    // the load right away and the one-block-ahead prefetch would hide nothing (it used to: 1.8 k cycles per block).
    float st_l[2], st_d[2];
    uint32_t st_m[2] = {0u, 0u};
    auto fetch_stats = [&](int blk) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int q = blk * BWD_CB + t * 32 + lane;   // tile-local query index; token = q + 1
        const int qc = min(q + 1, N - 1);
        st_l[t] = __ldg(lse_bh + qc);
        st_d[t] = p.delta[((int64_t)b * p.c.H + h) * N + qc];  // written by the dQ grid (plain load: not read-only data)
        if (dropout) st_m[t] = p.c.mask[(mbase + qc) * p.c.mask_words + ((k0 + warp * 32) >> 5)];
      }
    };
    const uint32_t stats_u32 = smem_u32(stats), mwords_u32 = smem_u32(mwords);
    // q and dO of token 0 (the query that stays outside the tiles) are needed after the loop: fetched now, straight
    // into this warp's shared-memory scratch by cp.async (no registers held, no latency left at the point of use)
    const uint32_t x0_u32 = smem_u32(smem + DKV_SMEM_TILES + 128 + 4 * 768) + (uint32_t)warp * 256u;
    if (lane < 16) {
      const bf16* src = (lane < 8 ? p.c.q + (int64_t)b * p.c.qkv_bs : p.dO + (int64_t)b * p.o_bs) + h * HD + 8 * (lane & 7);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(x0_u32 + (uint32_t)lane * 16u), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int64_t stat0 = ((int64_t)b * p.c.H + h) * N;
    const float lse_0 = __ldg(p.lse + stat0), dl_0 = p.delta[stat0];
    uint32_t kw_0 = 0xFFFFFFFFu;
    if (dropout) kw_0 = __ldg(p.c.mask + stat0 * p.c.mask_words + (min(krow, n - 1) >> 5));
    fetch_stats(0);
    for (int i = 0; i < nblk; ++i) {
      const int qb = i * BWD_CB;
      const int nvalid = min(BWD_CB, n - qb);
      const int nch = (nvalid + 31) >> 5;
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        // padded queries: lse = +inf makes P exactly 0 (their dO / Q rows are zero-filled anyway)
        st_l[t] = (qb + t * 32 + lane) < n ? st_l[t] * LOG2E : INFINITY;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(stats_u32 + (uint32_t)(t * 32 + lane) * 4), "f"(st_l[t]) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(stats_u32 + (uint32_t)(64 + t * 32 + lane) * 4), "f"(st_d[t]) : "memory");
        if (dropout)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(mwords_u32 + (uint32_t)(t * 32 + lane) * 4), "r"(st_m[t]) : "memory");
      }
      if (i + 1 < nblk) fetch_stats(i + 1);
      __syncwarp();
      PROF_MARK(i == 0 ? 7 : 0);
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      PROF_MARK(1);
      PROF_MARK(2);
      for (int c = 0; c < nch; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tST + lane_base + c * 32, sv);
        tmem_ld_32x32(tdPT + lane_base + c * 32, dv);
        tmem_ld_wait();
        float pt[32], ds[32];
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          float4 l2, dl;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(l2.x), "=f"(l2.y), "=f"(l2.z), "=f"(l2.w)
                       : "r"(stats_u32 + (uint32_t)(c * 32 + e4 * 4) * 4));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dl.x), "=f"(dl.y), "=f"(dl.z), "=f"(dl.w)
                       : "r"(stats_u32 + (uint32_t)(64 + c * 32 + e4 * 4) * 4));
          const float l2v[4] = {l2.x, l2.y, l2.z, l2.w}, dlv[4] = {dl.x, dl.y, dl.z, dl.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = e4 * 4 + k;
            const float pe = ex2(fmaf(__uint_as_float(sv[e]), cs, -l2v[k]));
            pt[e] = pe;
            ds[e] = pe * (__uint_as_float(dv[e]) - dlv[k]);
          }
        }
        if (dropout) {
          uint32_t mw[32];  // the chunk's 32 queries' mask words for this warp's 32 keys (explicit shared loads)
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(mw[4 * e4]), "=r"(mw[4 * e4 + 1]),
                         "=r"(mw[4 * e4 + 2]), "=r"(mw[4 * e4 + 3]) : "r"(mwords_u32 + (uint32_t)(c * 32 + e4 * 4) * 4));
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const bool keep = (mw[e] >> lane) & 1u;
            const float pe = pt[e];
            // ds was pe * (dp - dl); with the mask it is pe * (keep ? dp * ks : 0 - dl)
            const float dl_pe = fmaf(-pe, __uint_as_float(dv[e]), ds[e]);  // = -pe * dl
            ds[e] = keep ? fmaf(pe * p.c.keep_scale, __uint_as_float(dv[e]), dl_pe) : dl_pe;
            pt[e] = keep ? pe * p.c.keep_scale : 0.f;
          }
        }
        uint32_t w[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) w[e] = pack_bf16x2(pt[2 * e], pt[2 * e + 1]);
        if (c == 0 && i > 0) {   // the previous dV / dK MMAs (issued behind this block's scores) are done
          mbar_wait(pds_free, (i - 1) & 1);
          tc_fence_after();
        }
        store_operand_chunk(sPT_u32, row, c, w);
#pragma unroll
        for (int e = 0; e < 16; ++e) w[e] = pack_bf16x2(ds[2 * e], ds[2 * e + 1]);
        store_operand_chunk(sdST_u32, row, c, w);
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(pds_ready);
      PROF_MARK(3);
    }
    // query token 0 stays outside the tiles: its row of P / dS for this key, from the key's K / V rows in
    // shared memory (rank-1 terms p~_0j dO_0 and dS_0j q_0 added to the accumulators below)
    float pm_x, ds_x;
    mbar_wait(kv_full, 0);  // long since complete: orders this thread's reads of the TMA-written K / V tiles
    Vec64 q0v, do0;   // q and dO of token 0: used for the scores here and for the rank-1 terms below
    {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q0v.c[i].x), "=r"(q0v.c[i].y), "=r"(q0v.c[i].z), "=r"(q0v.c[i].w)
                     : "r"(x0_u32 + (uint32_t)i * 16u));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(do0.c[i].x), "=r"(do0.c[i].y), "=r"(do0.c[i].z), "=r"(do0.c[i].w)
                     : "r"(x0_u32 + 128u + (uint32_t)i * 16u));
      }
      const uint32_t kw = kw_0;
      const float lse2_0 = lse_0 * LOG2E;
      const float p_x = ex2(fmaf(row_dot(smem_u32(sK), row, q0v), cs, -lse2_0));
      const float dp_x = row_dot(smem_u32(sV), row, do0);
      const bool keep = (kw >> (krow & 31)) & 1u;
      pm_x = keep ? p_x * p.c.keep_scale : 0.f;
      ds_x = p_x * ((keep ? dp_x * p.c.keep_scale : 0.f) - dl_0);
    }
    PROF_MARK(6);
    mbar_wait(pds_free, (nblk - 1) & 1);
    tc_fence_after();
    PROF_MARK(4);
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      bf16* base = which == 0 ? p.dv : p.dk;
      const float mul = which == 0 ? 1.0f : p.c.scale;
      bf16* dst = base + (int64_t)b * p.d_bs + (int64_t)ktoken * p.d_rs + h * HD;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32((which == 0 ? tdV : tdK) + lane_base + c * 32, v);
        tmem_ld_wait();
        axpy_half(v, which == 0 ? pm_x : ds_x, which == 0 ? do0 : q0v, c);
        if (krow < n) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 t;
            t.x = pack_bf16x2(__uint_as_float(v[8 * i]) * mul, __uint_as_float(v[8 * i + 1]) * mul);
            t.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * mul, __uint_as_float(v[8 * i + 3]) * mul);
            t.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * mul, __uint_as_float(v[8 * i + 5]) * mul);
            t.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * mul, __uint_as_float(v[8 * i + 7]) * mul);
            *reinterpret_cast<uint4*>(dst + c * 32 + 8 * i) = t;
          }
        }
      }
    }
    PROF_MARK(5);
#ifdef NV_PROFILE
    if (role.tile == 1 && h == 3 && (b == 5 || b == 40) && (threadIdx.x == 0 || threadIdx.x == 70))
      PROF_DUMP(8 + (b == 40 ? 2 : 0) + (threadIdx.x == 70 ? 1 : 0));
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// Attention backward when only the cls query (token 0) of every sample carries gradient — the last block under a
// cls-pooled head (vit_3d.py:123): dS has a single non-zero row, so dK and dV are rank-1 in that row and dQ is one
// row: O(N d) per (batch, head) instead of O(N^2 d). One CTA per (head, batch); eight lanes share a key, each owning
// one 16-byte piece of its 128-byte k / v / dk / dv / dq rows, so a warp instruction moves four whole rows.
constexpr int CLS_THREADS = 256;
__global__ void __launch_bounds__(CLS_THREADS)
attn_cls_bwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, int64_t qkv_bs,
                    int64_t qkv_rs, const bf16* __restrict__ o, int64_t o_bs, const bf16* __restrict__ dO_cls,
                    int64_t do_bs, const float* __restrict__ lse, bf16* __restrict__ dq, bf16* __restrict__ dk,
                    bf16* __restrict__ dv, int64_t d_bs, int64_t d_rs, int N, int H, float scale,
                    const uint32_t* __restrict__ mask, int mask_words, float keep_scale) {
  static_assert(HD == 64, "eight lanes x eight dims per key row");
  __shared__ float red[CLS_THREADS / 32][HD];
  const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const int sub = t & 7;        // which 8 dims of the head this lane owns
  const int grp = t >> 3;       // key slot within the CTA's stride
  // this lane's slice of the cls query, its dO and O (delta = dO . O over the whole head: 8-lane reduction)
  float qc[8], doc[8];
  float delta = 0.f;
  {
    const uint4 qq = *reinterpret_cast<const uint4*>(q + (int64_t)b * qkv_bs + h * HD + 8 * sub);   // token 0
    const uint4 dd = *reinterpret_cast<const uint4*>(dO_cls + (int64_t)b * do_bs + h * HD + 8 * sub);
    const uint4 oo = *reinterpret_cast<const uint4*>(o + (int64_t)b * o_bs + h * HD + 8 * sub);
    const uint32_t qw[4] = {qq.x, qq.y, qq.z, qq.w}, dw[4] = {dd.x, dd.y, dd.z, dd.w}, ow[4] = {oo.x, oo.y, oo.z, oo.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = unpack_bf16x2(qw[j]), c = unpack_bf16x2(dw[j]), e = unpack_bf16x2(ow[j]);
      qc[2 * j] = a.x; qc[2 * j + 1] = a.y;
      doc[2 * j] = c.x; doc[2 * j + 1] = c.y;
      delta = fmaf(c.x, e.x, fmaf(c.y, e.y, delta));
    }
    delta += __shfl_xor_sync(0xffffffffu, delta, 1);
    delta += __shfl_xor_sync(0xffffffffu, delta, 2);
    delta += __shfl_xor_sync(0xffffffffu, delta, 4);
  }
  const int64_t bh = (int64_t)b * H + h;
  const float lse2 = lse[bh * N] * LOG2E;
  float dq_acc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) dq_acc[d] = 0.f;
  for (int key0 = 0; key0 < N; key0 += CLS_THREADS / 8) {   // uniform trip count: the shuffles below need whole warps
    const int key = key0 + grp;
    const bool live = key < N;
    const int64_t row = (int64_t)(live ? key : 0) * qkv_rs + h * HD + 8 * sub;
    const uint4 kk = *reinterpret_cast<const uint4*>(k + (int64_t)b * qkv_bs + row);
    const uint4 vv = *reinterpret_cast<const uint4*>(v + (int64_t)b * qkv_bs + row);
    const uint32_t kw[4] = {kk.x, kk.y, kk.z, kk.w}, vw[4] = {vv.x, vv.y, vv.z, vv.w};
    float kf[8];
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = unpack_bf16x2(kw[j]), c2 = unpack_bf16x2(vw[j]);
      kf[2 * j] = a.x; kf[2 * j + 1] = a.y;
      s = fmaf(qc[2 * j], a.x, fmaf(qc[2 * j + 1], a.y, s));
      dp = fmaf(doc[2 * j], c2.x, fmaf(doc[2 * j + 1], c2.y, dp));
    }
#pragma unroll
    for (int m = 1; m <= 4; m <<= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, m);
      dp += __shfl_xor_sync(0xffffffffu, dp, m);
    }
    const float p = ex2(fmaf(s, scale * LOG2E, -lse2));
    bool keep = true;
    if (mask != nullptr && live) {  // row of token 0; bit position of key token `key`
      const int pos = key_pos(key, N);
      keep = (mask[bh * N * mask_words + (pos >> 5)] >> (pos & 31)) & 1u;
    }
    const float pm = keep ? p * keep_scale : 0.f;
    const float ds = live ? p * ((keep ? dp * keep_scale : 0.f) - delta) * scale : 0.f;
    if (live) {
      const int64_t drow = (int64_t)b * d_bs + (int64_t)key * d_rs + h * HD + 8 * sub;
      uint4 a, bb;
      a.x = pack_bf16x2(ds * qc[0], ds * qc[1]); a.y = pack_bf16x2(ds * qc[2], ds * qc[3]);
      a.z = pack_bf16x2(ds * qc[4], ds * qc[5]); a.w = pack_bf16x2(ds * qc[6], ds * qc[7]);
      bb.x = pack_bf16x2(pm * doc[0], pm * doc[1]); bb.y = pack_bf16x2(pm * doc[2], pm * doc[3]);
      bb.z = pack_bf16x2(pm * doc[4], pm * doc[5]); bb.w = pack_bf16x2(pm * doc[6], pm * doc[7]);
      *reinterpret_cast<uint4*>(dk + drow) = a;
      *reinterpret_cast<uint4*>(dv + drow) = bb;
      if (key != 0) *reinterpret_cast<uint4*>(dq + drow) = make_uint4(0u, 0u, 0u, 0u);  // queries other than the cls token
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) dq_acc[d] = fmaf(ds, kf[d], dq_acc[d]);
  }
  // dQ of the cls row: fold the four key slots of a warp (lanes with equal `sub`), then the warps
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    float r = dq_acc[d];
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 16);
    if ((t & 31) < 8) red[t >> 5][8 * sub + d] = r;
  }
  __syncthreads();
  if (t < HD) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < CLS_THREADS / 32; ++w) r += red[w][t];
    dq[(int64_t)b * d_bs + h * HD + t] = __float2bfloat16(r);
  }
}

int make_map(CUtensorMap* m, const bf16* base, int64_t bs, int64_t rs, int B, int N, int H, int box_rows = 64) {
  const uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)N, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)rs * 2, (uint64_t)bs * 2};
  const uint32_t box[3] = {64, (uint32_t)box_rows, 1};
  return nv_encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int fill_common(Common& c, const bf16* q, const bf16* k, const bf16* v, int64_t qkv_bs, int64_t qkv_rs, int N, int H,
                float scale, float dropout_p, uint64_t seed, uint32_t* mask, int mask_ready = 0) {
  NV_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "attention: dropout_p %f out of range [0, 1)", dropout_p);
  NV_REQUIRE(N <= 16384, "attention: N=%d tokens exceed the SIMT cls-row staging (16384)", N);
  c.q = q; c.k = k; c.v = v; c.qkv_bs = qkv_bs; c.qkv_rs = qkv_rs;
  c.N = N; c.H = H; c.scale = scale; c.seed = seed;
  c.ntile_ctas = 0;  // set by the launchers (tile_grid)
  c.epoch = dropout_p > 0.f ? nv_rng_epoch_dev() : nullptr;
  c.drop_thr = nv_dropout_threshold(dropout_p);
  c.keep_scale = nv_dropout_keep_scale(c.drop_thr);
  c.mask = mask;
  c.mask_words = (N + 31) / 32;
  c.mask_ready = mask_ready;
  NV_REQUIRE(c.drop_thr == 0 || mask != nullptr, "attention: dropout needs the mask buffer [B*H, N, ceil(N/32)] uint32");
  return NV_OK;
}

int check_args(const void* p, int64_t bs, int64_t rs, const char* name) {
  NV_REQUIRE(p != nullptr, "attention: %s is null", name);
  NV_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0 && bs % 8 == 0 && rs % 8 == 0,
             "attention: %s must be 16-byte aligned with strides that are multiples of 8 elements", name);
  return NV_OK;
}

// 1-D grid of the tile kernels: tiles over tokens 1..N-1 for every (head, batch), then one SIMT CTA per simt_groups
// (head, batch) pairs for token 0. NV_ATTN_X=skipsimt drops the SIMT CTAs (timing experiments only: token 0 is then not computed).
int tile_grid(Common& c, int B, unsigned* grid) {
  static const char* x = getenv("NV_ATTN_X");
  const int64_t ntile = (int64_t)((c.N - 1 + BQ - 1) / BQ) * c.H * B;
  c.simt_groups = simt_groups_for(c.N);
  const int64_t ncta = ntile + ((x && strstr(x, "skipsimt")) ? 0 : ((int64_t)c.H * B + c.simt_groups - 1) / c.simt_groups);
  c.B = B;
  NV_REQUIRE(ncta < (1ll << 31) && ncta > 0, "attention: bad grid (%lld CTAs)", (long long)ncta);
  c.ntile_ctas = (int)ntile;
  *grid = (unsigned)ncta;
  return NV_OK;
}

// NV_ATTN_SYNC=1: synchronise and report after every attention kernel (debugging aid for the probes)
int dbg_sync(const char* what, cudaStream_t stream) {
  static const bool on = getenv("NV_ATTN_SYNC") != nullptr;
  if (!on) return NV_OK;
  cudaError_t e = cudaStreamSynchronize(stream);
  fprintf(stderr, "[attn] %s: %s\n", what, cudaGetErrorString(e));
  fflush(stderr);
  return nv_check_cuda(e, what);
}

template <typename K>
int set_smem(K kern, int bytes) {
  NV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  NV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  return NV_OK;
}

int fwd_attrs_once() {
  static uint64_t attr_done = 0;  // bit per device
  int s;
  if (nv_first_on_device(&attr_done)) {
    if ((s = set_smem(attn_tc_fwd_kernel<false, false>, FWD_SMEM)) != NV_OK) return s;
    if ((s = set_smem(attn_tc_fwd_kernel<true, false>, FWD_SMEM)) != NV_OK) return s;
    if ((s = set_smem(attn_tc_fwd_kernel<true, true>, FWD_SMEM)) != NV_OK) return s;
  }
  return NV_OK;
}

}  // namespace

int nv_attn_tc_fwd_launch(const bf16* q, const bf16* k, const bf16* v, int64_t qkv_bs, int64_t qkv_rs, bf16* o,
                          int64_t o_bs, int64_t o_rs, float* lse, int B, int N, int H, int head_dim, float scale,
                          float dropout_p, uint64_t seed, uint32_t* drop_mask, int mask_ready, cudaStream_t stream) {
  NV_REQUIRE(head_dim == HD, "attention: head_dim %d unsupported by the bf16 flash kernel (needs 64)", head_dim);
  NV_REQUIRE(B >= 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "attention: bad sizes B=%d N=%d H=%d", B, N, H);
  if (B == 0) return NV_OK;
  int s;
  if ((s = check_args(q, qkv_bs, qkv_rs, "q")) != NV_OK) return s;
  if ((s = check_args(k, qkv_bs, qkv_rs, "k")) != NV_OK) return s;
  if ((s = check_args(v, qkv_bs, qkv_rs, "v")) != NV_OK) return s;
  if ((s = check_args(o, o_bs, o_rs, "o")) != NV_OK) return s;
  NV_REQUIRE(lse != nullptr, "attention: lse is null");
  CUtensorMap tq, tk, tv;
  if ((s = make_map(&tq, q, qkv_bs, qkv_rs, B, N, H, BQ)) != NV_OK) return s;
  if ((s = make_map(&tk, k, qkv_bs, qkv_rs, B, N, H, FWD_KB)) != NV_OK) return s;
  if ((s = make_map(&tv, v, qkv_bs, qkv_rs, B, N, H, FWD_KB)) != NV_OK) return s;
  FwdParams p;
  if ((s = fill_common(p.c, q, k, v, qkv_bs, qkv_rs, N, H, scale, dropout_p, seed, drop_mask, mask_ready)) != NV_OK) return s;
  p.o = o; p.o_bs = o_bs; p.o_rs = o_rs; p.lse = lse;
  if ((s = fwd_attrs_once()) != NV_OK) return s;
  unsigned ncta;
  if ((s = tile_grid(p.c, B, &ncta)) != NV_OK) return s;
  dim3 grid(ncta);
  if (p.c.drop_thr != 0 && p.c.mask_ready) attn_tc_fwd_kernel<true, true><<<grid, NTHREADS, FWD_SMEM, stream>>>(tq, tk, tv, p);
  else if (p.c.drop_thr != 0) attn_tc_fwd_kernel<true, false><<<grid, NTHREADS, FWD_SMEM, stream>>>(tq, tk, tv, p);
  else attn_tc_fwd_kernel<false, false><<<grid, NTHREADS, FWD_SMEM, stream>>>(tq, tk, tv, p);
  NV_LAUNCH_CHECK("attn_tc_fwd_kernel");
  return dbg_sync("fwd", stream);
}

// Forward for the cls query only (token 0 of every sample): the last block under a cls-pooled head (vit_3d.py:123)
// uses nothing else of its attention output. Launches the forward kernel with zero tile CTAs, i.e. only the SIMT CTAs
// that serve token 0 in the full forward: same arithmetic, same dropout bits (row (b*H+h, token 0) of drop_mask is
// drawn inline and written, or read when mask_ready), o row b written at o + b*o_bs, lse at lse[(b*H+h)*N].
int nv_attn_cls_fwd_launch(const bf16* q, const bf16* k, const bf16* v, int64_t qkv_bs, int64_t qkv_rs, bf16* o,
                           int64_t o_bs, float* lse, int B, int N, int H, int head_dim, float scale, float dropout_p,
                           uint64_t seed, uint32_t* drop_mask, int mask_ready, cudaStream_t stream) {
  NV_REQUIRE(head_dim == HD, "attention: head_dim %d unsupported by the bf16 flash kernel (needs 64)", head_dim);
  NV_REQUIRE(B >= 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "attention: bad sizes B=%d N=%d H=%d", B, N, H);
  if (B == 0) return NV_OK;
  int s;
  if ((s = check_args(q, qkv_bs, qkv_rs, "q")) != NV_OK) return s;
  if ((s = check_args(k, qkv_bs, qkv_rs, "k")) != NV_OK) return s;
  if ((s = check_args(v, qkv_bs, qkv_rs, "v")) != NV_OK) return s;
  NV_REQUIRE(o != nullptr && lse != nullptr, "attention: o / lse is null");
  FwdParams p;
  if ((s = fill_common(p.c, q, k, v, qkv_bs, qkv_rs, N, H, scale, dropout_p, seed, drop_mask, mask_ready)) != NV_OK) return s;
  p.o = o; p.o_bs = o_bs; p.o_rs = 0; p.lse = lse;
  if ((s = fwd_attrs_once()) != NV_OK) return s;
  p.c.B = B;
  p.c.ntile_ctas = 0;   // every CTA takes the SIMT role (cta_role)
  p.c.simt_groups = simt_groups_for(N);
  const dim3 grid((unsigned)(((int64_t)H * B + p.c.simt_groups - 1) / p.c.simt_groups));
  CUtensorMap none;     // the SIMT role returns before any tensor map is touched
  memset(&none, 0, sizeof(none));
  if (p.c.drop_thr != 0) attn_tc_fwd_kernel<true, false><<<grid, NTHREADS, FWD_SMEM, stream>>>(none, none, none, p);
  else attn_tc_fwd_kernel<false, false><<<grid, NTHREADS, FWD_SMEM, stream>>>(none, none, none, p);
  NV_LAUNCH_CHECK("attn_tc_fwd_kernel (cls only)");
  return dbg_sync("fwd cls", stream);
}

int nv_attn_tc_bwd_launch(const bf16* q, const bf16* k, const bf16* v, int64_t qkv_bs, int64_t qkv_rs, const bf16* o,
                          const bf16* dO, int64_t o_bs, int64_t o_rs, const float* lse, float* delta_ws, bf16* dq,
                          bf16* dk, bf16* dv, int64_t d_bs, int64_t d_rs, int B, int N, int H, int head_dim,
                          float scale, float dropout_p, const uint32_t* drop_mask, cudaStream_t stream) {
  NV_REQUIRE(head_dim == HD, "attention: head_dim %d unsupported by the bf16 flash kernel (needs 64)", head_dim);
  NV_REQUIRE(B >= 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "attention: bad sizes B=%d N=%d H=%d", B, N, H);
  if (B == 0) return NV_OK;
  int s;
  if ((s = check_args(q, qkv_bs, qkv_rs, "q")) != NV_OK) return s;
  if ((s = check_args(k, qkv_bs, qkv_rs, "k")) != NV_OK) return s;
  if ((s = check_args(v, qkv_bs, qkv_rs, "v")) != NV_OK) return s;
  if ((s = check_args(o, o_bs, o_rs, "o")) != NV_OK) return s;
  if ((s = check_args(dO, o_bs, o_rs, "dO")) != NV_OK) return s;
  if ((s = check_args(dq, d_bs, d_rs, "dq")) != NV_OK) return s;
  if ((s = check_args(dk, d_bs, d_rs, "dk")) != NV_OK) return s;
  if ((s = check_args(dv, d_bs, d_rs, "dv")) != NV_OK) return s;
  NV_REQUIRE(lse != nullptr && delta_ws != nullptr, "attention: lse / delta workspace is null");
  CUtensorMap tq, tk, tv, tdo;
  if ((s = make_map(&tq, q, qkv_bs, qkv_rs, B, N, H)) != NV_OK) return s;
  if ((s = make_map(&tk, k, qkv_bs, qkv_rs, B, N, H)) != NV_OK) return s;
  if ((s = make_map(&tv, v, qkv_bs, qkv_rs, B, N, H)) != NV_OK) return s;
  if ((s = make_map(&tdo, dO, o_bs, o_rs, B, N, H)) != NV_OK) return s;
  CUtensorMap to;
  if ((s = make_map(&to, o, o_bs, o_rs, B, N, H)) != NV_OK) return s;
  BwdParams p;
  if ((s = fill_common(p.c, q, k, v, qkv_bs, qkv_rs, N, H, scale, dropout_p, 0, const_cast<uint32_t*>(drop_mask))) != NV_OK) return s;
  p.lse = lse; p.delta = delta_ws; p.dq = dq; p.dk = dk; p.dv = dv; p.d_bs = d_bs; p.d_rs = d_rs;
  p.o = o; p.dO = dO; p.o_bs = o_bs; p.o_rs = o_rs;
  static uint64_t attr_done = 0;  // bit per device
  if (nv_first_on_device(&attr_done)) {
    if ((s = set_smem(attn_tc_bwd_dq_kernel, DQ_SMEM)) != NV_OK) return s;
    if ((s = set_smem(attn_tc_bwd_dkv_kernel, DKV_SMEM)) != NV_OK) return s;
  }
  // dQ first (it also writes delta, which the dK/dV grid reads); the last x-slot of each grid is the SIMT CTA
  unsigned ncta;
  if ((s = tile_grid(p.c, B, &ncta)) != NV_OK) return s;
  dim3 grid(ncta);
  static const char* only = getenv("NV_ATTN_ONLY");   // timing experiments: launch one of the two kernels only
  if (!only || !strcmp(only, "dq")) attn_tc_bwd_dq_kernel<<<grid, NTHREADS, DQ_SMEM, stream>>>(tq, tk, tv, tdo, to, p);
  NV_LAUNCH_CHECK("attn_tc_bwd_dq_kernel");
  { int ds = dbg_sync("dq", stream); if (ds != NV_OK) return ds; }
  if (!only || !strcmp(only, "dkv")) attn_tc_bwd_dkv_kernel<<<grid, NTHREADS, DKV_SMEM, stream>>>(tq, tk, tv, tdo, p);
  NV_LAUNCH_CHECK("attn_tc_bwd_dkv_kernel");
  return dbg_sync("dkv", stream);
}

int nv_attn_cls_bwd_launch(const bf16* q, const bf16* k, const bf16* v, int64_t qkv_bs, int64_t qkv_rs, const bf16* o,
                           int64_t o_bs, const bf16* dO_cls, int64_t do_bs, const float* lse, bf16* dq, bf16* dk, bf16* dv,
                           int64_t d_bs, int64_t d_rs, int B, int N, int H, int head_dim, float scale, float dropout_p,
                           const uint32_t* drop_mask, cudaStream_t stream) {
  NV_REQUIRE(head_dim == HD, "attention: head_dim %d unsupported by the bf16 kernels (needs 64)", head_dim);
  NV_REQUIRE(B >= 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "attention: bad sizes B=%d N=%d H=%d", B, N, H);
  if (B == 0) return NV_OK;
  int s;
  if ((s = check_args(q, qkv_bs, qkv_rs, "q")) != NV_OK) return s;
  if ((s = check_args(k, qkv_bs, qkv_rs, "k")) != NV_OK) return s;
  if ((s = check_args(v, qkv_bs, qkv_rs, "v")) != NV_OK) return s;
  if ((s = check_args(dq, d_bs, d_rs, "dq")) != NV_OK) return s;
  if ((s = check_args(dk, d_bs, d_rs, "dk")) != NV_OK) return s;
  if ((s = check_args(dv, d_bs, d_rs, "dv")) != NV_OK) return s;
  NV_REQUIRE(o != nullptr && dO_cls != nullptr && lse != nullptr, "attention: null o / dO / lse");
  NV_REQUIRE(((reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(dO_cls)) & 15) == 0 && o_bs % 8 == 0 &&
                 do_bs % 8 == 0, "attention: o / dO of the cls row must be 16-byte aligned (16-byte vector loads)");
  Common c;
  if ((s = fill_common(c, q, k, v, qkv_bs, qkv_rs, N, H, scale, dropout_p, 0, const_cast<uint32_t*>(drop_mask))) != NV_OK) return s;
  attn_cls_bwd_kernel<<<dim3(H, B), CLS_THREADS, 0, stream>>>(q, k, v, qkv_bs, qkv_rs, o, o_bs, dO_cls, do_bs, lse, dq, dk, dv, d_bs,
                                                      d_rs, N, H, scale, c.drop_thr != 0 ? c.mask : nullptr, c.mask_words,
                                                      c.keep_scale);
  NV_LAUNCH_CHECK("attn_cls_bwd_kernel");
  return NV_OK;
}
