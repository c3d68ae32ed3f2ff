// tcgen05 GEMM for the ViT3D linear layers (reference: src/models/vit_3d.py:18-22,41-45,50,60,94 —
// every nn.Linear on the hot path, forward, dgrad and wgrad).
//
//   C[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )      bf16 operands, fp32 accumulate in TMEM
//
// Operand storage (row-major, leading dimension in elements):
//   A K-major : A[M, K]          A MN-major : A stored as [K, M]   (wgrad: dY^T without a transpose pass)
//   B K-major : B[N, K]          B MN-major : B stored as [K, N]   (wgrad: X, dgrad: W)
//
// Structure (persistent, one CTA per SM, 320 threads):
//   warps 0-7  epilogue: tcgen05.ld -> regs -> swizzled smem transpose -> coalesced global I/O, one 32x32 chunk
//              at a time in three straight-line passes (8 LDS, arithmetic, 8 predicated STG); 16 warps in the
//              GELU modes. Store mode 3 stages the fp32 residual through a cp.async ring (see epilogue_chunk)
//   warp  8    TMA producer (one lane): 128B-swizzled tiles into a STAGES-deep mbarrier ring
//   warp  9    TMEM allocator + MMA issuer (one lane): tcgen05.mma kind::f16, K=16 per instruction
// The accumulator is double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
// main loop of tile i+1. Split-K work units accumulate with red.global.add.f32 straight into the fp32
// gradient buffer (wgrad).
//
// CG = 1: one CTA computes a 128 x BLOCK_N tile (tcgen05.mma.cta_group::1, M=128).
// CG = 2: a CTA pair (cluster of 2, same TPC) computes a 256 x BLOCK_N tile with tcgen05.mma.cta_group::2
//         (M=256): each CTA stages its own 128 rows of A and HALF of the B tile, the leader CTA issues the
//         MMA for both, each CTA's TMEM holds its 128 accumulator rows. Per-SM L2->smem traffic drops by a
//         third (the 1-CTA 128x256 tile is L2-bandwidth bound on B200: 96 B/clk/SM x 148 SMs > L2's ~6.3 KB/clk).
#include "nv_common.cuh"
#include "nv_rng.cuh"
#include <cstdlib>

namespace {

constexpr int BLOCK_M = 128;  // accumulator rows per CTA (TMEM lanes)
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
// epilogue warps: 8 for the store epilogue (memory-bound: residual prefetch is double-buffered in registers),
// 16 for the two GELU epilogues (instruction-bound: four warps per SM sub-partition hide the MUFU/FMA chains)
__host__ __device__ constexpr int epi_warps(int epi_mode) { return (epi_mode == 0 || epi_mode == 3) ? 8 : 16; }
// 4 KB shared-memory slabs per epilogue warp: the transpose slab, plus a two-deep residual ring in mode 3
__host__ __device__ constexpr int epi_slabs(int epi_mode) { return epi_mode == 3 ? 3 : 1; }
__host__ __device__ constexpr int num_threads(int epi_mode) { return (epi_warps(epi_mode) + 2) * 32; }
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;          // per epilogue warp
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;       // shared::cluster address of the pair's leader CTA

enum : int { EPI_GELU = 1, EPI_ATOMIC = 2 };

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, k_splits, k_blocks_total, k_blocks_per_split;
  const float* bias;      // [N] or null
  const float* residual;  // [M, ld_res] fp32 or null (added after activation)
  const bf16* gelu_u;     // [M, ld_u] or null: acc *= gelu'(u)  (dgrad through GELU)
  float* out_f32;         // [M, ld_f32] or null
  bf16* out_bf16;         // [M, ld_bf16] or null (final value, bf16 copy)
  bf16* out_pre;          // [M, ld_pre] or null (pre-activation, only with EPI_GELU)
  float* colsum;          // [N] or null: += sum over rows of the final value (bias gradient, red.add)
  int64_t ld_res, ld_u, ld_f32, ld_bf16, ld_pre;
  int flags;
  float alpha;
  // dropout on the value (nn.Dropout after to_out / GELU / the MLP down projection, vit_3d.py:21,23,45;
  // applied before the residual add). Mask = f(seed, stream, row * N + col): backward regenerates it.
  uint32_t drop_thr;  // 0 = off
  float keep_scale;
  uint64_t drop_seed;
  const uint64_t* drop_epoch;  // device epoch counter added to the seed (CUDA-graph replays)
  const uint8_t* drop_bits;    // optional pre-generated keep bits, byte (row * N + col) / 8 (nv_dropout_bits)
  int drop_row_mul;            // mask row = output row * drop_row_mul (compact cls-row problems index the full site)
  uint32_t drop_stream;
};

template <int BLOCK_N, int STAGES, int CG, int NUM_EPI_WARPS, int EPI_SLABS = 1>
struct SmemLayout {
  static constexpr int B_ROWS = BLOCK_N / CG;  // B rows staged by one CTA
  static constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = A_OFF + STAGES * A_STAGE_BYTES;
  static constexpr int EPI_OFF = B_OFF + STAGES * B_STAGE_BYTES;
  static constexpr int BAR_OFF = EPI_OFF + NUM_EPI_WARPS * EPI_SLABS * EPI_STAGE_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024B alignment
};

// ---- cluster / cta_group::2 PTX --------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes are signalled on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK),
      "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at the same smem offset in BOTH CTAs of the pair once the issued MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// ---- epilogue ---------------------------------------------------------------------------------------
// EPI_MODE: 0 = store (alpha, bias, residual, fp32/bf16 out, split-K red.add, column sums)
//           1 = bias + exact-erf GELU forward (writes pre-activation and activation, bf16)
//           2 = dgrad through GELU: acc * gelu'(u) (+ column sums)
// erf is evaluated with Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below bf16 resolution): one
// MUFU.RCP + one MUFU.EX2 + 7 FMAs; gelu' reuses the same exponential (exp(-u^2/2) is erf's exp(-x^2)).
// (MUFU.RCP / MUFU.EX2 are issued as the bare approx instructions: the IEEE-rounded intrinsics expand to
// range checks with slow-path calls that break the instruction-level parallelism of the unrolled epilogue.)
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void erf_parts(float u, float& erf_v, float& expv) {
  const float x = u * 0.70710678118654752440f;
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  expv = ex2_approx(-x * x * 1.4426950408889634f);  // exp(-x^2) = exp(-u^2/2)
  erf_v = copysignf(fmaf(-poly, expv, 1.0f), x);
}
__device__ __forceinline__ float gelu_fast(float u) {
  float e, ex;
  erf_parts(u, e, ex);
  return 0.5f * u * (1.0f + e);
}
__device__ __forceinline__ float gelu_grad_fast(float u) {
  float e, ex;
  erf_parts(u, e, ex);
  return fmaf(u * 0.39894228040143267794f, ex, 0.5f * (1.0f + e));
}

struct EpiAux {  // global operands of one 32x32 chunk, prefetched one chunk ahead
  float4 res[8];
  uint2 uu[8];
  float4 bias4;
  uint32_t kb[4];  // pre-drawn dropout keep bytes of this lane's four row slices (p.drop_bits)
};

template <int EPI_MODE, bool DROP = false>
__device__ __forceinline__ void epi_prefetch(const GemmParams& p, EpiAux& x, int lane, int row0, int col0) {
  // Out-of-range rows/columns are CLAMPED to a valid address instead of predicated: a "ok ? load : 0"
  // select makes the warp wait for the load right here and defeats the prefetch; clamped values are
  // simply never stored.
  const int sub_r = lane >> 3, sub_c = lane & 7;
  const int gn_raw = col0 + sub_c * 4;
  const bool col_ok = gn_raw < p.N;
  const int gn = col_ok ? gn_raw : 0;
  if (EPI_MODE == 0 && p.residual != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gm = min(row0 + i * 4 + sub_r, p.M - 1);
      x.res[i] = *reinterpret_cast<const float4*>(p.residual + (int64_t)gm * p.ld_res + gn);
    }
  }
  if (EPI_MODE == 2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gm = min(row0 + i * 4 + sub_r, p.M - 1);
      x.uu[i] = *reinterpret_cast<const uint2*>(p.gelu_u + (int64_t)gm * p.ld_u + gn);
    }
  }
  x.bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (EPI_MODE != 2 && p.bias != nullptr) x.bias4 = *reinterpret_cast<const float4*>(p.bias + gn);
  if (DROP && p.drop_bits != nullptr) {  // bytes drawn ahead by nv_dropout_bits: fetched with the other operands
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int gm_t = min(row0 + ((sub_c & 1) * 4 + t) * 4 + sub_r, p.M - 1);
      x.kb[t] = __ldg(p.drop_bits + (((uint64_t)gm_t * p.drop_row_mul * p.N + (gn & ~7)) >> 3));
    }
  }
}

// explicit shared-space accesses for the staging slab: through a generic pointer they compile to LD.E / ST.E
// (generic path: long-scoreboard latency and the local/global queue instead of the shared-memory pipe)
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// predicated global stores: written as one `@p st` so that the eight row stores of a chunk stay straight-line code
// (an `if (ok) *ptr = v` per row compiles to a branch + reconvergence block around every store)
__device__ __forceinline__ void stg_v4_if(bool ok, float* ptr, const float4& v) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %5, 0; @p st.global.v4.f32 [%0], {%1, %2, %3, %4}; }" ::"l"(ptr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)ok) : "memory");
}
__device__ __forceinline__ void red_v4_if(bool ok, float* ptr, const float4& v) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %5, 0; @p red.global.add.v4.f32 [%0], {%1, %2, %3, %4}; }" ::"l"(ptr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)ok) : "memory");
}
__device__ __forceinline__ void stg_v2_if(bool ok, void* ptr, uint32_t lo, uint32_t hi) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %3, 0; @p st.global.v2.b32 [%0], {%1, %2}; }" ::"l"(ptr), "r"(lo), "r"(hi),
               "r"((int)ok) : "memory");
}

// EPI_MODE 3 = mode 0 for the forward linears that add the fp32 residual stream (K-major operands, CTA pairs). Every
// global load of an epilogue warp completes on one hardware scoreboard, so a register prefetch can never have more
// than "everything issued so far" granularity and is one chunk deep at best. Here the residual chunk is copied by
// cp.async (completion counted per commit group) into a two-slab ring per warp, in the swizzled layout of the
// transposed accumulator: two chunks are in flight while a third is processed. Each lane copies exactly the
// 16-byte pieces it reads back itself, so its own wait_group is all the synchronisation needed.
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_PENDING> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory");
}
__device__ __forceinline__ void residual_prefetch(const GemmParams& p, uint32_t slab, int lane, int row0, int col0) {
  const int sub_r = lane >> 3, sub_c = lane & 7;
  const int gn_raw = col0 + sub_c * 4;
  const int gn = gn_raw < p.N ? gn_raw : 0;  // clamped like epi_prefetch: never stored
  const int gm_last = p.M - 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + sub_r;
    cp_async16(slab + (uint32_t)((r * 32 + ((sub_c ^ (r & 7)) << 2)) * 4),
               p.residual + (int64_t)min(row0 + r, gm_last) * p.ld_res + gn);
  }
  cp_async_commit();
}

// next_col0 >= 0 (mode 3): once this chunk's residual has been read out of `res_slab`, refill the slab with the
// residual of the chunk at column next_col0
template <int EPI_MODE, bool DROP>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&v)[32], uint32_t stage, int lane,
                                               int row0, int col0, const EpiAux& x, uint32_t res_slab = 0,
                                               int next_col0 = -1) {
  const int sub_r = lane >> 3, sub_c = lane & 7;
  const int gn = col0 + sub_c * 4;
  const bool col_ok = gn < p.N;
  // row-per-thread -> swizzled staging (16B chunk j of row `lane` lands at chunk j^(lane&7))
#pragma unroll
  for (int j = 0; j < 8; ++j)
    sts_v4(stage + (uint32_t)((lane * 32 + ((j ^ (lane & 7)) << 2)) * 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  // dropout keep-bits: one Philox call covers 8 consecutive columns = the two lanes (sub_c, sub_c ^ 1) of a row
  // slice, so each lane draws the bits of four of its eight row slices and fetches the rest from its neighbour
  uint32_t kb[4] = {0u, 0u, 0u, 0u};
  if (DROP) {
    if (p.drop_bits != nullptr) {  // bits drawn ahead of time by nv_dropout_bits (same values), prefetched
#pragma unroll
      for (int t = 0; t < 4; ++t) kb[t] = x.kb[t];
    } else {
      const uint64_t seed = nv_seed(p.drop_seed, p.drop_epoch);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int gm_t = row0 + ((sub_c & 1) * 4 + t) * 4 + sub_r;
        kb[t] = nv_keep_bits8(seed, ((uint64_t)gm_t * p.drop_row_mul * p.N + (gn & ~7)) >> 3, p.drop_stream, p.drop_thr);
      }
    }
  }
  // Three straight-line passes (read the slab, arithmetic, stores) instead of one row at a time: the per-row form
  // re-read the kernel parameters and branched on them inside every row, which left each warp waiting on a chain
  // of LDS -> LDC -> branch -> address math -> STG eight times per chunk with nothing else to issue.
  // G rows at a time (all eight under the 168-register budget of the store epilogue, four under the 96 of the
  // 16-warp GELU epilogues)
  constexpr bool STORE = EPI_MODE == 0 || EPI_MODE == 3;
  constexpr int G = STORE ? 8 : 4;
  const bool has_res = EPI_MODE == 0 && p.residual != nullptr;
  const int gm0 = row0 + sub_r;
  // rows this lane may store: gm0 + 4 i < M  <=>  i < rows_ok
  const int rows_ok = col_ok ? min(8, (p.M - gm0 + 3) >> 2) : 0;
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int g = 0; g < 8; g += G) {
    float4 a[G];
    float4 rr[EPI_MODE == 3 ? G : 1];
    uint2 pre[EPI_MODE == 1 ? G : 1];
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const int r = (g + k) * 4 + sub_r;
      a[k] = lds_v4(stage + (uint32_t)((r * 32 + ((sub_c ^ (r & 7)) << 2)) * 4));
      if (EPI_MODE == 3) rr[EPI_MODE == 3 ? k : 0] = lds_v4(res_slab + (uint32_t)((r * 32 + ((sub_c ^ (r & 7)) << 2)) * 4));
    }
    if (EPI_MODE == 3 && next_col0 >= 0) residual_prefetch(p, res_slab, lane, row0, next_col0);
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const int i = g + k;
      uint32_t keep4 = 0xFu;
      if (DROP) {
        const uint32_t other = __shfl_xor_sync(0xffffffffu, kb[i & 3], 1);
        const uint32_t b8 = ((sub_c & 1) == (i >> 2)) ? kb[i & 3] : other;
        keep4 = (b8 >> ((sub_c & 1) * 4)) & 0xFu;
      }
      float4 t = a[k];
      if (STORE) {
        t.x = fmaf(t.x, p.alpha, x.bias4.x);
        t.y = fmaf(t.y, p.alpha, x.bias4.y);
        t.z = fmaf(t.z, p.alpha, x.bias4.z);
        t.w = fmaf(t.w, p.alpha, x.bias4.w);
        if (DROP) t = nv_dropout4(t, keep4, p.keep_scale);
        if (EPI_MODE == 3) {
          const float4 r4 = rr[EPI_MODE == 3 ? k : 0];
          t.x += r4.x; t.y += r4.y; t.z += r4.z; t.w += r4.w;
        } else if (has_res) {
          t.x += x.res[i].x; t.y += x.res[i].y; t.z += x.res[i].z; t.w += x.res[i].w;
        }
      } else if (EPI_MODE == 1) {
        t.x += x.bias4.x; t.y += x.bias4.y; t.z += x.bias4.z; t.w += x.bias4.w;
        pre[EPI_MODE == 1 ? k : 0] = make_uint2(pack_bf16x2(t.x, t.y), pack_bf16x2(t.z, t.w));
        t.x = gelu_fast(t.x); t.y = gelu_fast(t.y); t.z = gelu_fast(t.z); t.w = gelu_fast(t.w);
        if (DROP) t = nv_dropout4(t, keep4, p.keep_scale);
      } else {
        const float2 u01 = unpack_bf16x2(x.uu[i].x);
        const float2 u23 = unpack_bf16x2(x.uu[i].y);
        t.x *= gelu_grad_fast(u01.x); t.y *= gelu_grad_fast(u01.y);
        t.z *= gelu_grad_fast(u23.x); t.w *= gelu_grad_fast(u23.y);
        if (DROP) t = nv_dropout4(t, keep4, p.keep_scale);  // d/du of dropout(gelu(u)): the activation's forward mask
      }
      a[k] = t;
    }
    if (EPI_MODE == 1 && p.out_pre != nullptr) {
      __nv_bfloat16* dst = p.out_pre + (int64_t)gm0 * p.ld_pre + gn;
      const int64_t step = (int64_t)4 * p.ld_pre;
#pragma unroll
      for (int k = 0; k < G; ++k)
        stg_v2_if(g + k < rows_ok, dst + (g + k) * step, pre[EPI_MODE == 1 ? k : 0].x, pre[EPI_MODE == 1 ? k : 0].y);
    }
    if (EPI_MODE == 0 && (p.flags & EPI_ATOMIC)) {
      float* dst = p.out_f32 + (int64_t)gm0 * p.ld_f32 + gn;
      const int64_t step = (int64_t)4 * p.ld_f32;
#pragma unroll
      for (int k = 0; k < G; ++k) red_v4_if(g + k < rows_ok, dst + (g + k) * step, a[k]);
    } else if (p.out_f32 != nullptr) {
      float* dst = p.out_f32 + (int64_t)gm0 * p.ld_f32 + gn;
      const int64_t step = (int64_t)4 * p.ld_f32;
#pragma unroll
      for (int k = 0; k < G; ++k) stg_v4_if(g + k < rows_ok, dst + (g + k) * step, a[k]);
    }
    if (p.out_bf16 != nullptr) {
      __nv_bfloat16* dst = p.out_bf16 + (int64_t)gm0 * p.ld_bf16 + gn;
      const int64_t step = (int64_t)4 * p.ld_bf16;
#pragma unroll
      for (int k = 0; k < G; ++k)
        stg_v2_if(g + k < rows_ok, dst + (g + k) * step, pack_bf16x2(a[k].x, a[k].y), pack_bf16x2(a[k].z, a[k].w));
    }
    if (EPI_MODE != 1 && p.colsum != nullptr) {
#pragma unroll
      for (int k = 0; k < G; ++k) {
        const bool ok = g + k < rows_ok;
        csum.x += ok ? a[k].x : 0.f; csum.y += ok ? a[k].y : 0.f;
        csum.z += ok ? a[k].z : 0.f; csum.w += ok ? a[k].w : 0.f;
      }
    }
  }
  if (EPI_MODE != 1 && p.colsum != nullptr) {  // warp-uniform: fold the 4 row groups, one vector red per chunk
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o);
      csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
      csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o);
      csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
    }
    if (sub_r == 0 && col_ok)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.colsum + gn), "f"(csum.x), "f"(csum.y),
                   "f"(csum.z), "f"(csum.w) : "memory");
  }
  __syncwarp();
}

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, int CG, int EPI_MODE, bool DROP>
__global__ void __launch_bounds__(num_threads(EPI_MODE), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmParams p) {
  constexpr int NUM_EPI_WARPS = epi_warps(EPI_MODE);
  constexpr int TMA_WARP = NUM_EPI_WARPS;
  constexpr int MMA_WARP = NUM_EPI_WARPS + 1;
  using L = SmemLayout<BLOCK_N, STAGES, CG, NUM_EPI_WARPS, epi_slabs(EPI_MODE)>;
  extern __shared__ uint8_t smem_raw[];
  // the dynamic smem window starts at the same offset in every CTA, so the aligned layout (and therefore
  // every barrier / tile offset) is identical in both CTAs of a pair
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem + L::A_OFF;
  uint8_t* sB = smem + L::B_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], CG);  // leader's arrive.expect_tx (+ the peer's remote arrive)
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], NUM_EPI_WARPS * CG);
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    if (CG == 2) { tmem_alloc_pair(tmem_slot, 2 * BLOCK_N); tmem_relinquish_pair(); }
    else         { tmem_alloc(tmem_slot, 2 * BLOCK_N); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_units = p.num_m_tiles * p.num_n_tiles * p.k_splits;
  const int worker = blockIdx.x / CG;         // CTA (CG=1) or CTA-pair (CG=2) index
  const int num_workers = gridDim.x / CG;
  constexpr uint32_t STAGE_TX = (A_STAGE_BYTES + L::B_STAGE_BYTES) * CG;  // bytes landing per stage, both CTAs

  if (warp == TMA_WARP) {
    {  // all 32 lanes walk the loop (warp-uniform control flow keeps the TMA operands in uniform registers);
       // one elected lane arms the barrier and issues the copies
      int s = 0;
      uint32_t ph = 0;
      for (int unit = worker; unit < total_units; unit += num_workers) {
        // split-major unit order: the workers running at the same time share one K-slice, so the A / B rows of
        // that slice are fetched from HBM once and re-used out of L2 by all tiles
        const int ntiles = p.num_m_tiles * p.num_n_tiles;
        const int split = unit / ntiles;
        const int tile = unit % ntiles;
        const int n_blk = tile % p.num_n_tiles;
        const int m_blk = tile / p.num_n_tiles;
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        const int m0 = (m_blk * CG + (int)cta_rank) * BLOCK_M;
        const int n0 = n_blk * BLOCK_N + (int)cta_rank * L::B_ROWS;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (elect_one()) {
          if (CG == 1 || leader) mbar_arrive_expect_tx(&full_bar[s], STAGE_TX);
          uint8_t* a_dst = sA + s * A_STAGE_BYTES;
          uint8_t* b_dst = sB + s * L::B_STAGE_BYTES;
          if (!A_MN) {
            if (CG == 2) tma_load_2d_pair(a_dst, &tmap_a, &full_bar[s], kb * BLOCK_K, m0);
            else         tma_load_2d(a_dst, &tmap_a, &full_bar[s], kb * BLOCK_K, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i) {
              if (CG == 2) tma_load_2d_pair(a_dst + i * (BLOCK_K * 128), &tmap_a, &full_bar[s], m0 + i * 64, kb * BLOCK_K);
              else         tma_load_2d(a_dst + i * (BLOCK_K * 128), &tmap_a, &full_bar[s], m0 + i * 64, kb * BLOCK_K);
            }
          }
          if (!B_MN) {
            if (CG == 2) tma_load_2d_pair(b_dst, &tmap_b, &full_bar[s], kb * BLOCK_K, n0);
            else         tma_load_2d(b_dst, &tmap_b, &full_bar[s], kb * BLOCK_K, n0);
          } else {
#pragma unroll
            for (int i = 0; i < L::B_ROWS / 64; ++i) {
              if (CG == 2) tma_load_2d_pair(b_dst + i * (BLOCK_K * 128), &tmap_b, &full_bar[s], n0 + i * 64, kb * BLOCK_K);
              else         tma_load_2d(b_dst + i * (BLOCK_K * 128), &tmap_b, &full_bar[s], n0 + i * 64, kb * BLOCK_K);
            }
          }
          if (CG == 2 && !leader) mbar_arrive_remote(&full_bar[s], 0);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    if (leader) {  // whole warp walks the loop; the elected lane issues tcgen05.mma / tcgen05.commit
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M * CG, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // K-major: 8-row groups are 1024 B apart (SBO), LBO unused (1).
      // MN-major: 64-element column slabs are BLOCK_K*128 B apart (LBO), 8-k groups 1024 B (SBO).
      constexpr uint32_t A_LBO = A_MN ? BLOCK_K * 128 : 16;
      constexpr uint32_t B_LBO = B_MN ? BLOCK_K * 128 : 16;
      constexpr uint32_t A_KSTEP = A_MN ? UMMA_K * 128 : UMMA_K * 2;  // bytes per UMMA_K
      constexpr uint32_t B_KSTEP = B_MN ? UMMA_K * 128 : UMMA_K * 2;
      // descriptors differ between stages / k-steps only in the 14-bit start-address field: build once
      const uint64_t a_desc0 = umma_smem_desc_sw128(smem_u32(sA), A_LBO, 1024);
      const uint64_t b_desc0 = umma_smem_desc_sw128(smem_u32(sB), B_LBO, 1024);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (int unit = worker; unit < total_units; unit += num_workers) {
        const int split = unit / (p.num_m_tiles * p.num_n_tiles);
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = a_desc0 + (uint64_t)((s * A_STAGE_BYTES) >> 4);
            const uint64_t db = b_desc0 + (uint64_t)((s * L::B_STAGE_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
              if (CG == 2) umma_f16_ss_pair(d_tmem, da + (uint64_t)((k * A_KSTEP) >> 4), db + (uint64_t)((k * B_KSTEP) >> 4), idesc, accum);
              else         umma_f16_ss(d_tmem, da + (uint64_t)((k * A_KSTEP) >> 4), db + (uint64_t)((k * B_KSTEP) >> 4), idesc, accum);
            }
            // frees the smem stage (in both CTAs) once these MMAs retire
            if (CG == 2) umma_commit_pair(&empty_bar[s]); else umma_commit(&empty_bar[s]);
            if (kb == kb1 - 1) {
              if (CG == 2) umma_commit_pair(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else {
    // ------------------------------- epilogue warps -----------------------------------------
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int part = warp >> 2;   // which 32-column chunks (part, part + W, ...) this warp owns
    const uint32_t stage = smem_u32(smem + L::EPI_OFF + warp * epi_slabs(EPI_MODE) * EPI_STAGE_BYTES);
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int unit = worker; unit < total_units; unit += num_workers) {
      const int tile = unit % (p.num_m_tiles * p.num_n_tiles);
      const int n_blk = tile % p.num_n_tiles;
      const int m_blk = tile / p.num_n_tiles;
      const int row0 = (m_blk * CG + (int)cta_rank) * BLOCK_M + q * 32;
      constexpr int W = NUM_EPI_WARPS / 4;        // warps sharing one TMEM lane quarter split the columns
      constexpr int CHUNKS = BLOCK_N / (32 * W);  // 32-column chunks owned by this warp
      const int colbase = n_blk * BLOCK_N + part * 32;
      const uint32_t tm_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + part * 32;
      if constexpr (EPI_MODE == 3) {
        const uint32_t ring0 = stage + EPI_STAGE_BYTES, ring1 = stage + 2 * EPI_STAGE_BYTES;
        residual_prefetch(p, ring0, lane, row0, colbase);
        if (CHUNKS > 1) residual_prefetch(p, ring1, lane, row0, colbase + 32 * W);
        EpiAux aux[CHUNKS];  // bias / keep bytes of every chunk of the tile: one scoreboard wait per tile
#pragma unroll
        for (int ci = 0; ci < CHUNKS; ++ci) epi_prefetch<EPI_MODE, DROP>(p, aux[ci], lane, row0, colbase + ci * 32 * W);
        mbar_wait(&tmem_full[acc], acc_ph);
        tc_fence_after();
#pragma unroll
        for (int ci = 0; ci < CHUNKS; ++ci) {
          const int col0 = colbase + ci * 32 * W;
          uint32_t v[32];
          tmem_ld_32x32(tm_row + ci * 32 * W, v);
          if (ci + 1 < CHUNKS) cp_async_wait<1>(); else cp_async_wait<0>();  // this chunk's residual has landed
          tmem_ld_wait();
          epilogue_chunk<EPI_MODE, DROP>(p, v, stage, lane, row0, col0, aux[ci], (ci & 1) ? ring1 : ring0,
                                         ci + 2 < CHUNKS ? col0 + 2 * 32 * W : -1);
        }
      } else if constexpr (EPI_MODE == 0) {
        EpiAux aux[2];
        epi_prefetch<EPI_MODE, DROP>(p, aux[0], lane, row0, colbase);  // overlaps the wait for the accumulator
        mbar_wait(&tmem_full[acc], acc_ph);
        tc_fence_after();
#pragma unroll
        for (int ci = 0; ci < CHUNKS; ++ci) {
          const int col0 = colbase + ci * 32 * W;
          if (ci + 1 < CHUNKS) epi_prefetch<EPI_MODE, DROP>(p, aux[(ci + 1) & 1], lane, row0, col0 + 32 * W);
          if (col0 < p.N) {  // warp-uniform
            uint32_t v[32];
            tmem_ld_32x32(tm_row + ci * 32 * W, v);
            tmem_ld_wait();
            epilogue_chunk<EPI_MODE, DROP>(p, v, stage, lane, row0, col0, aux[ci & 1]);
          }
        }
      } else {
        EpiAux aux;
        epi_prefetch<EPI_MODE, DROP>(p, aux, lane, row0, colbase);
        mbar_wait(&tmem_full[acc], acc_ph);
        tc_fence_after();
#pragma unroll
        for (int ci = 0; ci < CHUNKS; ++ci) {
          const int col0 = colbase + ci * 32 * W;
          if (ci > 0) epi_prefetch<EPI_MODE, DROP>(p, aux, lane, row0, col0);
          if (col0 < p.N) {  // warp-uniform
            uint32_t v[32];
            tmem_ld_32x32(tm_row + ci * 32 * W, v);
            tmem_ld_wait();
            epilogue_chunk<EPI_MODE, DROP>(p, v, stage, lane, row0, col0, aux);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2 && !leader) mbar_arrive_remote(&tmem_empty[acc], 0);
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, 2 * BLOCK_N); else tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, int CG, int EPI_MODE, bool DROP = false>
int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid,
                   cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, STAGES, CG, epi_warps(EPI_MODE), epi_slabs(EPI_MODE)>;
  static_assert(L::DYN_BYTES <= 232448, "shared memory budget exceeded");
  auto kern = gemm_tc_kernel<BLOCK_N, STAGES, A_MN, B_MN, CG, EPI_MODE, DROP>;
  static uint64_t attr_done = 0;  // per instantiation, one bit per device
  if (nv_first_on_device(&attr_done))
    NV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(num_threads(EPI_MODE));
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NV_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  return NV_OK;
}

// pipeline depth: what fits in 227 KB next to the epilogue staging (4 KB per epilogue warp)
template <int BLOCK_N, int CG, int EPI_MODE>
constexpr int stages_for() {
  constexpr int stage_bytes = A_STAGE_BYTES + (BLOCK_N / CG) * BLOCK_K * 2;
  constexpr int budget = 232448 - 1024 - 256 - epi_warps(EPI_MODE) * epi_slabs(EPI_MODE) * EPI_STAGE_BYTES;
  constexpr int n = budget / stage_bytes;
  return n > 8 ? 8 : n;
}

template <int BLOCK_N, int CG>
int launch_major(int a_mn, int b_mn, int epi_mode, const CUtensorMap& ta, const CUtensorMap& tb,
                 const GemmParams& p, int grid, cudaStream_t stream) {
  // dropout is a compile-time variant (the Philox code costs registers and scheduling freedom in the epilogue)
  // built only for the three layouts that have a dropout site: forward linear, GELU forward, dgrad through GELU
  const bool drop = p.drop_thr != 0;
  if (epi_mode == 1) {  // GELU forward: activations x weights, both K-major
    NV_REQUIRE(!a_mn && !b_mn, "gemm: apply_gelu is only built for K-major operands (forward linear)");
    if (drop) return launch_variant<BLOCK_N, stages_for<BLOCK_N, CG, 1>(), false, false, CG, 1, true>(ta, tb, p, grid, stream);
    return launch_variant<BLOCK_N, stages_for<BLOCK_N, CG, 1>(), false, false, CG, 1>(ta, tb, p, grid, stream);
  }
  if (epi_mode == 2) {  // dgrad through GELU: dY [M,K] x W stored [K,N]
    NV_REQUIRE(!a_mn && b_mn, "gemm: gelu_u is only built for the dgrad layout (A K-major, B MN-major)");
    if (drop) return launch_variant<BLOCK_N, stages_for<BLOCK_N, CG, 2>(), false, true, CG, 2, true>(ta, tb, p, grid, stream);
    return launch_variant<BLOCK_N, stages_for<BLOCK_N, CG, 2>(), false, true, CG, 2>(ta, tb, p, grid, stream);
  }
  constexpr int S = stages_for<BLOCK_N, CG, 0>();
  if constexpr (CG == 2) {
    // forward linear that adds the fp32 residual stream: residual staged by cp.async (mode 3). The ring costs two
    // pipeline stages, which only a short K loop can spare (K = 2048: 84 -> 98 us with four stages).
    static const bool res_async = getenv("NV_GEMM_NO_RES_ASYNC") == nullptr;
    if (res_async && !a_mn && !b_mn && p.residual != nullptr && !(p.flags & EPI_ATOMIC) && p.K <= 1024 &&
        p.ld_res % 4 == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0) {
      constexpr int S3 = stages_for<BLOCK_N, CG, 3>();
      if (drop) return launch_variant<BLOCK_N, S3, false, false, CG, 3, true>(ta, tb, p, grid, stream);
      return launch_variant<BLOCK_N, S3, false, false, CG, 3>(ta, tb, p, grid, stream);
    }
  }
  if (drop) {
    NV_REQUIRE(!a_mn && !b_mn, "gemm: dropout in the store epilogue is only built for K-major operands (forward linear)");
    return launch_variant<BLOCK_N, S, false, false, CG, 0, true>(ta, tb, p, grid, stream);
  }
  if (!a_mn && !b_mn) return launch_variant<BLOCK_N, S, false, false, CG, 0>(ta, tb, p, grid, stream);
  if (!a_mn && b_mn) return launch_variant<BLOCK_N, S, false, true, CG, 0>(ta, tb, p, grid, stream);
  if (a_mn && !b_mn) return launch_variant<BLOCK_N, S, true, false, CG, 0>(ta, tb, p, grid, stream);
  return launch_variant<BLOCK_N, S, true, true, CG, 0>(ta, tb, p, grid, stream);
}

}  // namespace

// Host entry used by the C ABI (api.cu). All pointers are device pointers; lds in elements.
int nv_gemm_tc_launch(int a_mn, int b_mn, int M, int N, int K, const bf16* A, int64_t lda,
                      const bf16* B, int64_t ldb, const float* bias, const float* residual,
                      int64_t ld_res, const bf16* gelu_u, int64_t ld_u, float* out_f32,
                      int64_t ld_f32, bf16* out_bf16, int64_t ld_bf16, bf16* out_pre, int64_t ld_pre,
                      float* colsum, int apply_gelu, int accumulate, float alpha, int k_splits, int block_n,
                      int cta_group, float dropout_p, uint64_t dropout_seed, int dropout_stream,
                      const uint8_t* dropout_bits, int dropout_row_mul, cudaStream_t stream) {
  NV_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  NV_REQUIRE(N % 8 == 0, "gemm: N=%d must be a multiple of 8", N);
  NV_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements (TMA 16B strides)");
  NV_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
             "gemm: operands must be 16-byte aligned");
  NV_REQUIRE(out_f32 != nullptr || out_bf16 != nullptr, "gemm: no output buffer");
  NV_REQUIRE(!accumulate || out_f32 != nullptr, "gemm: accumulate needs an fp32 output");
  NV_REQUIRE(block_n == 128 || block_n == 256 || block_n == 0, "gemm: block_n must be 0, 128 or 256");
  NV_REQUIRE(cta_group >= 0 && cta_group <= 2, "gemm: cta_group must be 0 (auto), 1 or 2");

  // one row tile (the 64-row GEMMs of the cls-only last layer, small batches): 128-wide tiles put twice the CTAs on the
  // operand stream of the K loop — 64 x 1024 x 512: 12.3 -> 10.2 us, 64 x 1024 x 2048: 20.5 -> 16.4 us (tools/small_gemm_probe.py)
  if (block_n == 0) block_n = (N >= 256 && M > BLOCK_M) ? 256 : 128;
  const int num_sms = nv_num_sms();
  if (cta_group == 0) cta_group = (M > BLOCK_M && num_sms % 2 == 0) ? 2 : 1;
  const int b_rows = block_n / cta_group;  // B rows staged per CTA

  CUtensorMap ta, tb;
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!a_mn) { dims[0] = (uint64_t)K; dims[1] = (uint64_t)M; box[0] = BLOCK_K; box[1] = BLOCK_M; }
    else       { dims[0] = (uint64_t)M; dims[1] = (uint64_t)K; box[0] = 64;      box[1] = BLOCK_K; }
    strides[0] = (uint64_t)lda * 2;
    int s = nv_encode_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, strides, box,
                           CU_TENSOR_MAP_SWIZZLE_128B);
    if (s != NV_OK) return s;
    if (!b_mn) { dims[0] = (uint64_t)K; dims[1] = (uint64_t)N; box[0] = BLOCK_K; box[1] = (uint32_t)b_rows; }
    else       { dims[0] = (uint64_t)N; dims[1] = (uint64_t)K; box[0] = 64;      box[1] = BLOCK_K; }
    strides[0] = (uint64_t)ldb * 2;
    s = nv_encode_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, dims, strides, box,
                       CU_TENSOR_MAP_SWIZZLE_128B);
    if (s != NV_OK) return s;
  }

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_tiles = (M + BLOCK_M * cta_group - 1) / (BLOCK_M * cta_group);
  p.num_n_tiles = (N + block_n - 1) / block_n;
  p.k_blocks_total = (K + BLOCK_K - 1) / BLOCK_K;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > p.k_blocks_total) k_splits = p.k_blocks_total;
  p.k_blocks_per_split = (p.k_blocks_total + k_splits - 1) / k_splits;
  p.k_splits = (p.k_blocks_total + p.k_blocks_per_split - 1) / p.k_blocks_per_split;
  NV_REQUIRE(p.k_splits == 1 || accumulate, "gemm: split-K needs accumulate=1 (red.add epilogue)");
  p.bias = bias; p.residual = residual; p.gelu_u = gelu_u;
  p.out_f32 = out_f32; p.out_bf16 = out_bf16; p.out_pre = out_pre; p.colsum = colsum;
  NV_REQUIRE(colsum == nullptr || p.k_splits == 1, "gemm: colsum needs k_splits == 1");
  p.ld_res = ld_res; p.ld_u = ld_u; p.ld_f32 = ld_f32; p.ld_bf16 = ld_bf16; p.ld_pre = ld_pre;
  p.flags = (apply_gelu ? EPI_GELU : 0) | (accumulate ? EPI_ATOMIC : 0);
  p.alpha = alpha;
  NV_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "gemm: dropout_p %f out of range [0, 1)", dropout_p);
  NV_REQUIRE(dropout_p == 0.f || !accumulate, "gemm: dropout and accumulate (split-K) are mutually exclusive");
  p.drop_thr = nv_dropout_threshold(dropout_p);
  p.keep_scale = nv_dropout_keep_scale(p.drop_thr);
  p.drop_seed = dropout_seed;
  p.drop_epoch = p.drop_thr != 0 ? nv_rng_epoch_dev() : nullptr;
  p.drop_bits = p.drop_thr != 0 ? dropout_bits : nullptr;
  p.drop_row_mul = dropout_row_mul > 0 ? dropout_row_mul : 1;
  p.drop_stream = (uint32_t)dropout_stream;

  const int total_units = p.num_m_tiles * p.num_n_tiles * p.k_splits;
  const int max_workers = num_sms / cta_group;
  const int grid = (total_units < max_workers ? total_units : max_workers) * cta_group;
  NV_REQUIRE(!(apply_gelu && gelu_u != nullptr), "gemm: apply_gelu and gelu_u are mutually exclusive");
  NV_REQUIRE(!((apply_gelu || gelu_u != nullptr) && (residual != nullptr || accumulate || alpha != 1.0f)),
             "gemm: the GELU epilogues take no residual / accumulate / alpha");
  NV_REQUIRE(!(gelu_u != nullptr && bias != nullptr), "gemm: the GELU-grad epilogue takes no bias");
  const int epi_mode = apply_gelu ? 1 : (gelu_u != nullptr ? 2 : 0);
  if (cta_group == 2) {
    if (block_n == 256) return launch_major<256, 2>(a_mn, b_mn, epi_mode, ta, tb, p, grid, stream);
    return launch_major<128, 2>(a_mn, b_mn, epi_mode, ta, tb, p, grid, stream);
  }
  if (block_n == 256) return launch_major<256, 1>(a_mn, b_mn, epi_mode, ta, tb, p, grid, stream);
  return launch_major<128, 1>(a_mn, b_mn, epi_mode, ta, tb, p, grid, stream);
}
