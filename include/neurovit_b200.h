/*
 * neurovit_b200 — C ABI of the B200 (sm_100a) ViT3D / NeuroEncoder hot path.
 *
 * The reference (gillet-thomas/NeuroViT) is pure PyTorch and has no FFI of its own; every entry point
 * below therefore cites the reference *Python* call site it replaces (file:line under the reference
 * root). Host code (neurovit_b200/*.py, the drop-in nn.Modules) binds these symbols with ctypes; see
 * INTEGRATION.md for the binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates; the library never
 *     frees or retains memory past the call);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no hidden syncs;
 *   - leading dimensions / strides are in ELEMENTS; sizes are plain ints;
 *   - return value: 0 = NV_OK, 1 = bad argument, 2 = unsupported shape, 3 = CUDA error,
 *     4 = not initialised; nv_last_error() returns the thread-local message of the last failure;
 *   - there is NO CPU fallback: on a device that is not sm_100 nv_device_check() fails.
 */
#ifndef NEUROVIT_B200_H
#define NEUROVIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NV_ABI_VERSION 1

int nv_version(void);
const char* nv_last_error(void);
/* 0 iff the current CUDA device is compute capability 10.x (B200) */
int nv_device_check(void);

/* ---- linear layers: nn.Linear fwd / dgrad / wgrad ------------------------------------------------
 * replaces: vit_3d.py:94 (patch embedding), :41,50 (to_qkv), :43-45,60 (to_out), :19-22 (FeedForward)
 * and their autograd backward.  C = epilogue(alpha * A * B^T), bf16 operands, fp32 accumulation on
 * tcgen05 tensor cores (TMEM accumulators, TMA-fed 128B-swizzled tiles).
 *   a_mn = 0: A is [M,K] row-major (lda);  a_mn = 1: A is stored [K,M] row-major (lda)
 *   b_mn = 0: B is [N,K] row-major (ldb);  b_mn = 1: B is stored [K,N] row-major (ldb)
 * epilogue, in order: *alpha, +bias[N], *gelu'(gelu_u[M,N]) (dgrad through GELU), out_pre = value and
 * value = gelu(value) when apply_gelu, +residual[M,N] (fp32), then store: out_f32 (or red.add into it
 * when accumulate=1 — required for k_splits > 1) and/or the bf16 copy out_bf16; colsum[N] (optional)
 * += sum over rows of the stored value (the bias gradient of the layer below, fused into the dgrad).
 * dropout_p > 0: nn.Dropout on the value (vit_3d.py:21,23,45), applied after bias / GELU and BEFORE the
 * residual add; with gelu_u it multiplies the gradient by the forward mask of the activation. The keep
 * mask is a pure function of (dropout_seed, dropout_stream, row * N + col) — nv_dropout draws the same one.
 * dropout_row_mul (>= 1): the mask row is output row * dropout_row_mul — a compact problem over every n-th row of a
 * site (the cls rows in the last block's backward) then sees that site's forward mask.
 * block_n: 0 = auto, or 128 / 256 (CTA tile 128 x block_n). */
int nv_gemm_bf16(int a_mn, int b_mn, int M, int N, int K,
                 const void* A, int64_t lda, const void* B, int64_t ldb,
                 const float* bias, const float* residual, int64_t ld_res,
                 const void* gelu_u, int64_t ld_u,
                 float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16,
                 void* out_pre, int64_t ld_pre, float* colsum,
                 int apply_gelu, int accumulate, float alpha, int k_splits, int block_n, int cta_group,
                 float dropout_p, int64_t dropout_seed, int dropout_stream, const void* dropout_bits,
                 int dropout_row_mul, void* stream);

/* Keep bits drawn ahead of time: out[g] (one byte) = the 8 keep bits of elements [8g, 8g+8) of a dropout site
 * with the given (seed, stream_id) — exactly the bits the consumers would draw inline. nv_gemm_bf16
 * (dropout_bits), nv_layernorm_bwd (side_drop_bits) and nv_attention_fwd (drop_mask_ready) take them, which moves
 * the Philox arithmetic out of their epilogues / softmax rows onto a side stream. n_groups % 4 == 0. */
int nv_dropout_bits(void* out, int64_t n_groups, float p, int64_t seed, int stream_id, void* stream);

/* ---- element-wise dropout ---------------------------------------------------------------------------
 * replaces: nn.Dropout on the embedding (vit_3d.py:100,119) and, in backward, the mask applied to the
 * gradient of a dropped-out linear output.  v = in * keep / (1 - p_eff); out_f32 / out_bf16 = v (+ residual);
 * colsum[N] += column sums of v. p = 0 turns it into a copy / cast / column sum. N % 8 == 0. */
int nv_dropout(const float* in, int64_t ld_in, const float* residual, int64_t ld_res,
               float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16, float* colsum,
               int M, int N, float p, int64_t seed, int stream_id, int row_mul, void* stream);

/* fp32 verification GEMM (CUDA-core FMA, arbitrary strides, batch index z = z1*Z2 + z2):
 * C[z][m][n] = epilogue(alpha * sum_k A[z][m,k] * B[z][n,k]); same epilogue order as nv_gemm_bf16, all
 * fp32; residual/gelu_u/out_pre share C's batch offsets. Also serves the [B,1024]x[1024,num_classes]
 * head (vit_3d.py:107-110). */
int nv_gemm_f32(int M, int N, int K, int Z1, int Z2,
                const float* A, int64_t sa_m, int64_t sa_k, int64_t sa_z1, int64_t sa_z2,
                const float* B, int64_t sb_n, int64_t sb_k, int64_t sb_z1, int64_t sb_z2,
                float* C, int64_t sc_m, int64_t sc_z1, int64_t sc_z2,
                const float* bias, const float* residual, int64_t ld_res,
                const float* gelu_u, int64_t ld_u, float* out_pre, int64_t ld_pre,
                int apply_gelu, int accumulate, float alpha, void* stream);

/* ---- LayerNorm ------------------------------------------------------------------------------------
 * replaces: nn.LayerNorm at vit_3d.py:18,37,93,95,108 (eps 1e-5 default, biased variance).
 * Rows may be addressed through a grouped map r -> (r / group) * gstride + goff + r % group
 * (group = 0: identity) so the [B, n+1, D] token tensor can be walked without its cls rows.
 * add (optional): y += add[(r % add_mod) + add_off, :]  — the positional embedding (vit_3d.py:118). */
int nv_layernorm_fwd(const float* x, int64_t ld_x, int x_group, int x_gstride, int x_goff,
                     const float* gamma, const float* beta,
                     const float* add, int64_t ld_add, int add_mod, int add_off,
                     void* y, int y_is_bf16, int64_t ld_y, int y_group, int y_gstride, int y_goff,
                     float* mean, float* rstd, int M, int D, float eps, void* stream);
/* dx = LNbwd(dy) (+ dres); dgamma/dbeta/colsum are ACCUMULATED (atomicAdd) — zero or pre-load them.
 * colsum (optional) += sum_rows dx_out: the bias gradient of the linear feeding the residual stream.
 * dy is fp32, or bf16 when dy_is_bf16 (the dgrad GEMM's output in bf16 mode).
 * side_drop_p > 0: the bf16 copy dx_bf16 and colsum (NOT the fp32 dx) are multiplied by the dropout mask
 * (side_drop_seed, side_drop_stream, row * D + col) — the forward mask of the linear layer whose backward
 * consumes them (vit_3d.py:23,45), saving a separate masking pass. Needs M >= 64 and D % 8 == 0. */
int nv_layernorm_bwd(const void* dy, int dy_is_bf16, int64_t ld_dy, int dy_group, int dy_gstride, int dy_goff,
                     const float* x, int64_t ld_x, int x_group, int x_gstride, int x_goff,
                     const float* mean, const float* rstd, const float* gamma,
                     const float* dres, int64_t ld_dres,
                     float* dx, int64_t ld_dx, int dx_group, int dx_gstride, int dx_goff,
                     void* dx_bf16, int64_t ld_dxb,
                     float* dgamma, float* dbeta, float* colsum, int M, int D,
                     float side_drop_p, int64_t side_drop_seed, int side_drop_stream, const void* side_drop_bits,
                     void* stream);
/* x[b, 0, :] = cls + pos[0]  (vit_3d.py:116-118) */
int nv_cls_row(const float* cls, const float* pos, float* x, int64_t batch_stride, int B, int D, void* stream);

/* ---- 3D patch gather + LayerNorm(patch_dim) -------------------------------------------------------
 * replaces: Rearrange('b c (f pf) (h p1) (w p2) -> b (f h w) (p1 p2 pf c)') + nn.LayerNorm(patch_dim)
 * at vit_3d.py:92-93, on the strided view built by ViT3DEncoder.forward (NeuroEncoder.py:200-202).
 * dims[5] = {B,C,F,H,W}, strides[5] = element strides of the view, patch[3] = {pf,p1,p2}.
 * out [B*n, ld_out] (bf16 or fp32; columns >= patch_dim are zero), raw (optional, fp32 [B*n,patch_dim])
 * receives the un-normalised gathered patches — the bit-exact patch-index contract. */
int nv_patch_gather_ln(const float* video, const int64_t* dims, const int64_t* strides, const int64_t* patch,
                       const float* gamma, const float* beta, void* out, int out_is_bf16, int64_t ld_out,
                       float* raw, float* mean, float* rstd, float eps, void* stream);
/* dgamma/dbeta of that LayerNorm from dP = d(loss)/d(LN output) [B*n, ld_dp] (accumulated) */
int nv_patch_ln_param_grad(const float* video, const int64_t* dims, const int64_t* strides, const int64_t* patch,
                           const float* dP, int64_t ld_dp, const float* mean, const float* rstd,
                           float* dgamma, float* dbeta, void* stream);

/* LayerNorm(patch_dim) folded into the Linear(patch_dim, dim) behind it (vit_3d.py:93-94):
 * Linear(LN(x)) = xhat (W o gamma)^T + (W beta + b). nv_ln_fold writes Wf [D, ld_wf] (bf16 or fp32; columns P..ld_wf-1
 * zero: K padded to a multiple of 8) and bias_f [D] (b may be NULL); nv_patch_gather_ln with gamma = 1, beta = 0
 * delivers xhat. nv_ln_fold_grads turns G = de^T xhat [D, ld_g >= P] (one weight-gradient GEMM) and cs = column sums
 * of de [D] into the gradients of all four parameters, accumulating: dW[D,P] += G o gamma + cs beta^T,
 * dgamma[P] += sum_k W o G, dbeta[P] += W^T cs, db[D] += cs (db may be NULL). W [D, P] fp32, contiguous. */
int nv_ln_fold(const float* W, const float* gamma, const float* beta, const float* b, void* Wf, int wf_is_bf16,
               int64_t ld_wf, float* bias_f, int D, int P, void* stream);
int nv_ln_fold_grads(const float* G, int64_t ld_g, const float* W, const float* gamma, const float* beta, const float* cs,
                     float* dW, float* dgamma, float* dbeta, float* db, int D, int P, void* stream);

/* ---- attention --------------------------------------------------------------------------------------
 * replaces: vit_3d.py:51-59. q/k/v are read in place from the QKV projection output
 * [B, N, 3*H*64] (pointers to the q, k, v column blocks; shared batch/row strides), O is written as
 * [B, N, H*64]; lse [B,H,N] fp32 is saved for backward. head_dim must be 64 (bf16 flash kernels,
 * tcgen05 / TMEM; token 0 of every sample — the cls token — is handled outside the 128-row tiles, which cover
 * tokens 1..N-1). dropout_p > 0 applies nn.Dropout to the probabilities (vit_3d.py:56): forward draws the
 * keep bits from (seed) and saves them in drop_mask, uint32 [B*H, N, ceil(N/32)]: bit k of word w of row q =
 * the score of query token q and the key at position 32w+k survives, where key token t sits at position t-1
 * for t >= 1 and key token 0 at position N-1; backward reads them. drop_mask may be NULL when dropout_p == 0.
 * drop_mask_ready = 1: drop_mask already holds the bits (nv_dropout_bits with the same seed, stream 0, over
 * B*H*N*4*ceil(N/32) groups) and forward reads them instead of drawing them inside its softmax rows. */
int nv_attention_fwd(const void* q, const void* k, const void* v, int64_t qkv_batch_stride, int64_t qkv_row_stride,
                     void* o, int64_t o_batch_stride, int64_t o_row_stride, float* lse,
                     int B, int N, int H, int head_dim, float scale, float dropout_p, int64_t seed,
                     void* drop_mask, int drop_mask_ready, void* stream);
/* Forward for the cls query only: with pool='cls' (vit_3d.py:123) the LAST block's attention output is used at token 0
 * alone, so that block computes one query row per (batch, head) — the same SIMT code, dropout bits and arithmetic that
 * serve token 0 inside nv_attention_fwd. o_cls row b (H*64 bf16) is written at o_cls + b*o_batch_stride; lse and
 * drop_mask keep nv_attention_fwd's layouts ([B,H,N] and [B*H, N, ceil(N/32)]) with only the token-0 entries / rows
 * written (drop_mask_ready = 1: read instead). Pairs with nv_attention_cls_bwd. */
int nv_attention_cls_fwd(const void* q, const void* k, const void* v, int64_t qkv_batch_stride, int64_t qkv_row_stride,
                         void* o_cls, int64_t o_batch_stride, float* lse, int B, int N, int H, int head_dim, float scale,
                         float dropout_p, int64_t seed, void* drop_mask, int drop_mask_ready, void* stream);
/* Backward when only the cls query (token 0 of every sample) carries gradient (last block, pool='cls'):
 * dO_cls [B, H*64] holds that row's dO (batch stride dO_batch_stride), o's token-0 rows are read at o + b*o_batch_stride.
 * Writes dq (zero except token 0), dk, dv for every token: O(N d) per (batch, head). */
int nv_attention_cls_bwd(const void* q, const void* k, const void* v, int64_t qkv_batch_stride, int64_t qkv_row_stride,
                         const void* o, int64_t o_batch_stride, const void* dO_cls, int64_t dO_batch_stride,
                         const float* lse, void* dq, void* dk, void* dv, int64_t dqkv_batch_stride,
                         int64_t dqkv_row_stride, int B, int N, int H, int head_dim, float scale, float dropout_p,
                         const void* drop_mask, void* stream);
/* delta_ws: fp32 workspace of B*H*N elements (delta_i = dO_i . O_i, written by the dQ kernel's prologue, read by
 * the dK/dV kernel: no separate pass) */
int nv_attention_bwd(const void* q, const void* k, const void* v, int64_t qkv_batch_stride, int64_t qkv_row_stride,
                     const void* o, const void* dO, int64_t o_batch_stride, int64_t o_row_stride,
                     const float* lse, float* delta_ws,
                     void* dq, void* dk, void* dv, int64_t dqkv_batch_stride, int64_t dqkv_row_stride,
                     int B, int N, int H, int head_dim, float scale, float dropout_p, const void* drop_mask,
                     void* stream);
/* fp32 verification path: materialised softmax (vit_3d.py:55) and its backward, in place */
int nv_softmax_fwd(float* s, int64_t rows, int n, void* stream);
int nv_softmax_bwd(const float* P, float* dP, int64_t rows, int n, void* stream);

/* ---- helpers around the GEMMs ---------------------------------------------------------------------- */
int nv_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream);
/* out[r,c] (optional) and outT[c,r] = bf16(in[r,c]): bf16 weight caches for fwd and dgrad */
int nv_cast_transpose_f32_bf16(const float* in, void* out, void* outT, int R, int C, void* stream);
/* out[c] += sum_r in[r,c]  (bias gradients) */
int nv_colsum(const void* in, int in_is_bf16, int64_t ld, float* out, int M, int N, void* stream);
/* out[j] += sum_b in[b*batch_stride + j], j < L  (pos_embedding / cls_token gradients) */
int nv_batch_sum(const float* in, int64_t batch_stride, float* out, int B, int64_t L, void* stream);
/* pool='mean' (vit_3d.py:123) */
int nv_mean_pool_fwd(const float* x, float* pooled, int B, int N, int D, void* stream);
int nv_mean_pool_bwd(const float* dpooled, float* dx, void* dx_bf16, int B, int N, int D, void* stream);

/* ---- classification head on the cls token --------------------------------------------------------------
 * replaces: x[:, 0] -> nn.LayerNorm(dim) -> nn.Linear(dim, num_classes) at vit_3d.py:105-110,123-126 (pool='cls').
 * x: [B, ...] fp32 with the cls row of sample b at x + b*ld_x; y [B,D] (LayerNorm output, saved), mean/rstd [B],
 * logits [B,C]. bwd: dx rows (stride ld_dx, optional bf16 copy) get the cls-token gradient; dW [C,D], db [C],
 * dgamma/dbeta [D] are ACCUMULATED (zero or pre-load them). fp32 in both precision modes. */
int nv_head_fwd(const float* x, int64_t ld_x, const float* gamma, const float* beta, const float* W, const float* bias,
                float* y, float* mean, float* rstd, float* logits, int B, int D, int C, float eps, void* stream);
int nv_head_bwd(const float* dl, const float* x, int64_t ld_x, const float* y, const float* mean, const float* rstd,
                const float* gamma, const float* W, float* dx, int64_t ld_dx, void* dx_bf16, int64_t ld_dxb,
                float* dW, float* db, float* dgamma, float* dbeta, int B, int D, int C, void* stream);

/* ---- fused AdamW over flat buffers -------------------------------------------------------------------
 * replaces: optim.AdamW(model.parameters(), lr, weight_decay) + optimizer.step() at src/Trainer.py:31,75
 * (torch semantics: p *= 1 - lr*wd; m, v moments; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)) when parameters,
 * gradients and moments live in flat fp32 buffers of n elements (n % 4 == 0). p_bf16 (optional) receives the
 * bf16 copy of the updated parameters — the weight cache the next forward's GEMMs read. step counts from 1;
 * step_dev (optional): the step count as a device float (read by the kernel instead of `step`, so a captured
 * CUDA graph keeps counting — bump it with nv_counter_add before the launch). */
int nv_adamw_flat(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                  const float* step_dev, void* stream);
/* *counter += inc on the stream (device-side step counters of captured graphs) */
int nv_counter_add(float* counter, float inc, void* stream);
/* Every dropout mask is Philox(seed + epoch * c, ...) where epoch is a device-resident counter (0 until advanced):
 * a training step captured in a CUDA graph ends with nv_rng_epoch_advance so each replay draws new masks even
 * though the host-drawn seeds are baked into the graph. Forward and backward of a step share the epoch. */
int nv_rng_epoch_advance(void* stream);
/* Current value of that counter, copied to *out_host (synchronises the stream). With it the keep bits of any site are
 * a closed-form function of (seed, epoch, stream, element index): oracle/rng_oracle.py restates it in numpy and
 * tests/test_gpu_kernels.py pins nv_dropout_bits / nv_dropout to it bit for bit. */
int nv_rng_epoch_get(unsigned long long* out_host, void* stream);

/* ---- 4D temporal head -------------------------------------------------------------------------------
 * replaces: TemporalTransformer (nn.TransformerEncoderLayer(d_model=2, nhead=2, batch_first=True),
 * post-norm, ReLU, dim_ff = F) + mean over T + ProjectionHead Linear(2,2); NeuroEncoder.py:63-66,207-230.
 * params: packed fp32 vector, layout documented in neurovit_b200/functional.py (TEMPORAL_LAYOUT).
 * x [B,T,2] -> out [B,2] (optional) and/or seq_out [B,T,2] = the encoder layer's per-timepoint output
 * (TemporalTransformer.forward called on its own); saved [B, T*4] keeps the LN1 output and LN2 input.
 * bwd takes dout [B,2] and/or dseq [B,T,2] (either may be null), writes per-sequence parameter
 * gradients to dparams_ws [B, P] (reduce with nv_batch_sum) and dx [B,T,2] (optional).
 * p_attn / p_drop1 / p_ffn / p_drop2: training-mode dropout of the layer's four nn.Dropout sites (attention
 * weights, after the self-attention block, after ReLU, after linear2; torch default 0.1, 0 = eval); masks are
 * Philox bits of (seed, site, element index) — pass the same seed to bwd. */
int nv_temporal_fwd(const float* x, const float* params, float* out, float* seq_out, float* saved,
                    int B, int T, int F, float eps,
                    float p_attn, float p_drop1, float p_ffn, float p_drop2, int64_t seed, void* stream);
int nv_temporal_bwd(const float* x, const float* params, const float* saved, const float* dout, const float* dseq,
                    float* dparams_ws, float* dx, int B, int T, int F, float eps,
                    float p_attn, float p_drop1, float p_ffn, float p_drop2, int64_t seed, void* stream);

/* ---- 4D input pipeline ---------------------------------------------------------------------------------
 * replaces: NeuroEncoder.py:54-56 (fmri.permute(0, 4, 1, 2, 3) + reshape(B*T, H, W, D): the time axis leaves the
 * innermost position) and, optionally fused into the same pass, the dataset's per-sample z-score of
 * src/data/DatasetADNI_4D.py:84-86 ((x - mean) / (std + eps) over all H*W*D*T values, population std).
 * x [B, S, T] fp32 contiguous (S = H*W*D) -> y [B, T, S]. stats_ws NULL: plain de-interleave, bit-exact with the
 * reference's strided copy. stats_ws = DEVICE workspace of 2*B doubles: z-score, moments accumulated in fp64. */
int nv_fmri_deinterleave(const float* x, float* y, int B, int64_t S, int T, void* stats_ws, double eps, void* stream);

/* ---- data-parallel exchange step ---------------------------------------------------------------------
 * New work (the reference is single-GPU): the gradient all-reduce at the step boundary of
 * src/Trainer.py:65-76, over NCCL on NVLink 5 / NVSwitch. The library owns one communicator per device; NCCL is
 * resolved with dlopen (the copy PyTorch ships; no link-time dependency). The all-reduce is enqueued on the caller's
 * stream like any kernel, so it can be captured into the training step's CUDA graph and overlapped with backward
 * (neurovit_b200/trainer.py launches one per gradient bucket on a side stream, followed by that bucket's AdamW).
 *   nv_dp_load(path)            dlopen NCCL (path NULL/"" = "libnccl.so.2"); 4 (not initialised) if unavailable
 *   nv_dp_nccl_version()        NCCL version code (22809 = 2.28.9), 0 when not loaded   [returns the value]
 *   nv_dp_unique_id(out128)     rank 0: 128-byte id (HOST pointer) to hand to every rank
 *   nv_dp_init(uid128, r, w, c) collective: create the communicator of rank r of w on the current device; c > 0 caps
 *                               the CTAs of its collectives (ncclConfig_t.maxCTAs: per communicator, unlike NCCL_MAX_CTAS,
 *                               which NCCL reads once per process)
 *   nv_dp_register(buf, bytes)  register a long-lived DEVICE buffer (the flat gradient buffer): zero-copy / NVLS
 *   nv_dp_allreduce_bucket      in place over `count` elements; dtype 0 = fp32, 1 = bf16; op 0 = sum, 1 = average
 *   nv_dp_world(rank*, world*)  HOST int pointers
 *   nv_dp_destroy()             */
int nv_dp_load(const char* path);
/* SMs (rounded up to even) the library's persistent kernels (GEMM, LayerNorm backward) leave free on the current
 * device for a concurrent communication kernel; 0 = use every SM (default). */
int nv_set_sm_reserve(int n);
int nv_dp_nccl_version(void);
int nv_dp_unique_id(void* out128);
int nv_dp_init(const void* uid128, int rank, int world, int max_ctas);
int nv_dp_register(void* buf, int64_t bytes);
int nv_dp_allreduce_bucket(void* buf, int64_t count, int dtype, int op, void* stream);
int nv_dp_world(int* rank, int* world);
int nv_dp_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* NEUROVIT_B200_H */
