// extern "C" surface of libneurovit_b200.so — thin argument checks + launches. See include/neurovit_b200.h.
#include "nv_common.cuh"
#include "../../include/neurovit_b200.h"

const char* nv_last_error_impl();
int nv_set_sm_reserve_impl(int n);

// launchers implemented in the kernel translation units
int nv_gemm_tc_launch(int a_mn, int b_mn, int M, int N, int K, const bf16* A, int64_t lda, const bf16* B, int64_t ldb,
                      const float* bias, const float* residual, int64_t ld_res, const bf16* gelu_u, int64_t ld_u,
                      float* out_f32, int64_t ld_f32, bf16* out_bf16, int64_t ld_bf16, bf16* out_pre, int64_t ld_pre,
                      float* colsum, int apply_gelu, int accumulate, float alpha, int k_splits, int block_n, int cta_group,
                      float dropout_p, uint64_t dropout_seed, int dropout_stream, const uint8_t* dropout_bits,
                      int dropout_row_mul, cudaStream_t stream);
int nv_head_fwd_launch(const float* x, int64_t ld_x, const float* gamma, const float* beta, const float* W,
                       const float* bias, float* y, float* mean, float* rstd, float* logits, int B, int D, int C,
                       float eps, cudaStream_t stream);
int nv_head_bwd_launch(const float* dl, const float* x, int64_t ld_x, const float* y, const float* mean,
                       const float* rstd, const float* gamma, const float* W, float* dx, int64_t ld_dx, bf16* dx_bf16,
                       int64_t ld_dxb, float* dW, float* db, float* dgamma, float* dbeta, int B, int D, int C,
                       cudaStream_t stream);
int nv_dropout_bits_launch(uint32_t* out, int64_t n_groups, float p, uint64_t seed, int stream_id, cudaStream_t stream);
int nv_adamw_flat_launch(float* p, const float* g, float* m, float* v, bf16* p_bf16, int64_t n, float lr, float beta1,
                         float beta2, float eps, float weight_decay, int step, const float* step_dev,
                         cudaStream_t stream);
int nv_rng_epoch_advance_launch(cudaStream_t stream);
int nv_rng_epoch_read(uint64_t* out, cudaStream_t stream);
int nv_counter_add_launch(float* counter, float inc, cudaStream_t stream);
int nv_dropout_launch(const float* in, int64_t ld_in, const float* residual, int64_t ld_res, float* out_f32,
                      int64_t ld_f32, bf16* out_bf16, int64_t ld_bf16, float* colsum, int M, int N, float p,
                      uint64_t seed, int stream_id, int row_mul, cudaStream_t stream);
int nv_simt_gemm_launch(int M, int N, int K, int Z1, int Z2, const float* A, int64_t sa_m, int64_t sa_k, int64_t sa_z1,
                        int64_t sa_z2, const float* B, int64_t sb_n, int64_t sb_k, int64_t sb_z1, int64_t sb_z2,
                        float* C, int64_t sc_m, int64_t sc_z1, int64_t sc_z2, const float* bias, const float* residual,
                        int64_t ld_res, const float* gelu_u, int64_t ld_u, float* out_pre, int64_t ld_pre,
                        int apply_gelu, int accumulate, float alpha, cudaStream_t stream);
int nv_softmax_fwd_launch(float* s, int64_t rows, int n, cudaStream_t stream);
int nv_softmax_bwd_launch(const float* P, float* dP, int64_t rows, int n, cudaStream_t stream);
int nv_ln_fwd_launch(const float* x, int64_t ld_x, int xg, int xs, int xo, const float* gamma, const float* beta,
                     const float* add, int64_t ld_add, int add_mod, int add_off, void* y, int y_is_bf16, int64_t ld_y,
                     int yg, int ys, int yo, float* mean, float* rstd, int M, int D, float eps, cudaStream_t stream);
int nv_ln_bwd_launch(const void* dy, int dy_is_bf16, int64_t ld_dy, int dyg, int dys, int dyo, const float* x, int64_t ld_x, int xg,
                     int xs, int xo, const float* mean, const float* rstd, const float* gamma, const float* dres,
                     int64_t ld_dres, float* dx, int64_t ld_dx, int dxg, int dxs, int dxo, bf16* dx_bf16,
                     int64_t ld_dxb, float* dgamma, float* dbeta, float* colsum, int M, int D, float side_drop_p,
                     uint64_t side_drop_seed, int side_drop_stream, const uint8_t* side_drop_bits, cudaStream_t stream);
int nv_cls_row_launch(const float* cls, const float* pos, float* x, int64_t batch_stride, int B, int D,
                      cudaStream_t stream);
int nv_patch_gather_ln_launch(const float* video, const int64_t* dims, const int64_t* strides, const int64_t* patch,
                              const float* gamma, const float* beta, void* out, int out_is_bf16, int64_t ld_out,
                              float* raw, float* mean, float* rstd, float eps, cudaStream_t stream);
int nv_ln_fold_launch(const float* W, const float* gamma, const float* beta, const float* b, void* Wf, int wf_is_bf16,
                      int64_t ld_wf, float* bias_f, int D, int P, cudaStream_t stream);
int nv_ln_fold_grads_launch(const float* G, int64_t ld_g, const float* W, const float* gamma, const float* beta,
                            const float* cs, float* dW, float* dgamma, float* dbeta, float* db, int D, int P,
                            cudaStream_t stream);
int nv_patch_ln_param_grad_launch(const float* video, const int64_t* dims, const int64_t* strides,
                                  const int64_t* patch, const float* dP, int64_t ld_dp, const float* mean,
                                  const float* rstd, float* dgamma, float* dbeta, cudaStream_t stream);
int nv_attn_tc_fwd_launch(const bf16* q, const bf16* k, const bf16* v, int64_t qkv_bs, int64_t qkv_rs, bf16* o,
                          int64_t o_bs, int64_t o_rs, float* lse, int B, int N, int H, int head_dim, float scale,
                          float dropout_p, uint64_t seed, uint32_t* drop_mask, int mask_ready, cudaStream_t stream);
int nv_attn_tc_bwd_launch(const bf16* q, const bf16* k, const bf16* v, int64_t qkv_bs, int64_t qkv_rs, const bf16* o,
                          const bf16* dO, int64_t o_bs, int64_t o_rs, const float* lse, float* delta_ws, bf16* dq,
                          bf16* dk, bf16* dv, int64_t d_bs, int64_t d_rs, int B, int N, int H, int head_dim,
                          float scale, float dropout_p, const uint32_t* drop_mask, cudaStream_t stream);
int nv_attn_cls_fwd_launch(const bf16* q, const bf16* k, const bf16* v, int64_t qkv_bs, int64_t qkv_rs, bf16* o,
                           int64_t o_bs, float* lse, int B, int N, int H, int head_dim, float scale, float dropout_p,
                           uint64_t seed, uint32_t* drop_mask, int mask_ready, cudaStream_t stream);
int nv_attn_cls_bwd_launch(const bf16* q, const bf16* k, const bf16* v, int64_t qkv_bs, int64_t qkv_rs, const bf16* o,
                           int64_t o_bs, const bf16* dO_cls, int64_t do_bs, const float* lse, bf16* dq, bf16* dk, bf16* dv,
                           int64_t d_bs, int64_t d_rs, int B, int N, int H, int head_dim, float scale, float dropout_p,
                           const uint32_t* drop_mask, cudaStream_t stream);
int nv_cast_f32_bf16_launch(const float* in, bf16* out, int64_t n, cudaStream_t stream);
int nv_cast_transpose_launch(const float* in, bf16* out, bf16* outT, int R, int C, cudaStream_t stream);
int nv_colsum_launch(const void* in, int in_is_bf16, int64_t ld, float* out, int M, int N, cudaStream_t stream);
int nv_batch_sum_launch(const float* in, int64_t batch_stride, float* out, int B, int64_t L, cudaStream_t stream);
int nv_mean_pool_fwd_launch(const float* x, float* pooled, int B, int N, int D, cudaStream_t stream);
int nv_mean_pool_bwd_launch(const float* dpooled, float* dx, bf16* dx_bf16, int B, int N, int D, cudaStream_t stream);
int nv_fmri_deinterleave_launch(const float* x, float* y, int B, int64_t S, int T, double* stats_ws, double eps,
                                cudaStream_t stream);
int nv_temporal_fwd_launch(const float* x, const float* params, float* out, float* seq_out, float* saved, int B,
                           int T, int F, float eps, const float* drop_p4, uint64_t seed, cudaStream_t stream);
int nv_temporal_bwd_launch(const float* x, const float* params, const float* saved, const float* dout,
                           const float* dseq, float* dparams_ws, float* dx, int B, int T, int F, float eps,
                           const float* drop_p4, uint64_t seed, cudaStream_t stream);

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int nv_version(void) { return NV_ABI_VERSION; }
int nv_set_sm_reserve(int n) { return nv_set_sm_reserve_impl(n); }
const char* nv_last_error(void) { return nv_last_error_impl(); }

int nv_device_check(void) {
  int dev = 0;
  NV_CUDA(cudaGetDevice(&dev));
  int major = 0;
  NV_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    nv_set_error("neurovit_b200 kernels are built for sm_100a only; current device has compute capability %d.x", major);
    return NV_ERR_UNSUPPORTED;
  }
  return NV_OK;
}

int nv_gemm_bf16(int a_mn, int b_mn, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb,
                 const float* bias, const float* residual, int64_t ld_res, const void* gelu_u, int64_t ld_u,
                 float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16, void* out_pre, int64_t ld_pre,
                 float* colsum, int apply_gelu, int accumulate, float alpha, int k_splits, int block_n, int cta_group,
                 float dropout_p, int64_t dropout_seed, int dropout_stream, const void* dropout_bits,
                 int dropout_row_mul, void* stream) {
  return nv_gemm_tc_launch(a_mn, b_mn, M, N, K, (const bf16*)A, lda, (const bf16*)B, ldb, bias, residual, ld_res,
                           (const bf16*)gelu_u, ld_u, out_f32, ld_f32, (bf16*)out_bf16, ld_bf16, (bf16*)out_pre,
                           ld_pre, colsum, apply_gelu, accumulate, alpha, k_splits, block_n, cta_group, dropout_p,
                           (uint64_t)dropout_seed, dropout_stream, (const uint8_t*)dropout_bits, dropout_row_mul,
                           ST(stream));
}

int nv_head_fwd(const float* x, int64_t ld_x, const float* gamma, const float* beta, const float* W, const float* bias,
                float* y, float* mean, float* rstd, float* logits, int B, int D, int C, float eps, void* stream) {
  return nv_head_fwd_launch(x, ld_x, gamma, beta, W, bias, y, mean, rstd, logits, B, D, C, eps, ST(stream));
}
int nv_head_bwd(const float* dl, const float* x, int64_t ld_x, const float* y, const float* mean, const float* rstd,
                const float* gamma, const float* W, float* dx, int64_t ld_dx, void* dx_bf16, int64_t ld_dxb, float* dW,
                float* db, float* dgamma, float* dbeta, int B, int D, int C, void* stream) {
  return nv_head_bwd_launch(dl, x, ld_x, y, mean, rstd, gamma, W, dx, ld_dx, (bf16*)dx_bf16, ld_dxb, dW, db, dgamma,
                            dbeta, B, D, C, ST(stream));
}

int nv_dropout_bits(void* out, int64_t n_groups, float p, int64_t seed, int stream_id, void* stream) {
  return nv_dropout_bits_launch((uint32_t*)out, n_groups, p, (uint64_t)seed, stream_id, ST(stream));
}

int nv_adamw_flat(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, const float* step_dev, void* stream) {
  return nv_adamw_flat_launch(p, g, m, v, (bf16*)p_bf16, n, lr, beta1, beta2, eps, weight_decay, step, step_dev,
                              ST(stream));
}
int nv_rng_epoch_advance(void* stream) { return nv_rng_epoch_advance_launch(ST(stream)); }
int nv_rng_epoch_get(unsigned long long* out_host, void* stream) {
  return nv_rng_epoch_read(reinterpret_cast<uint64_t*>(out_host), ST(stream));
}
int nv_counter_add(float* counter, float inc, void* stream) { return nv_counter_add_launch(counter, inc, ST(stream)); }

int nv_dropout(const float* in, int64_t ld_in, const float* residual, int64_t ld_res, float* out_f32, int64_t ld_f32,
               void* out_bf16, int64_t ld_bf16, float* colsum, int M, int N, float p, int64_t seed, int stream_id,
               int row_mul, void* stream) {
  return nv_dropout_launch(in, ld_in, residual, ld_res, out_f32, ld_f32, (bf16*)out_bf16, ld_bf16, colsum, M, N, p,
                           (uint64_t)seed, stream_id, row_mul, ST(stream));
}

int nv_gemm_f32(int M, int N, int K, int Z1, int Z2, const float* A, int64_t sa_m, int64_t sa_k, int64_t sa_z1,
                int64_t sa_z2, const float* B, int64_t sb_n, int64_t sb_k, int64_t sb_z1, int64_t sb_z2, float* C,
                int64_t sc_m, int64_t sc_z1, int64_t sc_z2, const float* bias, const float* residual, int64_t ld_res,
                const float* gelu_u, int64_t ld_u, float* out_pre, int64_t ld_pre, int apply_gelu, int accumulate,
                float alpha, void* stream) {
  return nv_simt_gemm_launch(M, N, K, Z1, Z2, A, sa_m, sa_k, sa_z1, sa_z2, B, sb_n, sb_k, sb_z1, sb_z2, C, sc_m,
                             sc_z1, sc_z2, bias, residual, ld_res, gelu_u, ld_u, out_pre, ld_pre, apply_gelu,
                             accumulate, alpha, ST(stream));
}

int nv_layernorm_fwd(const float* x, int64_t ld_x, int x_group, int x_gstride, int x_goff, const float* gamma,
                     const float* beta, const float* add, int64_t ld_add, int add_mod, int add_off, void* y,
                     int y_is_bf16, int64_t ld_y, int y_group, int y_gstride, int y_goff, float* mean, float* rstd,
                     int M, int D, float eps, void* stream) {
  return nv_ln_fwd_launch(x, ld_x, x_group, x_gstride, x_goff, gamma, beta, add, ld_add, add_mod, add_off, y,
                          y_is_bf16, ld_y, y_group, y_gstride, y_goff, mean, rstd, M, D, eps, ST(stream));
}

int nv_layernorm_bwd(const void* dy, int dy_is_bf16, int64_t ld_dy, int dy_group, int dy_gstride, int dy_goff, const float* x,
                     int64_t ld_x, int x_group, int x_gstride, int x_goff, const float* mean, const float* rstd,
                     const float* gamma, const float* dres, int64_t ld_dres, float* dx, int64_t ld_dx, int dx_group,
                     int dx_gstride, int dx_goff, void* dx_bf16, int64_t ld_dxb, float* dgamma, float* dbeta,
                     float* colsum, int M, int D, float side_drop_p, int64_t side_drop_seed, int side_drop_stream,
                     const void* side_drop_bits, void* stream) {
  return nv_ln_bwd_launch(dy, dy_is_bf16, ld_dy, dy_group, dy_gstride, dy_goff, x, ld_x, x_group, x_gstride, x_goff, mean, rstd,
                          gamma, dres, ld_dres, dx, ld_dx, dx_group, dx_gstride, dx_goff, (bf16*)dx_bf16, ld_dxb,
                          dgamma, dbeta, colsum, M, D, side_drop_p, (uint64_t)side_drop_seed, side_drop_stream,
                          (const uint8_t*)side_drop_bits, ST(stream));
}

int nv_cls_row(const float* cls, const float* pos, float* x, int64_t batch_stride, int B, int D, void* stream) {
  return nv_cls_row_launch(cls, pos, x, batch_stride, B, D, ST(stream));
}

int nv_patch_gather_ln(const float* video, const int64_t* dims, const int64_t* strides, const int64_t* patch,
                       const float* gamma, const float* beta, void* out, int out_is_bf16, int64_t ld_out, float* raw,
                       float* mean, float* rstd, float eps, void* stream) {
  return nv_patch_gather_ln_launch(video, dims, strides, patch, gamma, beta, out, out_is_bf16, ld_out, raw, mean,
                                   rstd, eps, ST(stream));
}

int nv_ln_fold(const float* W, const float* gamma, const float* beta, const float* b, void* Wf, int wf_is_bf16,
               int64_t ld_wf, float* bias_f, int D, int P, void* stream) {
  return nv_ln_fold_launch(W, gamma, beta, b, Wf, wf_is_bf16, ld_wf, bias_f, D, P, ST(stream));
}
int nv_ln_fold_grads(const float* G, int64_t ld_g, const float* W, const float* gamma, const float* beta, const float* cs,
                     float* dW, float* dgamma, float* dbeta, float* db, int D, int P, void* stream) {
  return nv_ln_fold_grads_launch(G, ld_g, W, gamma, beta, cs, dW, dgamma, dbeta, db, D, P, ST(stream));
}

int nv_patch_ln_param_grad(const float* video, const int64_t* dims, const int64_t* strides, const int64_t* patch,
                           const float* dP, int64_t ld_dp, const float* mean, const float* rstd, float* dgamma,
                           float* dbeta, void* stream) {
  return nv_patch_ln_param_grad_launch(video, dims, strides, patch, dP, ld_dp, mean, rstd, dgamma, dbeta, ST(stream));
}

int nv_attention_fwd(const void* q, const void* k, const void* v, int64_t qkv_batch_stride, int64_t qkv_row_stride,
                     void* o, int64_t o_batch_stride, int64_t o_row_stride, float* lse, int B, int N, int H,
                     int head_dim, float scale, float dropout_p, int64_t seed, void* drop_mask, int drop_mask_ready,
                     void* stream) {
  return nv_attn_tc_fwd_launch((const bf16*)q, (const bf16*)k, (const bf16*)v, qkv_batch_stride, qkv_row_stride,
                               (bf16*)o, o_batch_stride, o_row_stride, lse, B, N, H, head_dim, scale, dropout_p,
                               (uint64_t)seed, (uint32_t*)drop_mask, drop_mask_ready, ST(stream));
}

int nv_attention_cls_fwd(const void* q, const void* k, const void* v, int64_t qkv_batch_stride, int64_t qkv_row_stride,
                         void* o_cls, int64_t o_batch_stride, float* lse, int B, int N, int H, int head_dim, float scale,
                         float dropout_p, int64_t seed, void* drop_mask, int drop_mask_ready, void* stream) {
  return nv_attn_cls_fwd_launch((const bf16*)q, (const bf16*)k, (const bf16*)v, qkv_batch_stride, qkv_row_stride,
                                (bf16*)o_cls, o_batch_stride, lse, B, N, H, head_dim, scale, dropout_p, (uint64_t)seed,
                                (uint32_t*)drop_mask, drop_mask_ready, ST(stream));
}

int nv_attention_bwd(const void* q, const void* k, const void* v, int64_t qkv_batch_stride, int64_t qkv_row_stride,
                     const void* o, const void* dO, int64_t o_batch_stride, int64_t o_row_stride, const float* lse,
                     float* delta_ws, void* dq, void* dk, void* dv, int64_t dqkv_batch_stride,
                     int64_t dqkv_row_stride, int B, int N, int H, int head_dim, float scale, float dropout_p,
                     const void* drop_mask, void* stream) {
  return nv_attn_tc_bwd_launch((const bf16*)q, (const bf16*)k, (const bf16*)v, qkv_batch_stride, qkv_row_stride,
                               (const bf16*)o, (const bf16*)dO, o_batch_stride, o_row_stride, lse, delta_ws,
                               (bf16*)dq, (bf16*)dk, (bf16*)dv, dqkv_batch_stride, dqkv_row_stride, B, N, H, head_dim,
                               scale, dropout_p, (const uint32_t*)drop_mask, ST(stream));
}

int nv_attention_cls_bwd(const void* q, const void* k, const void* v, int64_t qkv_batch_stride, int64_t qkv_row_stride,
                         const void* o, int64_t o_batch_stride, const void* dO_cls, int64_t dO_batch_stride,
                         const float* lse, void* dq, void* dk, void* dv, int64_t dqkv_batch_stride,
                         int64_t dqkv_row_stride, int B, int N, int H, int head_dim, float scale, float dropout_p,
                         const void* drop_mask, void* stream) {
  return nv_attn_cls_bwd_launch((const bf16*)q, (const bf16*)k, (const bf16*)v, qkv_batch_stride, qkv_row_stride,
                                (const bf16*)o, o_batch_stride, (const bf16*)dO_cls, dO_batch_stride, lse, (bf16*)dq,
                                (bf16*)dk, (bf16*)dv, dqkv_batch_stride, dqkv_row_stride, B, N, H, head_dim, scale,
                                dropout_p, (const uint32_t*)drop_mask, ST(stream));
}

int nv_softmax_fwd(float* s, int64_t rows, int n, void* stream) { return nv_softmax_fwd_launch(s, rows, n, ST(stream)); }
int nv_softmax_bwd(const float* P, float* dP, int64_t rows, int n, void* stream) {
  return nv_softmax_bwd_launch(P, dP, rows, n, ST(stream));
}

int nv_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream) {
  return nv_cast_f32_bf16_launch(in, (bf16*)out, n, ST(stream));
}
int nv_cast_transpose_f32_bf16(const float* in, void* out, void* outT, int R, int C, void* stream) {
  return nv_cast_transpose_launch(in, (bf16*)out, (bf16*)outT, R, C, ST(stream));
}
int nv_colsum(const void* in, int in_is_bf16, int64_t ld, float* out, int M, int N, void* stream) {
  return nv_colsum_launch(in, in_is_bf16, ld, out, M, N, ST(stream));
}
int nv_batch_sum(const float* in, int64_t batch_stride, float* out, int B, int64_t L, void* stream) {
  return nv_batch_sum_launch(in, batch_stride, out, B, L, ST(stream));
}
int nv_mean_pool_fwd(const float* x, float* pooled, int B, int N, int D, void* stream) {
  return nv_mean_pool_fwd_launch(x, pooled, B, N, D, ST(stream));
}
int nv_mean_pool_bwd(const float* dpooled, float* dx, void* dx_bf16, int B, int N, int D, void* stream) {
  return nv_mean_pool_bwd_launch(dpooled, dx, (bf16*)dx_bf16, B, N, D, ST(stream));
}

int nv_fmri_deinterleave(const float* x, float* y, int B, int64_t S, int T, void* stats_ws, double eps, void* stream) {
  return nv_fmri_deinterleave_launch(x, y, B, S, T, static_cast<double*>(stats_ws), eps, ST(stream));
}

int nv_temporal_fwd(const float* x, const float* params, float* out, float* seq_out, float* saved, int B, int T,
                    int F, float eps, float p_attn, float p_drop1, float p_ffn, float p_drop2, int64_t seed,
                    void* stream) {
  const float p4[4] = {p_attn, p_drop1, p_ffn, p_drop2};
  return nv_temporal_fwd_launch(x, params, out, seq_out, saved, B, T, F, eps, p4, (uint64_t)seed, ST(stream));
}
int nv_temporal_bwd(const float* x, const float* params, const float* saved, const float* dout, const float* dseq,
                    float* dparams_ws, float* dx, int B, int T, int F, float eps, float p_attn, float p_drop1,
                    float p_ffn, float p_drop2, int64_t seed, void* stream) {
  const float p4[4] = {p_attn, p_drop1, p_ffn, p_drop2};
  return nv_temporal_bwd_launch(x, params, saved, dout, dseq, dparams_ws, dx, B, T, F, eps, p4, (uint64_t)seed,
                                ST(stream));
}

}  // extern "C"
