// 4D NeuroEncoder temporal head: one CTA per fMRI sequence, everything in shared memory / registers.
// Reference: src/models/NeuroEncoder.py:63-66 (temporal_transformer -> mean over T -> projection_head),
// :207-217 TemporalTransformer = nn.TransformerEncoder(nn.TransformerEncoderLayer(d_model=2, nhead=2,
// batch_first=True), num_layers=1): post-norm, ReLU, dim_feedforward F (2048), LN eps 1e-5, head_dim 1
// (softmax scale 1), and :219-230 ProjectionHead = Linear(2,2).
//   qkv = W_in x + b_in;  per head h: o[t,h] = sum_s softmax_s(q[t,h] k[s,h]) v[s,h]
//   x1 = LN1(x + W_o o + b_o);  x2 = LN2(x1 + W_2 relu(W_1 x1 + b_1) + b_2);  out = W_p mean_t(x2) + b_p
// LayerNorm over 2 elements is degenerate (xhat = +-d/sqrt(d^2+eps)), so eps is honoured exactly.
// Packed parameter vector (fp32), F = dim_feedforward:
//   in_w[6,2]@0  in_b[6]@12  out_w[2,2]@18  out_b[2]@22  l1_w[F,2]@24  l1_b[F]@24+2F  l2_w[2,F]@24+3F
//   l2_b[2]@24+5F  n1_w@26+5F n1_b@28+5F n2_w@30+5F n2_b@32+5F  ph_w[2,2]@34+5F  ph_b[2]@38+5F ; P=40+5F
#include "nv_common.cuh"
#include "nv_rng.cuh"

namespace {

constexpr int TT = 256;  // threads per CTA

// The four nn.Dropout sites of nn.TransformerEncoderLayer in training mode (torch defaults p = 0.1):
//   0 attention probabilities [B, 2, T, T]   1 dropout1 on the self-attention output [B, T, 2]
//   2 dropout on relu(linear1) [B, T, F]     3 dropout2 on linear2's output [B, T, 2]
// Masks are Philox bits of (seed, site, element index) with rows padded to a multiple of 8 elements
// (Tp, Fp), so forward and backward regenerate them and ops.dropout_keep_mask can replay them in tests.
struct TDrop {
  uint32_t thr[4];
  float ks[4];
  uint64_t seed;
  const uint64_t* epoch;  // device epoch counter added to the seed (CUDA-graph replays)
};
// keep-scale (0 or 1/(1-p)) of one element
__device__ __forceinline__ float tkeep(const TDrop& d, int site, uint64_t idx) {
  if (d.thr[site] == 0) return 1.0f;
  const uint32_t b8 = nv_keep_bits8(d.seed, idx >> 3, (uint32_t)site, d.thr[site]);
  return (b8 >> (idx & 7)) & 1u ? d.ks[site] : 0.f;
}
// keep-scale of element (row, i) while i walks a row in order: one Philox call per 8 elements
struct RowMask {
  const TDrop& d;
  int site;
  uint64_t base;  // row * padded row length
  uint32_t b8;
  __device__ __forceinline__ RowMask(const TDrop& d_, int site_, uint64_t base_) : d(d_), site(site_), base(base_), b8(0) {}
  __device__ __forceinline__ float operator()(int i) {
    if (d.thr[site] == 0) return 1.0f;
    if ((i & 7) == 0) b8 = nv_keep_bits8(d.seed, (base + i) >> 3, (uint32_t)site, d.thr[site]);
    return (b8 >> (i & 7)) & 1u ? d.ks[site] : 0.f;
  }
};

struct TParams {
  const float *in_w, *in_b, *out_w, *out_b, *l1_w, *l1_b, *l2_w, *l2_b, *n1_w, *n1_b, *n2_w, *n2_b, *ph_w, *ph_b;
  __device__ TParams(const float* p, int F) {
    in_w = p; in_b = p + 12; out_w = p + 18; out_b = p + 22; l1_w = p + 24; l1_b = p + 24 + 2 * F;
    l2_w = p + 24 + 3 * F; l2_b = p + 24 + 5 * F; n1_w = p + 26 + 5 * F; n1_b = p + 28 + 5 * F;
    n2_w = p + 30 + 5 * F; n2_b = p + 32 + 5 * F; ph_w = p + 34 + 5 * F; ph_b = p + 38 + 5 * F;
  }
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < TT / 32; ++i) s += red[i];
  return s;
}

__device__ __forceinline__ void ln2_fwd(float y0, float y1, const float* w, const float* b, float eps, float& o0,
                                        float& o1) {
  const float mean = 0.5f * (y0 + y1);
  const float d0 = y0 - mean, d1 = y1 - mean;
  const float rstd = rsqrtf(0.5f * (d0 * d0 + d1 * d1) + eps);
  o0 = d0 * rstd * w[0] + b[0];
  o1 = d1 * rstd * w[1] + b[1];
}
// returns dy for LN over 2 elements; xh = normalised input, accumulates nothing
__device__ __forceinline__ void ln2_bwd(float y0, float y1, const float* w, float eps, float g0, float g1, float& dy0,
                                        float& dy1, float& xh0, float& xh1) {
  const float mean = 0.5f * (y0 + y1);
  const float d0 = y0 - mean, d1 = y1 - mean;
  const float rstd = rsqrtf(0.5f * (d0 * d0 + d1 * d1) + eps);
  xh0 = d0 * rstd; xh1 = d1 * rstd;
  const float gh0 = g0 * w[0], gh1 = g1 * w[1];
  const float m1 = 0.5f * (gh0 + gh1);
  const float m2 = 0.5f * (gh0 * xh0 + gh1 * xh1);
  dy0 = rstd * (gh0 - m1 - xh0 * m2);
  dy1 = rstd * (gh1 - m1 - xh1 * m2);
}

// shared: xs, q, k, v, o : [T][2] each
__device__ void temporal_attn_fwd(const float* xb, const TParams& P, int T, float* xs, float* q, float* k, float* v,
                                  float* o, const TDrop& drop, int b) {
  const int Tp = (T + 7) & ~7;
  for (int t = threadIdx.x; t < T; t += TT) {
    const float x0 = xb[2 * t], x1 = xb[2 * t + 1];
    xs[2 * t] = x0; xs[2 * t + 1] = x1;
    float r[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) r[i] = P.in_w[2 * i] * x0 + P.in_w[2 * i + 1] * x1 + P.in_b[i];
    q[2 * t] = r[0]; q[2 * t + 1] = r[1];
    k[2 * t] = r[2]; k[2 * t + 1] = r[3];
    v[2 * t] = r[4]; v[2 * t + 1] = r[5];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * T; idx += TT) {
    const int hh = idx & 1;
    const float qv = q[idx];
    float mx = -INFINITY;
    for (int s = 0; s < T; ++s) mx = fmaxf(mx, qv * k[2 * s + hh]);
    float sum = 0.f, acc = 0.f;
    RowMask mask(drop, 0, (((uint64_t)b * 2 + hh) * T + (idx >> 1)) * Tp);
    for (int s = 0; s < T; ++s) {
      const float e = expf(qv * k[2 * s + hh] - mx);
      sum += e;
      acc += e * mask(s) * v[2 * s + hh];  // dropout on the normalised weights = masked numerator
    }
    o[idx] = acc / sum;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TT)
temporal_fwd_kernel(const float* __restrict__ x, const float* __restrict__ params, float* __restrict__ out,
                    float* __restrict__ seq_out, float* __restrict__ saved, int T, int F, float eps, const TDrop drop_in) {
  TDrop drop = drop_in;
  drop.seed = nv_seed(drop_in.seed, drop_in.epoch);
  extern __shared__ float sm[];
  __shared__ float red[TT / 32];
  float *xs = sm, *q = sm + 2 * T, *k = sm + 4 * T, *v = sm + 6 * T, *o = sm + 8 * T;
  const int b = blockIdx.x;
  const TParams P(params, F);
  temporal_attn_fwd(x + (int64_t)b * T * 2, P, T, xs, q, k, v, o, drop, b);
  const int Fp = (F + 7) & ~7;
  float m0 = 0.f, m1 = 0.f;
  for (int t = threadIdx.x; t < T; t += TT) {
    const uint64_t bt = (uint64_t)b * T + t;
    const float a0 = (P.out_w[0] * o[2 * t] + P.out_w[1] * o[2 * t + 1] + P.out_b[0]) * tkeep(drop, 1, bt * 2);
    const float a1 = (P.out_w[2] * o[2 * t] + P.out_w[3] * o[2 * t + 1] + P.out_b[1]) * tkeep(drop, 1, bt * 2 + 1);
    float u0, u1;
    ln2_fwd(xs[2 * t] + a0, xs[2 * t + 1] + a1, P.n1_w, P.n1_b, eps, u0, u1);
    float f0 = P.l2_b[0], f1 = P.l2_b[1];
    RowMask hmask(drop, 2, bt * Fp);
    for (int j = 0; j < F; ++j) {
      const float h = fmaxf(P.l1_w[2 * j] * u0 + P.l1_w[2 * j + 1] * u1 + P.l1_b[j], 0.f) * hmask(j);
      f0 += P.l2_w[j] * h;
      f1 += P.l2_w[F + j] * h;
    }
    const float y0 = u0 + f0 * tkeep(drop, 3, bt * 2), y1 = u1 + f1 * tkeep(drop, 3, bt * 2 + 1);
    float* sv = saved + ((int64_t)b * T + t) * 4;
    sv[0] = u0; sv[1] = u1; sv[2] = y0; sv[3] = y1;
    float z0, z1;
    ln2_fwd(y0, y1, P.n2_w, P.n2_b, eps, z0, z1);
    if (seq_out) { seq_out[((int64_t)b * T + t) * 2] = z0; seq_out[((int64_t)b * T + t) * 2 + 1] = z1; }
    m0 += z0; m1 += z1;
  }
  m0 = block_sum(m0, red) / (float)T;
  m1 = block_sum(m1, red) / (float)T;
  if (threadIdx.x == 0 && out != nullptr) {
    out[2 * b] = P.ph_w[0] * m0 + P.ph_w[1] * m1 + P.ph_b[0];
    out[2 * b + 1] = P.ph_w[2] * m0 + P.ph_w[3] * m1 + P.ph_b[1];
  }
}

__global__ void __launch_bounds__(TT)
temporal_bwd_kernel(const float* __restrict__ x, const float* __restrict__ params, const float* __restrict__ saved,
                    const float* __restrict__ dout, const float* __restrict__ dseq, float* __restrict__ dparams_ws,
                    float* __restrict__ dx_out, int T, int F, float eps, const TDrop drop_in) {
  TDrop drop = drop_in;
  drop.seed = nv_seed(drop_in.seed, drop_in.epoch);
  extern __shared__ float sm[];
  __shared__ float red[TT / 32];
  // [T][2] arrays: xs q k v o | x1 df dx1 da dq dk dv
  float *xs = sm, *q = sm + 2 * T, *k = sm + 4 * T, *v = sm + 6 * T, *o = sm + 8 * T;
  float *x1 = sm + 10 * T, *df = sm + 12 * T, *dx1 = sm + 14 * T, *da = sm + 16 * T;
  float *dq = sm + 18 * T, *dk = sm + 20 * T, *dv = sm + 22 * T, *dxs = sm + 24 * T;
  float* dfm = sm + 26 * T;  // df through dropout2 (gradient of linear2's output)
  const int Tp = (T + 7) & ~7, Fp = (F + 7) & ~7;
  const int b = blockIdx.x;
  const TParams P(params, F);
  const int PN = 40 + 5 * F;
  float* G = dparams_ws + (int64_t)b * PN;
  float* g_in_w = G; float* g_in_b = G + 12; float* g_out_w = G + 18; float* g_out_b = G + 22;
  float* g_l1_w = G + 24; float* g_l1_b = G + 24 + 2 * F; float* g_l2_w = G + 24 + 3 * F; float* g_l2_b = G + 24 + 5 * F;
  float* g_n1_w = G + 26 + 5 * F; float* g_n1_b = G + 28 + 5 * F; float* g_n2_w = G + 30 + 5 * F;
  float* g_n2_b = G + 32 + 5 * F; float* g_ph_w = G + 34 + 5 * F; float* g_ph_b = G + 38 + 5 * F;

  temporal_attn_fwd(x + (int64_t)b * T * 2, P, T, xs, q, k, v, o, drop, b);

  // ---- projection head + mean + LN2 backward ----
  const float do0 = dout ? dout[2 * b] : 0.f, do1 = dout ? dout[2 * b + 1] : 0.f;
  float m0 = 0.f, m1 = 0.f;
  for (int t = threadIdx.x; t < T; t += TT) {
    const float* sv = saved + ((int64_t)b * T + t) * 4;
    float z0, z1;
    ln2_fwd(sv[2], sv[3], P.n2_w, P.n2_b, eps, z0, z1);
    m0 += z0; m1 += z1;
  }
  m0 = block_sum(m0, red) / (float)T;
  m1 = block_sum(m1, red) / (float)T;
  if (threadIdx.x == 0) {
    g_ph_w[0] = do0 * m0; g_ph_w[1] = do0 * m1; g_ph_w[2] = do1 * m0; g_ph_w[3] = do1 * m1;
    g_ph_b[0] = do0; g_ph_b[1] = do1;
  }
  const float dzm0 = (P.ph_w[0] * do0 + P.ph_w[2] * do1) / (float)T;
  const float dzm1 = (P.ph_w[1] * do0 + P.ph_w[3] * do1) / (float)T;
  float gw0 = 0.f, gw1 = 0.f, gb0 = 0.f, gb1 = 0.f, sb0 = 0.f, sb1 = 0.f;
  for (int t = threadIdx.x; t < T; t += TT) {
    const float* sv = saved + ((int64_t)b * T + t) * 4;
    x1[2 * t] = sv[0]; x1[2 * t + 1] = sv[1];
    const float dz0 = dzm0 + (dseq ? dseq[((int64_t)b * T + t) * 2] : 0.f);
    const float dz1 = dzm1 + (dseq ? dseq[((int64_t)b * T + t) * 2 + 1] : 0.f);
    float dy0, dy1, xh0, xh1;
    ln2_bwd(sv[2], sv[3], P.n2_w, eps, dz0, dz1, dy0, dy1, xh0, xh1);
    gw0 += dz0 * xh0; gw1 += dz1 * xh1; gb0 += dz0; gb1 += dz1;
    df[2 * t] = dy0; df[2 * t + 1] = dy1;           // residual path x1 -> y2
    const uint64_t bt = (uint64_t)b * T + t;
    const float e0 = dy0 * tkeep(drop, 3, bt * 2), e1 = dy1 * tkeep(drop, 3, bt * 2 + 1);
    dfm[2 * t] = e0; dfm[2 * t + 1] = e1;           // through dropout2 into linear2
    sb0 += e0; sb1 += e1;
  }
  gw0 = block_sum(gw0, red); gw1 = block_sum(gw1, red);
  gb0 = block_sum(gb0, red); gb1 = block_sum(gb1, red);
  sb0 = block_sum(sb0, red); sb1 = block_sum(sb1, red);
  if (threadIdx.x == 0) {
    g_n2_w[0] = gw0; g_n2_w[1] = gw1; g_n2_b[0] = gb0; g_n2_b[1] = gb1;
    g_l2_b[0] = sb0; g_l2_b[1] = sb1;
  }
  __syncthreads();
  // ---- feed-forward backward: parameter grads (thread per hidden unit j) ----
  for (int j = threadIdx.x; j < F; j += TT) {
    const float w0 = P.l1_w[2 * j], w1 = P.l1_w[2 * j + 1], bj = P.l1_b[j];
    const float v0 = P.l2_w[j], v1 = P.l2_w[F + j];
    float gw_0 = 0.f, gw_1 = 0.f, gbj = 0.f, gv0 = 0.f, gv1 = 0.f;
    for (int t = 0; t < T; ++t) {
      const float u0 = x1[2 * t], u1 = x1[2 * t + 1];
      const float pre = w0 * u0 + w1 * u1 + bj;
      if (pre > 0.f) {
        const float mk = tkeep(drop, 2, ((uint64_t)b * T + t) * Fp + j);  // hidden-unit dropout mask
        const float d0 = dfm[2 * t], d1 = dfm[2 * t + 1];
        const float dh = (d0 * v0 + d1 * v1) * mk;
        gv0 += d0 * pre * mk; gv1 += d1 * pre * mk;
        gw_0 += dh * u0; gw_1 += dh * u1; gbj += dh;
      }
    }
    g_l1_w[2 * j] = gw_0; g_l1_w[2 * j + 1] = gw_1; g_l1_b[j] = gbj;
    g_l2_w[j] = gv0; g_l2_w[F + j] = gv1;
  }
  // ---- feed-forward backward: input grads (thread per timepoint), then LN1 backward ----
  float n1w0 = 0.f, n1w1 = 0.f, n1b0 = 0.f, n1b1 = 0.f;
  float ow[4] = {0.f, 0.f, 0.f, 0.f}, ob0 = 0.f, ob1 = 0.f;
  for (int t = threadIdx.x; t < T; t += TT) {
    const float u0 = x1[2 * t], u1 = x1[2 * t + 1];
    const float d0 = dfm[2 * t], d1 = dfm[2 * t + 1];
    float g0 = df[2 * t], g1 = df[2 * t + 1];  // residual path x1 -> y2
    const uint64_t bt = (uint64_t)b * T + t;
    RowMask hmask(drop, 2, bt * Fp);
    for (int j = 0; j < F; ++j) {
      const float w0 = P.l1_w[2 * j], w1 = P.l1_w[2 * j + 1];
      const float pre = w0 * u0 + w1 * u1 + P.l1_b[j];
      const float mk = hmask(j);
      if (pre > 0.f) {
        const float dh = (d0 * P.l2_w[j] + d1 * P.l2_w[F + j]) * mk;
        g0 += dh * w0; g1 += dh * w1;
      }
    }
    // LN1: input y1 = x + dropout1(W_o o + b_o)
    const float k0 = tkeep(drop, 1, bt * 2), k1 = tkeep(drop, 1, bt * 2 + 1);
    const float a0 = (P.out_w[0] * o[2 * t] + P.out_w[1] * o[2 * t + 1] + P.out_b[0]) * k0;
    const float a1 = (P.out_w[2] * o[2 * t] + P.out_w[3] * o[2 * t + 1] + P.out_b[1]) * k1;
    float dy0, dy1, xh0, xh1;
    ln2_bwd(xs[2 * t] + a0, xs[2 * t + 1] + a1, P.n1_w, eps, g0, g1, dy0, dy1, xh0, xh1);
    n1w0 += g0 * xh0; n1w1 += g1 * xh1; n1b0 += g0; n1b1 += g1;
    dxs[2 * t] = dy0; dxs[2 * t + 1] = dy1;  // residual path x -> y1
    dy0 *= k0; dy1 *= k1;                    // through dropout1 into the attention output projection
    ow[0] += dy0 * o[2 * t]; ow[1] += dy0 * o[2 * t + 1]; ow[2] += dy1 * o[2 * t]; ow[3] += dy1 * o[2 * t + 1];
    ob0 += dy0; ob1 += dy1;
    // d o = W_o^T dy
    da[2 * t] = P.out_w[0] * dy0 + P.out_w[2] * dy1;
    da[2 * t + 1] = P.out_w[1] * dy0 + P.out_w[3] * dy1;
    dk[2 * t] = 0.f; dk[2 * t + 1] = 0.f; dv[2 * t] = 0.f; dv[2 * t + 1] = 0.f;
  }
  n1w0 = block_sum(n1w0, red); n1w1 = block_sum(n1w1, red);
  n1b0 = block_sum(n1b0, red); n1b1 = block_sum(n1b1, red);
#pragma unroll
  for (int i = 0; i < 4; ++i) ow[i] = block_sum(ow[i], red);
  ob0 = block_sum(ob0, red); ob1 = block_sum(ob1, red);
  if (threadIdx.x == 0) {
    g_n1_w[0] = n1w0; g_n1_w[1] = n1w1; g_n1_b[0] = n1b0; g_n1_b[1] = n1b1;
#pragma unroll
    for (int i = 0; i < 4; ++i) g_out_w[i] = ow[i];
    g_out_b[0] = ob0; g_out_b[1] = ob1;
  }
  __syncthreads();
  // ---- attention backward, one (t, head) row per thread ----
  for (int idx = threadIdx.x; idx < 2 * T; idx += TT) {
    const int hh = idx & 1;
    const float qv = q[idx], dov = da[idx];
    float mx = -INFINITY;
    for (int s = 0; s < T; ++s) mx = fmaxf(mx, qv * k[2 * s + hh]);
    float sum = 0.f;
    for (int s = 0; s < T; ++s) sum += expf(qv * k[2 * s + hh] - mx);
    const float inv = 1.0f / sum;
    const float delta = dov * o[idx];  // = sum_s p_s dP_s also with dropout (o already holds the masked sum)
    float dqa = 0.f;
    RowMask mask(drop, 0, (((uint64_t)b * 2 + hh) * T + (idx >> 1)) * Tp);
    for (int s = 0; s < T; ++s) {
      const float p = expf(qv * k[2 * s + hh] - mx) * inv;
      const float mk = mask(s);
      const float ds = p * (dov * v[2 * s + hh] * mk - delta);
      dqa += ds * k[2 * s + hh];
      atomicAdd(&dk[2 * s + hh], ds * qv);
      atomicAdd(&dv[2 * s + hh], p * mk * dov);
    }
    dq[idx] = dqa;
  }
  __syncthreads();
  // ---- in_proj backward ----
  float iw[12], ib[6];
#pragma unroll
  for (int i = 0; i < 12; ++i) iw[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) ib[i] = 0.f;
  for (int t = threadIdx.x; t < T; t += TT) {
    const float r[6] = {dq[2 * t], dq[2 * t + 1], dk[2 * t], dk[2 * t + 1], dv[2 * t], dv[2 * t + 1]};
    const float x0 = xs[2 * t], x1v = xs[2 * t + 1];
    float gx0 = dxs[2 * t], gx1 = dxs[2 * t + 1];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      iw[2 * i] += r[i] * x0; iw[2 * i + 1] += r[i] * x1v; ib[i] += r[i];
      gx0 += P.in_w[2 * i] * r[i]; gx1 += P.in_w[2 * i + 1] * r[i];
    }
    if (dx_out) { dx_out[((int64_t)b * T + t) * 2] = gx0; dx_out[((int64_t)b * T + t) * 2 + 1] = gx1; }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) iw[i] = block_sum(iw[i], red);
#pragma unroll
  for (int i = 0; i < 6; ++i) ib[i] = block_sum(ib[i], red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 12; ++i) g_in_w[i] = iw[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) g_in_b[i] = ib[i];
  }
}

int fill_tdrop(TDrop& d, const float* p4, uint64_t seed) {
  for (int i = 0; i < 4; ++i) {
    const float p = p4 ? p4[i] : 0.f;
    NV_REQUIRE(p >= 0.f && p < 1.f, "temporal: dropout p[%d] = %f out of range [0, 1)", i, p);
    d.thr[i] = nv_dropout_threshold(p);
    d.ks[i] = nv_dropout_keep_scale(d.thr[i]);
  }
  d.seed = seed;
  bool any = false;
  for (int i = 0; i < 4; ++i) any = any || d.thr[i] != 0;
  d.epoch = any ? nv_rng_epoch_dev() : nullptr;
  return NV_OK;
}

}  // namespace

int nv_temporal_fwd_launch(const float* x, const float* params, float* out, float* seq_out, float* saved, int B,
                           int T, int F, float eps, const float* drop_p4, uint64_t seed, cudaStream_t stream) {
  NV_REQUIRE(B >= 0 && T > 0 && F > 0 && T <= 2048, "temporal: bad sizes B=%d T=%d F=%d", B, T, F);
  TDrop drop;
  int st = fill_tdrop(drop, drop_p4, seed);
  if (st != NV_OK) return st;
  if (B == 0) return NV_OK;
  const size_t smem = (size_t)10 * T * sizeof(float);
  if (smem > 48 * 1024)
    NV_CUDA(cudaFuncSetAttribute(temporal_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  temporal_fwd_kernel<<<B, TT, smem, stream>>>(x, params, out, seq_out, saved, T, F, eps, drop);
  NV_LAUNCH_CHECK("temporal_fwd_kernel");
  return NV_OK;
}

int nv_temporal_bwd_launch(const float* x, const float* params, const float* saved, const float* dout,
                           const float* dseq, float* dparams_ws, float* dx, int B, int T, int F, float eps,
                           const float* drop_p4, uint64_t seed, cudaStream_t stream) {
  NV_REQUIRE(B >= 0 && T > 0 && F > 0 && T <= 2048, "temporal: bad sizes B=%d T=%d F=%d", B, T, F);
  TDrop drop;
  int st = fill_tdrop(drop, drop_p4, seed);
  if (st != NV_OK) return st;
  if (B == 0) return NV_OK;
  const size_t smem = (size_t)28 * T * sizeof(float);
  if (smem > 48 * 1024)
    NV_CUDA(cudaFuncSetAttribute(temporal_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  temporal_bwd_kernel<<<B, TT, smem, stream>>>(x, params, saved, dout, dseq, dparams_ws, dx, T, F, eps, drop);
  NV_LAUNCH_CHECK("temporal_bwd_kernel");
  return NV_OK;
}
