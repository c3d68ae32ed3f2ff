"""Tensor-level wrappers over the C ABI: PyTorch supplies device memory and the current stream, the
library does all arithmetic. No torch math ops are used on the data path here."""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32


class _KernelProfile:
    """CUDA-event timing of the dominant kernel (the tcgen05 GEMM) on the launching stream, used by bench.py
    for the roofline object. Off by default; events are only recorded while `enabled`."""

    def __init__(self):
        self.enabled = False
        self.events = []

    def reset(self):
        self.events = []

    def region(self, flops):
        prof = self

        class _R:
            def __enter__(self_r):
                self_r.e0 = torch.cuda.Event(enable_timing=True)
                self_r.e1 = torch.cuda.Event(enable_timing=True)
                self_r.e0.record()

            def __exit__(self_r, *exc):
                self_r.e1.record()
                prof.events.append((self_r.e0, self_r.e1, flops))
                return False

        return _R()

    def summary(self):
        """(total ms, total flops, launches) — call after torch.cuda.synchronize()."""
        ms = sum(e0.elapsed_time(e1) for e0, e1, _ in self.events)
        return ms, sum(f for _, _, f in self.events), len(self.events)


PROFILE = _KernelProfile()

# tcgen05 CTA-group mode of the GEMM: 0 = auto (CTA pairs, 256-row tiles), 1 = single CTA, 2 = CTA pair
GEMM_CTA_GROUP = int(os.environ.get("NEUROVIT_GEMM_CG", "0"))


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise _lib.NeuroViTLibraryError(
            "neurovit_b200 runs on CUDA (sm_100a) only; got a CPU tensor and there is no CPU fallback")
    _lib.require_device(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _rowmajor(t: torch.Tensor, name: str) -> int:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} "
                         f"strides {t.stride()}")
    return t.stride(0)


# ------------------------------------------------------------------------------------------- GEMMs
def gemm_bf16(a, b, *, a_mn=False, b_mn=False, bias=None, residual=None, gelu_u=None, out_f32=None,
              out_bf16=None, out_pre=None, colsum=None, apply_gelu=False, accumulate=False, alpha=1.0, k_splits=1,
              block_n=0, cta_group=None, dropout=None):
    """C[M,N] = epilogue(alpha * A @ B^T) on tcgen05. a: [M,K] (or [K,M] if a_mn); b: [N,K] (or [K,N]).
    dropout = (p, seed, stream) applies nn.Dropout to the value before the residual add."""
    _dev(a)
    assert a.dtype == BF16 and b.dtype == BF16, "gemm_bf16 operands must be bfloat16"
    lda, ldb = _rowmajor(a, "a"), _rowmajor(b, "b")
    (K, M) = a.shape if a_mn else (a.shape[1], a.shape[0])
    (Kb, N) = b.shape if b_mn else (b.shape[1], b.shape[0])
    if K != Kb:
        raise ValueError(f"gemm_bf16: reduction dims differ ({K} vs {Kb})")
    for t, nm, dt in ((out_f32, "out_f32", F32), (out_bf16, "out_bf16", BF16), (out_pre, "out_pre", BF16),
                      (residual, "residual", F32), (gelu_u, "gelu_u", BF16)):
        if t is not None:
            assert t.dtype == dt and tuple(t.shape) == (M, N), f"{nm}: expected {dt} [{M},{N}], got {t.dtype} {tuple(t.shape)}"
            _rowmajor(t, nm)
    if bias is not None:
        assert bias.dtype == F32 and bias.numel() == N and bias.is_contiguous()
    ld = lambda t: 0 if t is None else t.stride(0)
    if cta_group is None:
        cta_group = GEMM_CTA_GROUP
    dp, dseed, dstream = dropout[:3] if dropout is not None else (0.0, 0, 0)
    dbits = dropout[3] if dropout is not None and len(dropout) > 3 else None
    dmul = dropout[4] if dropout is not None and len(dropout) > 4 else 1
    args = (int(a_mn), int(b_mn), M, N, K, _ptr(a), lda, _ptr(b), ldb, _ptr(bias),
            _ptr(residual), ld(residual), _ptr(gelu_u), ld(gelu_u), _ptr(out_f32), ld(out_f32),
            _ptr(out_bf16), ld(out_bf16), _ptr(out_pre), ld(out_pre), _ptr(colsum), int(apply_gelu),
            int(accumulate), float(alpha), int(k_splits), int(block_n), int(cta_group), float(dp), int(dseed),
            int(dstream), _ptr(dbits), int(dmul), _stream())
    if PROFILE.enabled:
        with PROFILE.region(2.0 * M * N * K):
            _lib.call("nv_gemm_bf16", *args)
        return
    _lib.call("nv_gemm_bf16", *args)


def head_fwd(x, ld_x, gamma, beta, W, bias, y, mean, rstd, logits, B, D, C, eps):
    _dev(x)
    _lib.call("nv_head_fwd", _ptr(x), ld_x, _ptr(gamma), _ptr(beta), _ptr(W), _ptr(bias), _ptr(y), _ptr(mean),
              _ptr(rstd), _ptr(logits), B, D, C, float(eps), _stream())


def head_bwd(dl, x, ld_x, y, mean, rstd, gamma, W, dx, ld_dx, dx_bf16, ld_dxb, dW, db, dgamma, dbeta, B, D, C):
    _dev(x)
    _lib.call("nv_head_bwd", _ptr(dl), _ptr(x), ld_x, _ptr(y), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(W), _ptr(dx),
              ld_dx, _ptr(dx_bf16), ld_dxb, _ptr(dW), _ptr(db), _ptr(dgamma), _ptr(dbeta), B, D, C, _stream())


def rng_epoch_advance():
    """Advance the device-side dropout epoch (end of a training step that may be replayed from a CUDA graph)."""
    _lib.call("nv_rng_epoch_advance", _stream())


def counter_add(counter, inc=1.0):
    _dev(counter)
    assert counter.dtype == F32 and counter.numel() == 1
    _lib.call("nv_counter_add", _ptr(counter), float(inc), _stream())


def adamw_flat(p, g, m, v, p_bf16, *, lr, beta1, beta2, eps, weight_decay, step, step_dev=None):
    """One fused AdamW step over flat fp32 buffers (+ bf16 copy of the new parameters). step_dev: device float
    holding the step count (read instead of `step`)."""
    _dev(p)
    n = p.numel()
    assert all(t.dtype == F32 and t.is_contiguous() and t.numel() == n for t in (p, g, m, v))
    assert p_bf16 is None or (p_bf16.dtype == BF16 and p_bf16.numel() == n and p_bf16.is_contiguous())
    _lib.call("nv_adamw_flat", _ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p_bf16), n, float(lr), float(beta1),
              float(beta2), float(eps), float(weight_decay), int(step), _ptr(step_dev), _stream())


def rng_epoch() -> int:
    """The device-side dropout epoch right now (synchronises the current stream): with (seed, epoch, site stream,
    element index) every keep bit is a closed form — oracle/rng_oracle.py."""
    out = ctypes.c_ulonglong(0)
    _lib.call("nv_rng_epoch_get", ctypes.byref(out), _stream())
    return int(out.value)


def dropout(x, *, p, seed, stream, residual=None, out_f32=None, out_bf16=None, colsum=None, row_mul=1):
    """v = x * keep / (1 - p) with the (seed, stream, row * N + col) mask of the GEMM epilogues; out = v (+ residual);
    colsum += column sums of v. x: fp32 [M, N] (unit inner stride), N % 8 == 0."""
    _dev(x)
    assert x.dtype == F32 and x.dim() == 2 and x.stride(1) == 1
    M, N = x.shape
    ld = lambda t: 0 if t is None else t.stride(0)
    for t, dt in ((residual, F32), (out_f32, F32), (out_bf16, BF16)):
        if t is not None:
            assert t.dtype == dt and tuple(t.shape) == (M, N) and t.stride(1) == 1
    _lib.call("nv_dropout", _ptr(x), ld(x), _ptr(residual), ld(residual), _ptr(out_f32), ld(out_f32), _ptr(out_bf16),
              ld(out_bf16), _ptr(colsum), M, N, float(p), int(seed), int(stream), int(row_mul), _stream())


def dropout_bits(out, *, p, seed, stream):
    """out (uint8 / int32 buffer, numel*itemsize % 4 == 0): byte g = keep bits of elements [8g, 8g+8) of the dropout
    site (seed, stream) — what the GEMM epilogues / LayerNorm backward / attention forward would draw inline."""
    _dev(out)
    assert out.is_contiguous()
    n_groups = out.numel() * out.element_size()
    _lib.call("nv_dropout_bits", _ptr(out), n_groups, float(p), int(seed), int(stream), _stream())


def dropout_flat(x, out, *, p, seed, stream):
    """Dropout over a contiguous fp32 tensor addressed by its flat element index (the fp32 verification mode's
    materialised attention probabilities, whose row length need not be a multiple of 8). The flat length is
    padded up to a multiple of 8 by viewing whole rows of 8; a ragged tail is handled by a second tiny call."""
    _dev(x)
    assert x.dtype == F32 and out.dtype == F32 and x.is_contiguous() and out.is_contiguous() and x.shape == out.shape
    L = x.numel()
    body = L // 8 * 8
    if body:
        dropout(x.view(-1)[:body].view(-1, 8), p=p, seed=seed, stream=stream, out_f32=out.view(-1)[:body].view(-1, 8))
    if L - body:  # < 8 trailing elements: own stream offset keeps the bits independent of the body's
        pad = torch.zeros(1, 8, device=x.device)
        pad[0, :L - body] = x.view(-1)[body:]
        res = torch.empty_like(pad)
        dropout(pad, p=p, seed=seed, stream=stream + 1000, out_f32=res)
        out.view(-1)[body:] = res[0, :L - body]


def dropout_keep_mask(M, N, *, p, seed, stream, device="cuda"):
    """The keep mask (1.0 / 0.0, fp32 [M, N]) of a dropout site — test / debugging helper."""
    ones = torch.ones(M, N, device=device)
    out = torch.empty_like(ones)
    dropout(ones, p=p, seed=seed, stream=stream, out_f32=out)
    return (out != 0).float()


def gemm_f32(M, N, K, a, sa, b, sb, c, sc, *, Z1=1, Z2=1, bias=None, residual=None, gelu_u=None, out_pre=None,
             ld_aux=0, apply_gelu=False, accumulate=False, alpha=1.0):
    """fp32 verification GEMM. sa=(m,k,z1,z2) sb=(n,k,z1,z2) sc=(m,z1,z2) element strides; a/b/c may be
    tensors or (tensor, element_offset) pairs. residual/gelu_u/out_pre share C's batch offsets and use
    row stride ld_aux."""
    def base(t):
        if isinstance(t, tuple):
            t, off = t
            _dev(t)
            assert t.dtype == F32
            return ctypes.c_void_p(t.data_ptr() + 4 * off)
        _dev(t)
        assert t.dtype == F32
        return _ptr(t)

    _lib.call("nv_gemm_f32", M, N, K, Z1, Z2, base(a), sa[0], sa[1], sa[2], sa[3], base(b), sb[0], sb[1], sb[2],
              sb[3], base(c), sc[0], sc[1], sc[2], _ptr(bias), _ptr(residual), ld_aux, _ptr(gelu_u), ld_aux,
              _ptr(out_pre), ld_aux, int(apply_gelu), int(accumulate), float(alpha), _stream())


def linear_f32(x, w, *, bias=None, residual=None, gelu_u=None, out=None, out_pre=None, apply_gelu=False,
               accumulate=False, alpha=1.0, w_kn=False, x_km=False):
    """fp32 helper on top of gemm_f32: out[M,N] = x[M,K] @ w[N,K]^T (w_kn: w is [K,N]; x_km: x is [K,M])."""
    (K, M) = x.shape if x_km else (x.shape[1], x.shape[0])
    (Kw, N) = w.shape if w_kn else (w.shape[1], w.shape[0])
    assert K == Kw, f"linear_f32: {K} vs {Kw}"
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=F32)
    sa = (x.stride(1), x.stride(0), 0, 0) if x_km else (x.stride(0), x.stride(1), 0, 0)
    sb = (w.stride(1), w.stride(0), 0, 0) if w_kn else (w.stride(0), w.stride(1), 0, 0)
    assert out.stride(1) == 1
    for t in (residual, gelu_u, out_pre):
        if t is not None:
            assert t.stride(1) == 1 and t.stride(0) == out.stride(0) and t.dtype == F32
    gemm_f32(M, N, K, x, sa, w, sb, out, (out.stride(0), 0, 0), bias=bias, residual=residual, gelu_u=gelu_u,
             out_pre=out_pre, ld_aux=out.stride(0), apply_gelu=apply_gelu, accumulate=accumulate, alpha=alpha)
    return out


# --------------------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x, gamma, beta, y, *, M, D, ld_x=None, ld_y=None, xmap=(0, 0, 0), ymap=(0, 0, 0), add=None,
                  ld_add=0, add_mod=1, add_off=0, mean=None, rstd=None, eps=1e-5):
    _dev(x)
    assert x.dtype == F32 and y.dtype in (F32, BF16)
    _lib.call("nv_layernorm_fwd", _ptr(x), D if ld_x is None else ld_x, *xmap, _ptr(gamma), _ptr(beta), _ptr(add),
              ld_add, add_mod, add_off, _ptr(y), int(y.dtype == BF16), D if ld_y is None else ld_y, *ymap,
              _ptr(mean), _ptr(rstd), M, D, float(eps), _stream())


def layernorm_bwd(dy, x, mean, rstd, gamma, *, M, D, ld_dy=None, ld_x=None, dymap=(0, 0, 0), xmap=(0, 0, 0),
                  dres=None, ld_dres=None, dx=None, ld_dx=None, dxmap=(0, 0, 0), dx_bf16=None, ld_dxb=None,
                  dgamma=None, dbeta=None, colsum=None, side_drop=None):
    """side_drop = (p, seed, stream): mask dx_bf16 / colsum (not dx) with that dropout site's forward mask."""
    _dev(x)
    assert dy.dtype in (F32, BF16) and x.dtype == F32
    sp, sseed, sstream = side_drop[:3] if side_drop is not None else (0.0, 0, 0)
    sbits = side_drop[3] if side_drop is not None and len(side_drop) > 3 else None
    _lib.call("nv_layernorm_bwd", _ptr(dy), int(dy.dtype == BF16), D if ld_dy is None else ld_dy, *dymap, _ptr(x),
              D if ld_x is None else ld_x, *xmap, _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dres),
              D if ld_dres is None else ld_dres, _ptr(dx), D if ld_dx is None else ld_dx, *dxmap, _ptr(dx_bf16),
              D if ld_dxb is None else ld_dxb, _ptr(dgamma), _ptr(dbeta), _ptr(colsum), M, D, float(sp), int(sseed),
              int(sstream), _ptr(sbits), _stream())


def cls_row(cls, pos, x, batch_stride, B, D):
    _dev(x)
    _lib.call("nv_cls_row", _ptr(cls), _ptr(pos), _ptr(x), batch_stride, B, D, _stream())


# ------------------------------------------------------------------------------------ patch gather
def _i64x(vals):
    return (ctypes.c_int64 * len(vals))(*[int(v) for v in vals])


def patch_gather_ln(video, patch, gamma, beta, out, *, raw=None, mean=None, rstd=None, eps=1e-5):
    """video: 5-D fp32 view [B,C,F,H,W] (any strides); patch=(pf,p1,p2); out [B*n, ld] bf16/fp32 or None."""
    _dev(video)
    assert video.dtype == F32 and video.dim() == 5
    dims, strides, p = _i64x(video.shape), _i64x(video.stride()), _i64x(patch)
    _lib.call("nv_patch_gather_ln", _ptr(video), dims, strides, p, _ptr(gamma), _ptr(beta), _ptr(out),
              int(out is not None and out.dtype == BF16), 0 if out is None else out.stride(0), _ptr(raw),
              _ptr(mean), _ptr(rstd), float(eps), _stream())


def ln_fold(W, gamma, beta, b, Wf, bias_f):
    """Fold LayerNorm's affine into the Linear behind it: Wf [D, ld] (bf16 / fp32, pad columns zeroed) = W o gamma,
    bias_f [D] = b + W beta. W [D, P] fp32 contiguous."""
    _dev(W)
    D, P = W.shape
    assert W.dtype == F32 and W.is_contiguous() and Wf.shape[0] == D and Wf.stride(1) == 1 and Wf.shape[1] >= P
    assert Wf.stride(0) == Wf.shape[1], "ln_fold zeroes columns P .. row pitch: the pitch must be the padded width"
    _lib.call("nv_ln_fold", _ptr(W), _ptr(gamma), _ptr(beta), _ptr(b), _ptr(Wf), int(Wf.dtype == BF16), Wf.stride(0),
              _ptr(bias_f), D, P, _stream())


def ln_fold_grads(G, W, gamma, beta, cs, dW, dgamma, dbeta, db=None):
    """Parameter gradients of LayerNorm -> Linear from G = de^T xhat [D, >= P] and cs = colsum(de) [D]; all outputs +=."""
    _dev(G)
    D, P = W.shape
    assert G.dtype == F32 and G.stride(1) == 1 and G.shape[0] == D and G.shape[1] >= P
    assert W.dtype == F32 and W.is_contiguous() and dW.dtype == F32 and dW.is_contiguous() and tuple(dW.shape) == (D, P)
    _lib.call("nv_ln_fold_grads", _ptr(G), G.stride(0), _ptr(W), _ptr(gamma), _ptr(beta), _ptr(cs), _ptr(dW),
              _ptr(dgamma), _ptr(dbeta), _ptr(db), D, P, _stream())


def patch_ln_param_grad(video, patch, dP, mean, rstd, dgamma, dbeta):
    _dev(video)
    assert dP.dtype == F32 and dP.stride(1) == 1
    dims, strides, p = _i64x(video.shape), _i64x(video.stride()), _i64x(patch)
    _lib.call("nv_patch_ln_param_grad", _ptr(video), dims, strides, p, _ptr(dP), dP.stride(0), _ptr(mean),
              _ptr(rstd), _ptr(dgamma), _ptr(dbeta), _stream())


# --------------------------------------------------------------------------------------- attention
def _off(t, elems):
    return ctypes.c_void_p(t.data_ptr() + elems * t.element_size())


def attention_fwd(qkv, o, lse, *, B, N, H, head_dim, scale, dropout_p=0.0, seed=0, drop_mask=None, mask_ready=False):
    """qkv [B*N, 3*H*hd] bf16 (q|k|v column blocks), o [B*N, H*hd] bf16, lse [B,H,N] fp32;
    drop_mask uint32-as-int32 [B*H, N, ceil(N/32)] when dropout_p > 0 (mask_ready: already drawn by dropout_bits)."""
    _dev(qkv)
    assert qkv.dtype == BF16 and o.dtype == BF16 and lse.dtype == F32
    inner = H * head_dim
    rs = qkv.stride(0)
    _lib.call("nv_attention_fwd", _off(qkv, 0), _off(qkv, inner), _off(qkv, 2 * inner), N * rs, rs, _ptr(o),
              N * o.stride(0), o.stride(0), _ptr(lse), B, N, H, head_dim, float(scale), float(dropout_p), int(seed),
              _ptr(drop_mask), int(bool(mask_ready)), _stream())


def attention_cls_fwd(qkv, o_cls, lse, *, B, N, H, head_dim, scale, dropout_p=0.0, seed=0, drop_mask=None,
                      mask_ready=False):
    """Attention forward for query token 0 only (last block under a cls-pooled head): o_cls [B, H*hd] bf16; lse [B,H,N]
    and drop_mask [B*H, N, ceil(N/32)] keep the full layouts, only their token-0 entries / rows are written."""
    _dev(qkv)
    assert qkv.dtype == BF16 and o_cls.dtype == BF16 and lse.dtype == F32
    inner = H * head_dim
    assert tuple(o_cls.shape) == (B, inner) and o_cls.stride(1) == 1 and tuple(lse.shape) == (B, H, N) and lse.is_contiguous()
    rs = qkv.stride(0)
    _lib.call("nv_attention_cls_fwd", _off(qkv, 0), _off(qkv, inner), _off(qkv, 2 * inner), N * rs, rs, _ptr(o_cls),
              o_cls.stride(0), _ptr(lse), B, N, H, head_dim, float(scale), float(dropout_p), int(seed), _ptr(drop_mask),
              int(bool(mask_ready)), _stream())


def attention_cls_bwd(qkv, o, dO_cls, lse, dqkv, *, B, N, H, head_dim, scale, dropout_p=0.0, drop_mask=None, o_bs=None):
    """Attention backward when only token 0 of every sample has gradient: dO_cls [B, H*hd] bf16 (contiguous rows).
    o: the forward's output, [B*N, H*hd] (token-0 rows read at stride N) or, with o_bs given, any tensor whose sample b's
    token-0 row starts at element b * o_bs (the compact [B, H*hd] of attention_cls_fwd: o_bs = its row stride)."""
    _dev(qkv)
    inner = H * head_dim
    rs, drs = qkv.stride(0), dqkv.stride(0)
    assert dO_cls.dtype == BF16 and dO_cls.stride(1) == 1 and tuple(dO_cls.shape) == (B, inner)
    _lib.call("nv_attention_cls_bwd", _off(qkv, 0), _off(qkv, inner), _off(qkv, 2 * inner), N * rs, rs, _ptr(o),
              N * o.stride(0) if o_bs is None else int(o_bs), _ptr(dO_cls), dO_cls.stride(0), _ptr(lse), _off(dqkv, 0), _off(dqkv, inner),
              _off(dqkv, 2 * inner), N * drs, drs, B, N, H, head_dim, float(scale), float(dropout_p), _ptr(drop_mask),
              _stream())


def attention_bwd(qkv, o, dO, lse, delta_ws, dqkv, *, B, N, H, head_dim, scale, dropout_p=0.0, drop_mask=None):
    _dev(qkv)
    inner = H * head_dim
    rs, drs = qkv.stride(0), dqkv.stride(0)
    assert o.stride(0) == dO.stride(0)
    _lib.call("nv_attention_bwd", _off(qkv, 0), _off(qkv, inner), _off(qkv, 2 * inner), N * rs, rs, _ptr(o),
              _ptr(dO), N * o.stride(0), o.stride(0), _ptr(lse), _ptr(delta_ws), _off(dqkv, 0), _off(dqkv, inner),
              _off(dqkv, 2 * inner), N * drs, drs, B, N, H, head_dim, float(scale), float(dropout_p),
              _ptr(drop_mask), _stream())


def softmax_fwd(s, rows, n):
    _dev(s)
    _lib.call("nv_softmax_fwd", _ptr(s), rows, n, _stream())


def softmax_bwd(P, dP, rows, n):
    _dev(P)
    _lib.call("nv_softmax_bwd", _ptr(P), _ptr(dP), rows, n, _stream())


# ----------------------------------------------------------------------------------------- helpers
def cast_bf16(x, out=None):
    _dev(x)
    assert x.dtype == F32 and x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=BF16)
    _lib.call("nv_cast_f32_bf16", _ptr(x), _ptr(out), x.numel(), _stream())
    return out


def cast_transpose_bf16(w, out=None, outT=None):
    """w [R,C] fp32 -> (bf16 [R,C] or None, bf16 [C,R])."""
    _dev(w)
    assert w.dtype == F32 and w.dim() == 2 and w.is_contiguous()
    R, C = w.shape
    if outT is None:
        outT = torch.empty(C, R, device=w.device, dtype=BF16)
    _lib.call("nv_cast_transpose_f32_bf16", _ptr(w), _ptr(out), _ptr(outT), R, C, _stream())
    return out, outT


def colsum(x, out, *, M=None, N=None):
    _dev(x)
    assert x.dim() == 2 and x.stride(1) == 1 and out.dtype == F32
    M = x.shape[0] if M is None else M
    N = x.shape[1] if N is None else N
    _lib.call("nv_colsum", _ptr(x), int(x.dtype == BF16), x.stride(0), _ptr(out), M, N, _stream())


def batch_sum(x, batch_stride, out, B, L):
    _dev(x)
    _lib.call("nv_batch_sum", _ptr(x), batch_stride, _ptr(out), B, L, _stream())


def mean_pool_fwd(x, pooled, B, N, D):
    _dev(x)
    _lib.call("nv_mean_pool_fwd", _ptr(x), _ptr(pooled), B, N, D, _stream())


def mean_pool_bwd(dpooled, dx, dx_bf16, B, N, D):
    _dev(dx)
    _lib.call("nv_mean_pool_bwd", _ptr(dpooled), _ptr(dx), _ptr(dx_bf16), B, N, D, _stream())


def set_sm_reserve(n: int):
    """Leave n SMs (rounded up to even) free of the persistent kernels on the current device (for NCCL's CTAs)."""
    _lib.call("nv_set_sm_reserve", int(n))


def fmri_deinterleave(x, y, B, S, T, stats_ws=None, eps=1e-8):
    """x [B, S, T] fp32 contiguous -> y [B, T, S]; stats_ws (2*B float64, device) turns on the per-sample z-score."""
    _dev(x)
    _lib.call("nv_fmri_deinterleave", _ptr(x), _ptr(y), B, S, T, _ptr(stats_ws), float(eps), _stream())


def temporal_fwd(x, params, out, saved, B, T, F, eps=1e-5, seq_out=None, drop=(0.0, 0.0, 0.0, 0.0), seed=0):
    """drop = (p_attn, p_dropout1, p_ffn, p_dropout2) of nn.TransformerEncoderLayer in training mode."""
    _dev(x)
    _lib.call("nv_temporal_fwd", _ptr(x), _ptr(params), _ptr(out), _ptr(seq_out), _ptr(saved), B, T, F, float(eps),
              *[float(p) for p in drop], int(seed), _stream())


def temporal_bwd(x, params, saved, dout, dparams_ws, dx, B, T, F, eps=1e-5, dseq=None, drop=(0.0, 0.0, 0.0, 0.0),
                 seed=0):
    _dev(x)
    _lib.call("nv_temporal_bwd", _ptr(x), _ptr(params), _ptr(saved), _ptr(dout), _ptr(dseq), _ptr(dparams_ws),
              _ptr(dx), B, T, F, float(eps), *[float(p) for p in drop], int(seed), _stream())
