"""Per-kernel parity on the GPU, each C-ABI kernel against the matching torch sub-graph (fp32/fp64 on the
same device). Tolerances: bf16 kernels 2e-2 relative (north_star), fp32 kernels 1e-5, index work bit-exact."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # collected on CPU boxes too; every test here is gpu-marked
    pytest.skip("CUDA device required", allow_module_level=True)

from einops import rearrange  # noqa: E402

from neurovit_b200 import ops  # noqa: E402

DEV = "cuda"


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


# ------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("shape", [(256, 256, 128), (392, 1536, 1024), (136, 72, 200), (385, 264, 520)])
@pytest.mark.parametrize("cg", [1, 2])
def test_gemm_bf16_plain(a_mn, b_mn, block_n, shape, cg):
    M, N, K = shape
    torch.manual_seed(1)
    a = torch.randn((K, M) if a_mn else (M, K), device=DEV).to(torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=DEV).to(torch.bfloat16)
    if (a_mn and M % 8) or (b_mn and N % 8) or ((not a_mn or not b_mn) and K % 8):
        pytest.skip("TMA needs 16-byte row strides")
    out = torch.empty(M, N, device=DEV)
    ops.gemm_bf16(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out_f32=out, block_n=block_n, cta_group=cg)
    af = a.double().t() if a_mn else a.double()
    bf = b.double().t() if b_mn else b.double()
    assert rel_err(out, af @ bf.t()) < 1e-3


def test_gemm_bf16_epilogues():
    torch.manual_seed(2)
    M, N, K = 770, 1024, 512
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    ref = a.double() @ w.double().t() + bias.double()
    # bias + residual, fp32 out + bf16 copy
    out = torch.empty(M, N, device=DEV)
    outb = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm_bf16(a, w, bias=bias, residual=res, out_f32=out, out_bf16=outb)
    assert rel_err(out, ref + res.double()) < 1e-3
    assert rel_err(outb, ref + res.double()) < 1e-2
    # bias + GELU: pre-activation and activation in bf16
    pre = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    act = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm_bf16(a, w, bias=bias, out_pre=pre, out_bf16=act, apply_gelu=True)
    assert rel_err(pre, ref) < 1e-2
    assert rel_err(act, torch.nn.functional.gelu(ref)) < 1e-2
    # dgrad through GELU: acc * gelu'(u)
    u = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    ud = u.double().requires_grad_(True)
    g, = torch.autograd.grad(torch.nn.functional.gelu(ud).sum(), ud)
    cs = torch.ones(N, device=DEV)
    wt = w.t().contiguous()  # dgrad layout: B stored [K, N]
    ops.gemm_bf16(a, wt, b_mn=True, gelu_u=u, out_f32=out, colsum=cs)
    assert rel_err(out, (a.double() @ w.double().t()) * g) < 1e-3
    assert rel_err(cs, ((a.double() @ w.double().t()) * g).sum(0) + 1) < 1e-3   # fused bias-gradient column sum
    # split-K accumulate into a pre-loaded fp32 buffer (wgrad): dW[N_out,K_in] = dY^T X
    dy = torch.randn(M, 256, device=DEV).to(torch.bfloat16)
    x = torch.randn(M, 384, device=DEV).to(torch.bfloat16)
    dw = torch.full((256, 384), 0.5, device=DEV)
    ops.gemm_bf16(dy, x, a_mn=True, b_mn=True, out_f32=dw, accumulate=True, k_splits=5)
    assert rel_err(dw, dy.double().t() @ x.double() + 0.5) < 1e-3


@pytest.mark.parametrize("shape", [(385, 264, 520), (770, 1024, 1024), (130, 72, 64), (1000, 520, 136)])
@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("inplace", [False, True])
def test_gemm_residual_ring_ragged(shape, block_n, inplace):
    """Forward linear + fp32 residual on CTA pairs with K <= 1024 takes the cp.async residual ring (store mode 3):
    ragged last row tile, a last column tile that is partly (or, for whole 32-column chunks, entirely) outside N, and
    the in-place form x = x + linear(a) the transformer blocks use (vit_3d.py:60,73 `x = attn(x) + x`)."""
    M, N, K = shape
    torch.manual_seed(5)
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    out = res if inplace else torch.full((M, N), float("nan"), device=DEV)
    cs = torch.zeros(N, device=DEV)
    ops.gemm_bf16(a, w, bias=bias, residual=res, out_f32=out, colsum=cs, block_n=block_n, cta_group=2)
    assert rel_err(out, ref) < 1e-3
    assert rel_err(cs, ref.sum(0)) < 1e-3


def test_gemm_f32_matches_torch():
    torch.manual_seed(3)
    M, N, K = 200, 136, 300
    x = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    out = ops.linear_f32(x, w, bias=bias, residual=res)
    assert rel_err(out, x.double() @ w.double().t() + bias.double() + res.double()) < 1e-5
    out2 = ops.linear_f32(x.t().contiguous(), w.t().contiguous(), x_km=True, w_kn=True)
    assert rel_err(out2, x.double() @ w.double().t()) < 1e-5
    pre = torch.empty(M, N, device=DEV)
    out3 = ops.linear_f32(x, w, bias=bias, out_pre=pre, apply_gelu=True)
    assert rel_err(pre, x.double() @ w.double().t() + bias.double()) < 1e-5
    assert rel_err(out3, torch.nn.functional.gelu(x.double() @ w.double().t() + bias.double())) < 1e-5


# -------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("D", [64, 512, 1024])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd(D, out_dtype):
    torch.manual_seed(4)
    M = 777
    x = (torch.randn(M, D, device=DEV) * 2 + 0.5).requires_grad_(True)
    g = (torch.randn(D, device=DEV) * 0.5 + 1).requires_grad_(True)
    b = torch.randn(D, device=DEV).requires_grad_(True)
    y = torch.empty(M, D, device=DEV, dtype=out_dtype)
    mean = torch.empty(M, device=DEV)
    rstd = torch.empty(M, device=DEV)
    ops.layernorm_fwd(x.detach(), g.detach(), b.detach(), y, M=M, D=D, mean=mean, rstd=rstd)
    ref = torch.nn.functional.layer_norm(x.double(), (D,), g.double(), b.double(), 1e-5)
    assert rel_err(y, ref) < (1e-5 if out_dtype == torch.float32 else 1e-2)
    dy = torch.randn(M, D, device=DEV)
    dres = torch.randn(M, D, device=DEV)
    gx, gg, gb = torch.autograd.grad(ref, (x, g, b), dy.double(), retain_graph=True)
    dx = torch.empty(M, D, device=DEV)
    dxb = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    dg = torch.zeros(D, device=DEV)
    db = torch.zeros(D, device=DEV)
    cs = torch.zeros(D, device=DEV)
    ops.layernorm_bwd(dy, x.detach(), mean, rstd, g.detach(), M=M, D=D, dres=dres, dx=dx, dx_bf16=dxb, dgamma=dg,
                      dbeta=db, colsum=cs)
    assert rel_err(dx, gx.double() + dres.double()) < 1e-5
    assert rel_err(dxb, gx.double() + dres.double()) < 1e-2
    assert rel_err(dg, gg) < 1e-4
    assert rel_err(db, gb) < 1e-4
    assert rel_err(cs, (gx.double() + dres.double()).sum(0)) < 1e-4
    # bf16 dy (the dgrad GEMM's output in bf16 mode): same math on the rounded input
    dyb = dy.to(torch.bfloat16)
    gx2, gg2, gb2 = torch.autograd.grad(ref, (x, g, b), dyb.double())
    dg.zero_(); db.zero_(); cs.zero_()
    ops.layernorm_bwd(dyb, x.detach(), mean, rstd, g.detach(), M=M, D=D, dres=dres, dx=dx, dx_bf16=dxb, dgamma=dg,
                      dbeta=db, colsum=cs)
    assert rel_err(dx, gx2.double() + dres.double()) < 1e-5
    assert rel_err(dg, gg2) < 1e-4 and rel_err(db, gb2) < 1e-4


def test_layernorm_row_maps_and_pos_add():
    """LN over the patch rows written into x[B, n+1, D] at token offset 1 with the pos-embedding add."""
    torch.manual_seed(5)
    B, n, D = 3, 10, 128
    e = torch.randn(B * n, D, device=DEV)
    g = torch.randn(D, device=DEV)
    b = torch.randn(D, device=DEV)
    pos = torch.randn(n + 1, D, device=DEV)
    cls = torch.randn(D, device=DEV)
    x = torch.zeros(B, n + 1, D, device=DEV)
    ops.layernorm_fwd(e, g, b, x, M=B * n, D=D, ymap=(n, n + 1, 1), add=pos, ld_add=D, add_mod=n, add_off=1)
    ops.cls_row(cls, pos, x, (n + 1) * D, B, D)
    ref = torch.nn.functional.layer_norm(e, (D,), g, b).view(B, n, D) + pos[1:]
    ref = torch.cat([(cls + pos[0]).expand(B, 1, D), ref], 1)
    assert rel_err(x, ref) < 1e-5


# ----------------------------------------------------------------------------------- patch gather
@pytest.mark.parametrize("layout", ["neuro_view", "contiguous"])
@pytest.mark.parametrize("geom", [(2, 1, 16, 16, 8, 8), (2, 1, 18, 18, 18, 9), (1, 2, 8, 12, 4, 4),
                                  (5, 1, 64, 64, 48, 8)])
def test_patch_gather_bit_exact(layout, geom):
    """(the ViT3DEncoder view with patch 8 takes the 5-D TMA box kernels, everything else the strided-load ones)"""
    B, C, H, W, D_, p = geom
    torch.manual_seed(6)
    if layout == "neuro_view":
        if C != 1:
            pytest.skip("ViT3DEncoder view has one channel")
        x = torch.randn(B, H, W, D_, device=DEV)
        video = x.permute(0, 3, 1, 2).unsqueeze(1)  # [B,1,D,H,W] view, NeuroEncoder.py:201-202
    else:
        video = torch.randn(B, C, D_, H, W, device=DEV)
    ref = rearrange(video, 'b c (f pf) (h p1) (w p2) -> b (f h w) (p1 p2 pf c)', p1=p, p2=p, pf=p)
    n, P = ref.shape[1], ref.shape[2]
    raw = torch.empty(B * n, P, device=DEV)
    ld = (P + 7) // 8 * 8
    out = torch.empty(B * n, ld, device=DEV)
    g = torch.randn(P, device=DEV)
    b = torch.randn(P, device=DEV)
    mean = torch.empty(B * n, device=DEV)
    rstd = torch.empty(B * n, device=DEV)
    ops.patch_gather_ln(video, (p, p, p), g, b, out, raw=raw, mean=mean, rstd=rstd)
    assert torch.equal(raw.view(B, n, P), ref), "patch index mapping must be bit-exact"
    lnref = torch.nn.functional.layer_norm(ref.double(), (P,), g.double(), b.double())
    assert rel_err(out[:, :P], lnref.view(B * n, P)) < 1e-5
    assert (out[:, P:] == 0).all()
    # parameter gradients of that LayerNorm
    dP = torch.randn(B * n, ld, device=DEV)
    dg = torch.zeros(P, device=DEV)
    db = torch.zeros(P, device=DEV)
    ops.patch_ln_param_grad(video, (p, p, p), dP, mean, rstd, dg, db)
    xh = torch.nn.functional.layer_norm(ref.double(), (P,)).view(B * n, P)
    assert rel_err(dg, (dP[:, :P].double() * xh).sum(0)) < 1e-4
    assert rel_err(db, dP[:, :P].double().sum(0)) < 1e-4


def test_patch_gather_tma_and_strided_kernels_agree_bitwise():
    import os
    torch.manual_seed(16)
    B, H, W, D_, p = 3, 32, 24, 16, 8
    video = torch.randn(B, H, W, D_, device=DEV).permute(0, 3, 1, 2).unsqueeze(1)
    n, P = (H // p) * (W // p) * (D_ // p), p ** 3
    g, b = torch.randn(P, device=DEV), torch.randn(P, device=DEV)
    dP = torch.randn(B * n, P, device=DEV)
    res = []
    for no_tma in ("", "1"):
        if no_tma:
            os.environ["NV_PATCH_NO_TMA"] = "1"
        try:
            out = torch.empty(B * n, P, device=DEV, dtype=torch.bfloat16)
            raw = torch.empty(B * n, P, device=DEV)
            mean, rstd = torch.empty(B * n, device=DEV), torch.empty(B * n, device=DEV)
            ops.patch_gather_ln(video, (p, p, p), g, b, out, raw=raw, mean=mean, rstd=rstd)
            dg, db = torch.zeros(P, device=DEV), torch.zeros(P, device=DEV)
            ops.patch_ln_param_grad(video, (p, p, p), dP, mean, rstd, dg, db)
            res.append((out.clone(), raw.clone(), mean.clone(), rstd.clone(), dg, db))
        finally:
            os.environ.pop("NV_PATCH_NO_TMA", None)
    for a, c in zip(res[0][:4], res[1][:4]):
        assert torch.equal(a, c)
    assert rel_err(res[0][4], res[1][4]) < 1e-5 and rel_err(res[0][5], res[1][5]) < 1e-5  # atomics: order differs


# -------------------------------------------------------------------------------------- attention
def _attn_ref(qkv, B, N, H, hd):
    q, k, v = qkv.view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    p = s.softmax(-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * N, H * hd)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,N,H", [(2, 385, 8), (1, 64, 2), (3, 9, 2), (1, 1729, 1), (2, 1001, 2), (2, 1, 2), (1, 2, 1),
                                   (1, 129, 2), (2, 130, 1), (1, 97, 1), (1, 98, 2)])
def test_attention_fwd_bwd(B, N, H):
    hd = 64
    torch.manual_seed(7)
    qkv = torch.randn(B * N, 3 * H * hd, device=DEV).to(torch.bfloat16)
    o = torch.empty(B * N, H * hd, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=DEV)
    ops.attention_fwd(qkv, o, lse, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5)
    qd = qkv.double().requires_grad_(True)
    oref, lref = _attn_ref(qd, B, N, H, hd)
    assert rel_err(o, oref) < 1e-2
    assert rel_err(lse, lref) < 1e-3
    dO = torch.randn(B * N, H * hd, device=DEV).to(torch.bfloat16)
    gref, = torch.autograd.grad(oref, qd, dO.double())
    dqkv = torch.zeros_like(qkv)
    ws = torch.empty(B * H * N, device=DEV)
    ops.attention_bwd(qkv, o, dO, lse, ws, dqkv, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5)
    inner = H * hd
    # N = 1: softmax over one key has dS = 0 exactly, so dq = dk = 0 in exact arithmetic while the kernel leaves the
    # bf16 round-off of p (dP - delta); errors are therefore measured against the scale of the whole gradient
    floor = 1e-3 * gref.abs().max().item()

    def err(a, b):
        return ((a.double() - b.double()).abs().max() / max(b.abs().max().item(), floor)).item()

    for nm, sl in (("dq", slice(0, inner)), ("dk", slice(inner, 2 * inner)), ("dv", slice(2 * inner, 3 * inner))):
        assert err(dqkv[:, sl], gref[:, sl]) < 2e-2, nm
    # token 0 (handled outside the tensor-core tiles, as a query and as a key) and the last token, on their own
    for tok in (0, N - 1):
        assert rel_err(o.view(B, N, inner)[:, tok], oref.view(B, N, inner)[:, tok]) < 1e-2, tok
        assert err(dqkv.view(B, N, 3 * inner)[:, tok], gref.view(B, N, 3 * inner)[:, tok]) < 2e-2, tok


@pytest.mark.parametrize("B,N,H", [(3, 385, 8), (2, 64, 2), (2, 9, 1), (1, 1729, 2), (2, 1, 1)])
@pytest.mark.parametrize("p_drop", [0.0, 0.25])
def test_attention_cls_bwd_matches_full_backward(B, N, H, p_drop):
    """nv_attention_cls_bwd (only token 0 of every sample has gradient: the last block under pool='cls',
    vit_3d.py:123) against nv_attention_bwd fed the same dO with every other row zero, and against autograd."""
    hd = 64
    inner = H * hd
    torch.manual_seed(21)
    qkv = torch.randn(B * N, 3 * inner, device=DEV).to(torch.bfloat16)
    o = torch.empty(B * N, inner, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=DEV)
    mask = torch.zeros(B * H, N, (N + 31) // 32, device=DEV, dtype=torch.int32) if p_drop > 0 else None
    ops.attention_fwd(qkv, o, lse, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=p_drop, seed=77, drop_mask=mask)
    dO_cls = torch.randn(B, inner, device=DEV).to(torch.bfloat16)
    dO = torch.zeros(B * N, inner, device=DEV, dtype=torch.bfloat16)
    dO.view(B, N, inner)[:, 0] = dO_cls
    full = torch.zeros_like(qkv)
    ws = torch.empty(B * H * N, device=DEV)
    ops.attention_bwd(qkv, o, dO, lse, ws, full, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=p_drop,
                      drop_mask=mask)
    got = torch.full_like(qkv, float("nan"))   # the kernel must write every element, zeros included
    ops.attention_cls_bwd(qkv, o, dO_cls, lse, got, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=p_drop,
                          drop_mask=mask)
    assert torch.isfinite(got.float()).all()
    for nm, sl in (("dq", slice(0, inner)), ("dk", slice(inner, 2 * inner)), ("dv", slice(2 * inner, 3 * inner))):
        assert rel_err(got[:, sl], full[:, sl]) < 1e-2, nm
    assert (got.view(B, N, 3 * inner)[:, 1:, :inner] == 0).all()      # dQ rows of the other tokens
    if p_drop == 0.0:
        qd = qkv.double().requires_grad_(True)
        oref, _ = _attn_ref(qd, B, N, H, hd)
        gref, = torch.autograd.grad(oref, qd, dO.double())
        assert rel_err(got, gref) < 2e-2


@pytest.mark.parametrize("B,N,H", [(3, 385, 8), (5, 64, 2), (2, 9, 1), (1, 1729, 2), (2, 1, 1), (7, 129, 3)])
@pytest.mark.parametrize("p_drop,predrawn", [(0.0, False), (0.25, False), (0.25, True)])
def test_attention_cls_fwd_equals_token0_of_full_forward(B, N, H, p_drop, predrawn):
    """nv_attention_cls_fwd (query token 0 only: the last block under pool='cls', vit_3d.py:123) is the full forward's
    own token-0 code run alone: output row, lse entry and the mask row it draws are bit-identical, and it feeds
    nv_attention_cls_bwd (compact o, batch stride given) to the same gradients."""
    hd = 64
    inner = H * hd
    torch.manual_seed(23)
    qkv = torch.randn(B * N, 3 * inner, device=DEV).to(torch.bfloat16)
    kw = dict(B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=p_drop)
    words = (N + 31) // 32

    def new_mask():
        if p_drop == 0:
            return None
        m = torch.zeros(B * H, N, words, device=DEV, dtype=torch.int32)
        if predrawn:
            ops.dropout_bits(m, p=p_drop, seed=99, stream=0)
        return m

    o = torch.empty(B * N, inner, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=DEV)
    mask = new_mask()
    ops.attention_fwd(qkv, o, lse, seed=99, drop_mask=mask, mask_ready=predrawn, **kw)
    o_cls = torch.full((B, inner), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse_c = torch.full((B, H, N), float("nan"), device=DEV)
    mask_c = new_mask()
    ops.attention_cls_fwd(qkv, o_cls, lse_c, seed=99, drop_mask=mask_c, mask_ready=predrawn, **kw)
    assert torch.equal(o_cls, o.view(B, N, inner)[:, 0])
    assert torch.equal(lse_c[:, :, 0], lse[:, :, 0])
    if mask is not None:
        assert torch.equal(mask_c[:, 0], mask[:, 0])                       # the cls query's keep bits, drawn or read
        if not predrawn and N > 1:
            assert (mask_c[:, 1:] == 0).all()                              # nothing else is touched
    dO_cls = torch.randn(B, inner, device=DEV).to(torch.bfloat16)
    want = torch.full_like(qkv, float("nan"))
    ops.attention_cls_bwd(qkv, o, dO_cls, lse, want, drop_mask=mask, **kw)
    got = torch.full_like(qkv, float("nan"))
    ops.attention_cls_bwd(qkv, o_cls, dO_cls, lse_c, got, drop_mask=mask_c, o_bs=o_cls.stride(0), **kw)
    assert torch.equal(got, want)


def _mask_bits_to_keys(mask, B, H, N):
    """Saved keep-bit words [B*H, N, ceil(N/32)] -> float keep mask [B, H, N(query), N(key)]. Bit position p of a row
    is key token p + 1 for p < N - 1 and key token 0 for p = N - 1 (csrc/attention_tc.cu: token 0 is handled outside
    the tiles, so the tiled keys 1..N-1 keep word-aligned positions)."""
    w = mask.view(B, H, N, -1).to(torch.int64) & 0xFFFFFFFF
    bits = ((w.unsqueeze(-1) >> torch.arange(32, device=w.device)) & 1).reshape(B, H, N, -1)[..., :N]
    return torch.cat([bits[..., N - 1:N], bits[..., :N - 1]], dim=-1).double()


@pytest.mark.parametrize("B,N,H", [(2, 385, 4), (1, 1001, 2), (2, 64, 2), (3, 9, 1), (2, 1, 1), (1, 161, 2)])
@pytest.mark.parametrize("predrawn", [False, True])
def test_attention_dropout_fwd_bwd_matches_masked_reference(B, N, H, predrawn):
    """Attention dropout (vit_3d.py:56): the keep bits the kernels drew (inline, or ahead of time by nv_dropout_bits —
    identical bits) are replayed into an fp64 torch reference; forward and all three gradients must agree."""
    hd, p_drop, seed = 64, 0.3, 1234
    inner = H * hd
    torch.manual_seed(5)
    qkv = torch.randn(B * N, 3 * inner, device=DEV).to(torch.bfloat16)
    o = torch.empty(B * N, inner, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=DEV)
    mw = (N + 31) // 32
    mask = torch.zeros(B * H, N, mw, device=DEV, dtype=torch.int32)
    if predrawn:
        ops.dropout_bits(mask, p=p_drop, seed=seed, stream=0)
    ops.attention_fwd(qkv, o, lse, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=p_drop, seed=seed,
                      drop_mask=mask, mask_ready=predrawn)
    inline = torch.zeros_like(mask)
    ops.dropout_bits(inline, p=p_drop, seed=seed, stream=0)
    keep = _mask_bits_to_keys(mask, B, H, N)
    assert torch.equal(keep, _mask_bits_to_keys(inline, B, H, N))       # inline draw == nv_dropout_bits, every used bit
    assert 0.6 < keep.mean().item() < 0.8 or N < 16
    ks = 65536.0 / (65536 - int(p_drop * 65536 + 0.5))
    qd = qkv.double().requires_grad_(True)
    q, k, v = qd.view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    sc = (q @ k.transpose(-1, -2)) * hd ** -0.5
    oref = ((sc.softmax(-1) * keep * ks) @ v).permute(0, 2, 1, 3).reshape(B * N, inner)
    assert rel_err(o, oref) < 1e-2
    assert rel_err(lse, torch.logsumexp(sc, -1)) < 1e-3
    dO = torch.randn(B * N, inner, device=DEV).to(torch.bfloat16)
    gref, = torch.autograd.grad(oref, qd, dO.double())
    dqkv = torch.full_like(qkv, float("nan"))
    ws = torch.empty(B * H * N, device=DEV)
    ops.attention_bwd(qkv, o, dO, lse, ws, dqkv, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=p_drop,
                      drop_mask=mask)
    assert torch.isfinite(dqkv.float()).all()
    if N == 1:
        # one key: P = 1 and dS = p (dP - delta) is exactly 0; the kernel's delta uses the bf16-rounded O, so its
        # dq / dk are rounding noise around zero — bound them against the size of the operands instead
        assert dqkv[:, :2 * inner].abs().max().item() < 2e-2 * dO.abs().max().item()
        assert rel_err(dqkv[:, 2 * inner:], gref[:, 2 * inner:]) < 2e-2
        return
    for nm, sl in (("dq", slice(0, inner)), ("dk", slice(inner, 2 * inner)), ("dv", slice(2 * inner, 3 * inner))):
        assert rel_err(dqkv[:, sl], gref[:, sl]) < 2e-2, nm
    # per-token check of the rows that are handled outside the tiles (token 0 as query and as key)
    g3, r3 = dqkv.view(B, N, 3 * inner).double(), gref.view(B, N, 3 * inner)
    assert rel_err(g3[:, 0], r3[:, 0]) < 2e-2
    assert rel_err(o.view(B, N, inner)[:, 0], oref.view(B, N, inner)[:, 0]) < 1e-2


def test_attention_fwd_rising_scores_rescale_path():
    """The forward keeps a lazily raised softmax reference (TMEM accumulator rescaled only when a block's
    maximum exceeds it by 2^8): scores that climb along the key axis force that path in every block for the
    rows with a positive query, while rows with a negative query never raise after the first block — both
    kinds share warps."""
    hd, B, N, H = 64, 2, 385, 2
    torch.manual_seed(11)
    inner = H * hd
    qkv = torch.randn(B * N, 3 * inner, device=DEV) * 0.05
    sign = torch.where(torch.rand(B * N, 1, device=DEV) < 0.5, -1.0, 1.0)
    ramp = (torch.arange(N, device=DEV, dtype=torch.float32) / N).repeat(B).unsqueeze(1)
    qkv[:, :inner] += sign                          # q = +-1 (+ noise)
    qkv[:, inner:2 * inner] += 8.0 * ramp           # k_j grows with j: scores span ~0 .. +-64 per row
    qkv = qkv.to(torch.bfloat16)
    o = torch.empty(B * N, inner, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=DEV)
    ops.attention_fwd(qkv, o, lse, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5)
    oref, lref = _attn_ref(qkv.double(), B, N, H, hd)
    assert torch.isfinite(o.float()).all()
    assert rel_err(o, oref) < 1e-2
    assert rel_err(lse, lref) < 1e-3


def test_softmax_fwd_bwd():
    torch.manual_seed(8)
    rows, n = 333, 385
    s = torch.randn(rows, n, device=DEV) * 3
    ref = s.double().requires_grad_(True)
    p = ref.softmax(-1)
    P = s.clone()
    ops.softmax_fwd(P, rows, n)
    assert rel_err(P, p) < 1e-5
    dP = torch.randn(rows, n, device=DEV)
    g, = torch.autograd.grad(p, ref, dP.double())
    ops.softmax_bwd(P, dP, rows, n)
    assert rel_err(dP, g) < 1e-4


# ---------------------------------------------------------------------------------------- helpers
def test_cast_colsum_batchsum_pool():
    torch.manual_seed(9)
    w = torch.randn(300, 136, device=DEV)
    wb, wt = ops.cast_transpose_bf16(w, out=torch.empty(300, 136, device=DEV, dtype=torch.bfloat16))
    assert torch.equal(wb, w.to(torch.bfloat16)) and torch.equal(wt, w.t().contiguous().to(torch.bfloat16))
    v = torch.randn(1027, device=DEV)
    assert torch.equal(ops.cast_bf16(v), v.to(torch.bfloat16))
    x = torch.randn(1000, 200, device=DEV)
    out = torch.ones(200, device=DEV)
    ops.colsum(x, out)
    assert rel_err(out, x.double().sum(0) + 1) < 1e-5
    xb = x.to(torch.bfloat16)
    out.zero_()
    ops.colsum(xb, out)
    assert rel_err(out, xb.double().sum(0)) < 1e-5
    y = torch.randn(7, 50, 32, device=DEV)
    acc = torch.zeros(50 * 32, device=DEV)
    ops.batch_sum(y, 50 * 32, acc, 7, 50 * 32)
    assert rel_err(acc, y.double().sum(0).flatten()) < 1e-5
    pooled = torch.empty(7, 32, device=DEV)
    ops.mean_pool_fwd(y, pooled, 7, 50, 32)
    assert rel_err(pooled, y.double().mean(1)) < 1e-5
    dx = torch.empty_like(y)
    ops.mean_pool_bwd(pooled, dx, None, 7, 50, 32)
    assert rel_err(dx, (pooled / 50).unsqueeze(1).expand_as(y)) < 1e-6


# --------------------------------------------------------------------------------------- temporal
def test_temporal_head_matches_torch():
    from neurovit_b200.functional import pack_temporal_params, TEMPORAL_KEYS
    torch.manual_seed(10)
    B, T, F = 5, 140, 2048
    layer = torch.nn.TransformerEncoderLayer(d_model=2, nhead=2, batch_first=True, dropout=0.0).to(DEV).double()
    enc = torch.nn.TransformerEncoder(layer, num_layers=1, enable_nested_tensor=False)
    head = torch.nn.Linear(2, 2).to(DEV).double()
    # keep both 2-element LayerNorms out of saturation (|x0 - x1| ~ sqrt(eps)), otherwise every gradient is
    # the round-off residue of an exact cancellation and nothing meaningful is compared
    with torch.no_grad():
        layer.norm1.weight.mul_(0.003)
        layer.norm1.bias.zero_()
    x = (0.005 * torch.randn(B, T, 2, device=DEV)).double().requires_grad_(True)
    ref = head(enc(x).mean(1))
    dout = torch.randn(B, 2, device=DEV).double()
    named = {"temporal." + k: v for k, v in enc.layers[0].named_parameters()}
    named.update({"head.weight": head.weight, "head.bias": head.bias})
    tensors = [named[k] for k in TEMPORAL_KEYS]
    grads = torch.autograd.grad(ref, [x] + tensors, dout)
    params = pack_temporal_params([t.detach().float() for t in tensors])
    out = torch.empty(B, 2, device=DEV)
    saved = torch.empty(B, T * 4, device=DEV)
    ops.temporal_fwd(x.detach().float(), params, out, saved, B, T, F)
    assert rel_err(out, ref) < 1e-4
    ws = torch.empty(B, params.numel(), device=DEV)
    dx = torch.empty(B, T, 2, device=DEV)
    ops.temporal_bwd(x.detach().float(), params, saved, dout.float(), ws, dx, B, T, F)
    gflat = torch.cat([g.flatten() for g in grads[1:]])
    assert rel_err(ws.double().sum(0), gflat) < 1e-3
    assert rel_err(dx, grads[0]) < 1e-3


def test_temporal_head_training_dropout_matches_masked_reference():
    """nn.TransformerEncoderLayer in training mode has four dropout sites (torch default p = 0.1); the kernel's
    Philox masks are replayed (ops.dropout_keep_mask draws the same bits) into a float64 restatement."""
    import math as _m
    from neurovit_b200.functional import pack_temporal_params, TEMPORAL_KEYS
    import torch.nn.functional as Fnn
    torch.manual_seed(13)
    B, T, F = 3, 140, 2048
    drop, seed = (0.1, 0.2, 0.1, 0.3), 4242
    layer = torch.nn.TransformerEncoderLayer(d_model=2, nhead=2, batch_first=True).to(DEV).double()
    head = torch.nn.Linear(2, 2).to(DEV).double()
    with torch.no_grad():
        layer.norm1.weight.mul_(0.003)
        layer.norm1.bias.zero_()
    x = (0.005 * torch.randn(B, T, 2, device=DEV)).double().requires_grad_(True)
    ks = lambda p: 65536.0 / (65536 - int(p * 65536 + 0.5))
    Tp, Fp = (T + 7) // 8 * 8, (F + 7) // 8 * 8
    km = lambda rows, cols, site: ops.dropout_keep_mask(rows, cols, p=drop[site], seed=seed, stream=site).double() * ks(drop[site])
    m_attn = km(B * 2 * T, Tp, 0).view(B, 2, T, Tp)[..., :T]
    flat2 = lambda site: km((B * T * 2 + 7) // 8, 8, site).view(-1)[:B * T * 2].view(B, T, 2)
    m1, m2 = flat2(1), flat2(3)
    m_ffn = km(B * T, Fp, 2).view(B, T, Fp)[..., :F]
    sa = layer.self_attn
    qkv = Fnn.linear(x, sa.in_proj_weight, sa.in_proj_bias)
    q, k, v = (t.reshape(B, T, 2, 1).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    attn = (q @ k.transpose(-1, -2)).softmax(-1) * m_attn                       # head_dim 1: scale 1
    o = (attn @ v).transpose(1, 2).reshape(B, T, 2)
    a = Fnn.linear(o, sa.out_proj.weight, sa.out_proj.bias) * m1
    x1 = Fnn.layer_norm(x + a, (2,), layer.norm1.weight, layer.norm1.bias)
    f = Fnn.linear(Fnn.relu(Fnn.linear(x1, layer.linear1.weight, layer.linear1.bias)) * m_ffn,
                   layer.linear2.weight, layer.linear2.bias) * m2
    x2 = Fnn.layer_norm(x1 + f, (2,), layer.norm2.weight, layer.norm2.bias)
    ref = head(x2.mean(1))
    dout = torch.randn(B, 2, device=DEV).double()
    named = {"temporal." + k_: v_ for k_, v_ in layer.named_parameters()}
    named.update({"head.weight": head.weight, "head.bias": head.bias})
    tensors = [named[k_] for k_ in TEMPORAL_KEYS]
    grads = torch.autograd.grad(ref, [x] + tensors, dout)
    params = pack_temporal_params([t.detach().float() for t in tensors])
    out = torch.empty(B, 2, device=DEV)
    saved = torch.empty(B, T * 4, device=DEV)
    ops.temporal_fwd(x.detach().float(), params, out, saved, B, T, F, drop=drop, seed=seed)
    assert rel_err(out, ref) < 1e-4
    ws = torch.empty(B, params.numel(), device=DEV)
    dx = torch.empty(B, T, 2, device=DEV)
    ops.temporal_bwd(x.detach().float(), params, saved, dout.float(), ws, dx, B, T, F, drop=drop, seed=seed)
    gflat = torch.cat([g.flatten() for g in grads[1:]])
    assert rel_err(ws.double().sum(0), gflat) < 1e-3
    assert rel_err(dx, grads[0]) < 1e-3
    assert 0.85 < (m_attn != 0).double().mean().item() < 0.95 and 0.65 < (m2 != 0).double().mean().item() < 0.75


@pytest.mark.parametrize("D,P,Kp", [(1024, 512, 512), (64, 729, 736), (128, 8, 8), (520, 40, 40)])
def test_layernorm_fold_into_linear_and_its_gradients(D, P, Kp):
    """nv_ln_fold / nv_ln_fold_grads (patch embedding, vit_3d.py:93-94): Linear(LN(p)) = xhat (W o gamma)^T + (W beta + b),
    and the gradients of W, b, gamma, beta from G = de^T xhat and cs = colsum(de), against autograd in float64."""
    torch.manual_seed(3)
    W = (torch.randn(D, P, device=DEV) * 0.1)
    g = 1 + 0.3 * torch.randn(P, device=DEV)
    bt = 0.2 * torch.randn(P, device=DEV)
    b = torch.randn(D, device=DEV)
    Wf = torch.full((D, Kp), float("nan"), device=DEV)
    bias_f = torch.empty(D, device=DEV)
    ops.ln_fold(W, g, bt, b, Wf, bias_f)
    assert torch.equal(Wf[:, :P], W * g) and (Wf[:, P:] == 0).all()
    assert rel_err(bias_f, b.double() + W.double() @ bt.double()) < 1e-6
    Wfb = torch.full((D, Kp), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.ln_fold(W, g, bt, None, Wfb, bias_f)
    assert torch.equal(Wfb[:, :P], (W * g).to(torch.bfloat16)) and (Wfb[:, P:] == 0).all()
    assert rel_err(bias_f, W.double() @ bt.double()) < 1e-6
    M = 300
    xhat = torch.randn(M, P, device=DEV, dtype=torch.float64)
    de = torch.randn(M, D, device=DEV, dtype=torch.float64)
    Wd, gd, btd, bd = (t.double().requires_grad_(True) for t in (W, g, bt, b))
    y = (xhat * gd + btd) @ Wd.t() + bd
    want = torch.autograd.grad((y * de).sum(), (Wd, gd, btd, bd))
    G = torch.zeros(D, Kp, device=DEV)
    G[:, :P] = (de.t() @ xhat).float()
    cs = de.sum(0).float()
    base = [torch.randn(D, P, device=DEV), torch.randn(P, device=DEV), torch.randn(P, device=DEV), torch.randn(D, device=DEV)]
    got = [t.clone() for t in base]                    # every output accumulates
    ops.ln_fold_grads(G, W, g, bt, cs, got[0], got[1], got[2], got[3])
    for nm, gt, b0, wt in zip(("dW", "dgamma", "dbeta", "db"), got, base, want):
        assert rel_err(gt - b0, wt) < 2e-5, nm


# ---------------------------------------------------------------------------------------- dropout
def test_dropout_kernel_mask_statistics_and_determinism():
    M, N, p = 1000, 1024, 0.1
    x = torch.randn(M, N, device=DEV)
    out = torch.empty_like(x)
    outb = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    cs = torch.zeros(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    ops.dropout(x, p=p, seed=123, stream=2, residual=res, out_f32=out, out_bf16=outb, colsum=cs)
    keep = ops.dropout_keep_mask(M, N, p=p, seed=123, stream=2)
    thr = int(p * 65536 + 0.5)
    ks = 65536.0 / (65536 - thr)
    assert abs(keep.mean().item() - (1 - thr / 65536)) < 3e-3
    assert torch.equal(out, x * keep * ks + res)                    # same mask, exact arithmetic
    assert rel_err(outb, out) < 1e-2
    assert rel_err(cs, (x * keep * ks).double().sum(0)) < 1e-4
    # a different stream or seed gives an independent mask
    other = ops.dropout_keep_mask(M, N, p=p, seed=123, stream=3)
    assert 0.70 < (other == keep).float().mean().item() < 0.90      # agreement of two Bernoulli(0.9) masks = 0.82
    # p = 0: plain copy + column sums
    cs.zero_()
    ops.dropout(x, p=0.0, seed=1, stream=0, out_f32=out, colsum=cs)
    assert torch.equal(out, x) and rel_err(cs, x.double().sum(0)) < 1e-4


@pytest.mark.parametrize("p", [0.1, 0.25, 0.5, 0.9])
def test_keep_bits_match_the_numpy_contract_bit_for_bit(p):
    """Integer work, pinned bit-exact: nv_dropout_bits (four Philox calls per 32-bit word, byte-transposed, bit-sliced
    compare) and the inline draw of nv_dropout (one call per 8 elements) both equal oracle/rng_oracle.py, which states
    the contract on explicitly assembled 16-bit uniforms (generator pinned to Random123's known-answer vectors in
    tests/test_oracle.py)."""
    import numpy as np
    from oracle import rng_oracle as R
    seed, stream = 0x1F2E3D4C5B6A7988 >> 2, 7
    epoch = ops.rng_epoch()
    n_groups = 4 * 5003                                  # ragged against the kernel's grid stride
    bits = torch.zeros(n_groups, device=DEV, dtype=torch.uint8)
    ops.dropout_bits(bits, p=p, seed=seed, stream=stream)
    want = R.dropout_bits(n_groups, p, seed, stream, epoch=epoch)
    assert np.array_equal(bits.cpu().numpy(), want)
    # element-indexed site drawn inline, incl. the row multiplier of the cls-row paths (index = row * row_mul * N + col)
    M, N = 37, 264
    ones = torch.ones(M, N, device=DEV)
    for row_mul in (1, 5):
        out = torch.empty_like(ones)
        ops.dropout(ones, p=p, seed=seed, stream=stream, out_f32=out, row_mul=row_mul)
        keep = (out != 0).cpu().numpy()
        assert np.array_equal(keep, R.keep_mask(M, N, p, seed, stream, epoch=epoch, row_mul=row_mul))
        ks = np.float32(65536.0) / np.float32(65536 - R.threshold(p))   # nv_dropout_keep_scale, in fp32
        assert np.array_equal(out.cpu().numpy(), np.where(keep, ks, np.float32(0)))
    # the epoch enters the seed: advance it and the same site draws the oracle's next-epoch bits
    ops.rng_epoch_advance()
    assert ops.rng_epoch() == epoch + 1
    ops.dropout_bits(bits, p=p, seed=seed, stream=stream)
    assert np.array_equal(bits.cpu().numpy(), R.dropout_bits(n_groups, p, seed, stream, epoch=epoch + 1))


def test_gemm_epilogue_dropout_matches_elementwise_mask():
    """The three GEMM epilogues draw the same (seed, stream, row * N + col) mask as nv_dropout."""
    torch.manual_seed(12)
    M, N, K, p, seed = 392, 512, 256, 0.25, 987654321
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) * 0.1).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    thr = int(p * 65536 + 0.5)
    ks = 65536.0 / (65536 - thr)
    lin = a.double() @ w.double().t() + bias.double()
    # store epilogue: dropout(linear) + residual
    keep = ops.dropout_keep_mask(M, N, p=p, seed=seed, stream=1).double()
    out = torch.empty(M, N, device=DEV)
    ops.gemm_bf16(a, w, bias=bias, residual=res, out_f32=out, dropout=(p, seed, 1))
    assert rel_err(out, lin * keep * ks + res.double()) < 2e-3
    # GELU epilogue: pre-activation untouched, activation dropped
    keep2 = ops.dropout_keep_mask(M, N, p=p, seed=seed, stream=2).double()
    pre = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    act = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm_bf16(a, w, bias=bias, out_bf16=act, out_pre=pre, apply_gelu=True, dropout=(p, seed, 2))
    assert rel_err(pre, lin) < 1e-2
    assert rel_err(act, torch.nn.functional.gelu(lin) * keep2 * ks) < 1e-2
    assert (act.double()[keep2 == 0] == 0).all()                    # dropped entries are exact zeros
    # dgrad through GELU + dropout: dY W * mask/(1-p) * gelu'(u)
    dy = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    wt = (torch.randn(K, N, device=DEV) * 0.1).to(torch.bfloat16)  # stored [N_out=K, K_in=N]
    u = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    du = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    cs = torch.zeros(N, device=DEV)
    ops.gemm_bf16(dy, wt, b_mn=True, gelu_u=u, out_bf16=du, colsum=cs, dropout=(p, seed, 2))
    ud = u.double().requires_grad_(True)
    gp, = torch.autograd.grad(torch.nn.functional.gelu(ud).sum(), ud)
    ref = (dy.double() @ wt.double()) * keep2 * ks * gp
    assert rel_err(du, ref) < 1e-2
    assert rel_err(cs, ref.sum(0)) < 1e-2


# ------------------------------------------------------------------------------ 4D input pipeline
@pytest.mark.parametrize("shape", [(2, 16, 16, 12, 140), (1, 18, 18, 18, 7), (3, 8, 8, 4, 300), (2, 5, 3, 2, 1),
                                   (1, 4, 4, 4, 257)])
def test_fmri_deinterleave_bit_exact(shape):
    """[B,H,W,D,T] -> [B*T,H,W,D] equals the reference's permute(0, 4, 1, 2, 3).reshape (NeuroEncoder.py:54-56)
    bit for bit: ragged S (not a multiple of the 64-wide CTA strip), T beyond one chunk, T = 1."""
    from neurovit_b200 import functional as Fn
    torch.manual_seed(11)
    B, H, W, D, T = shape
    x = torch.randn(*shape, device=DEV)
    y = Fn.fmri_to_volumes(x)
    ref = x.permute(0, 4, 1, 2, 3).reshape(B * T, H, W, D)
    assert y.shape == ref.shape and torch.equal(y, ref)
    # non-contiguous input (a permuted view) goes through the same path
    xv = torch.randn(B, T, H, W, D, device=DEV).permute(0, 2, 3, 4, 1)
    assert torch.equal(Fn.fmri_to_volumes(xv), xv.permute(0, 4, 1, 2, 3).reshape(B * T, H, W, D))
    # backward = the inverse interleave
    xg = x.clone().requires_grad_(True)
    g = torch.randn(B * T, H, W, D, device=DEV)
    Fn.fmri_to_volumes(xg).backward(g)
    xr = x.clone().requires_grad_(True)
    xr.permute(0, 4, 1, 2, 3).reshape(B * T, H, W, D).backward(g)
    assert torch.equal(xg.grad, xr.grad)


def test_fmri_deinterleave_zscore_matches_numpy_fp64():
    """Fused per-sample z-score (DatasetADNI_4D.py:84-86: (x - mean) / (std + 1e-8), numpy fp64, population std)."""
    from neurovit_b200 import functional as Fn
    torch.manual_seed(12)
    B, H, W, D, T = 2, 12, 10, 6, 140
    x = (torch.randn(B, H, W, D, T, device=DEV) * 37.0 + 410.0) * torch.tensor([1.0, 0.01], device=DEV).view(B, 1, 1, 1, 1)
    y = Fn.fmri_to_volumes(x, zscore=True)
    xn = x.cpu().numpy().astype("float64")
    for b in range(B):
        z = (xn[b] - xn[b].mean()) / (xn[b].std() + 1e-8)
        ref = torch.from_numpy(z).permute(3, 0, 1, 2).to(torch.float32)
        got = y[b * T:(b + 1) * T].cpu()
        assert (got - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
