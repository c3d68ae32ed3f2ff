"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's ViT3D / NeuroEncoder hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker (or the timed CPU baseline) — never as part of the product path.

Parity status: PINNED. The reference holds no tests or golden vectors of its own, so the
pin is against outputs of the reference itself: oracle/gen_golden.py imports the unmodified reference
(/root/reference/src/models/{vit_3d,NeuroEncoder}.py) in the build container, runs it on seeded inputs
and commits logits, loss and every parameter gradient under tests/golden/; tests/test_oracle.py checks
this restatement against those fixtures (bit-exact index mapping, fp32 round-off for the arithmetic).

The arithmetic of this path lives in PyTorch (unpinned; README badge "PyTorch 2.4", fixtures generated with
torch 2.11.0) and einops (unpinned, 0.8.2 here) — both third-party, neither vendored by the reference.
The restatement therefore uses torch *functional* CPU ops for the floating-point math (the same ATen
kernels the reference dispatches to, so fp32 results agree to round-off) and numpy closed-form integer
arithmetic for the patch-index mapping. Each function cites the reference lines it follows.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------ patch index mapping
def patch_index_map(C, Fr, H, W, pf, p1, p2):
    """Integer closed form of Rearrange('b c (f pf) (h p1) (w p2) -> b (f h w) (p1 p2 pf c)')
    (reference src/models/vit_3d.py:92). Returns int64 [n_tokens, patch_dim] of flat offsets into a
    contiguous [C, Fr, H, W] volume: out[t, j] = ((c*Fr + f)*H + h)*W + w with
      t = (fi*(H/p1) + hi)*(W/p2) + wi,  j = ((p1i*p2 + p2i)*pf + pfi)*C + c,
      f = fi*pf + pfi, h = hi*p1 + p1i, w = wi*p2 + p2i."""
    nf, nh, nw = Fr // pf, H // p1, W // p2
    fi, hi, wi, p1i, p2i, pfi, c = np.meshgrid(np.arange(nf), np.arange(nh), np.arange(nw), np.arange(p1),
                                               np.arange(p2), np.arange(pf), np.arange(C), indexing="ij")
    f = fi * pf + pfi
    h = hi * p1 + p1i
    w = wi * p2 + p2i
    off = ((c * Fr + f) * H + h) * W + w
    return off.reshape(nf * nh * nw, p1 * p2 * pf * C).astype(np.int64)


def patchify_np(video: np.ndarray, pf, p1, p2) -> np.ndarray:
    """video [B, C, F, H, W] (numpy, any dtype) -> patches [B, n, patch_dim] through the closed form."""
    B, C, Fr, H, W = video.shape
    idx = patch_index_map(C, Fr, H, W, pf, p1, p2)
    flat = np.ascontiguousarray(video).reshape(B, -1)
    return flat[:, idx]


def neuro_view(x: torch.Tensor) -> torch.Tensor:
    """ViT3DEncoder.forward layout adapter (reference src/models/NeuroEncoder.py:200-202):
    [B, H, W, D] -> permute(0, 3, 1, 2) -> unsqueeze(1) = non-contiguous view [B, 1, D, H, W]."""
    return x.permute(0, 3, 1, 2).unsqueeze(1)


# ----------------------------------------------------------------------------------------- ViT3D
_IDX_CACHE = {}


def _patch_index(C, Fr, H, W, pf, p1, p2, device):
    key = (C, Fr, H, W, pf, p1, p2, str(device))
    if key not in _IDX_CACHE:
        _IDX_CACHE[key] = torch.from_numpy(patch_index_map(C, Fr, H, W, pf, p1, p2)).to(device)
    return _IDX_CACHE[key]


def vit3d_forward(sd: dict, video: torch.Tensor, *, patch, heads, dim_head=64, pool="cls", prefix="", masks=None,
                  dropout_p=0.0, sdpa=False):
    """Functional forward of ViT (reference src/models/vit_3d.py:112-126) from a state_dict `sd`.
    patch = (pf, p1, p2). Dropout is the identity (p = 0 / eval) unless `masks` is given: a dict
    {"emb": m, (layer, "attn" | "out" | "gelu" | "down"): m} of multiplicative masks (keep / (1 - p), already
    scaled) applied exactly where the reference's nn.Dropout modules sit (vit_3d.py:21,23,39,45,100) — torch's
    own Philox stream cannot be replayed by a fused kernel, so dropout parity is checked with injected masks.
    sdpa=True swaps the explicit softmax attention of vit_3d.py:53-57 for F.scaled_dot_product_attention (same
    math; the faster stock-PyTorch variant bench.py --impl torch_gpu also times)."""
    g = lambda k: sd[prefix + k]
    if masks is not None:
        mk = lambda key, t: t * masks[key].to(t.dtype).reshape(t.shape)
    elif dropout_p > 0:  # training-mode nn.Dropout with torch's own generator (the timed CPU arm of bench.py)
        mk = lambda key, t: F.dropout(t, dropout_p, training=True)
    else:
        mk = lambda key, t: t
    pf, p1, p2 = patch
    B, C, Fr, H, W = video.shape
    # to_patch_embedding: Rearrange -> LN(patch_dim) -> Linear -> LN(dim)           vit_3d.py:91-96
    idx = _patch_index(C, Fr, H, W, pf, p1, p2, video.device)
    x = video.contiguous().reshape(B, -1)[:, idx]
    x = F.layer_norm(x, x.shape[-1:], g("to_patch_embedding.1.weight"), g("to_patch_embedding.1.bias"))
    x = F.linear(x, g("to_patch_embedding.2.weight"), g("to_patch_embedding.2.bias"))
    x = F.layer_norm(x, x.shape[-1:], g("to_patch_embedding.3.weight"), g("to_patch_embedding.3.bias"))
    n = x.shape[1]
    # cls token, positional embedding sliced to n+1                                  vit_3d.py:116-118
    x = torch.cat((g("cls_token").expand(B, 1, -1), x), dim=1)
    x = x + g("pos_embedding")[:, : n + 1]
    if (masks is not None and "emb" in masks) or (masks is None and dropout_p > 0):
        x = mk("emb", x)                                                                   # vit_3d.py:119
    depth = 0
    while f"{prefix}transformer.layers.{depth}.0.norm.weight" in sd:
        depth += 1
    scale = dim_head ** -0.5
    for i in range(depth):
        p = f"transformer.layers.{i}."
        # Attention                                                                vit_3d.py:48-60
        a = F.layer_norm(x, x.shape[-1:], g(p + "0.norm.weight"), g(p + "0.norm.bias"))
        qkv = F.linear(a, g(p + "0.to_qkv.weight"))
        q, k, v = (t.reshape(B, n + 1, heads, dim_head).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        if sdpa and masks is None:
            out = F.scaled_dot_product_attention(q, k, v, dropout_p=dropout_p, scale=scale)
        else:
            dots = torch.matmul(q, k.transpose(-1, -2)) * scale
            attn = mk((i, "attn"), dots.softmax(dim=-1))                                    # :55-56
            out = torch.matmul(attn, v)
        out = out.transpose(1, 2).reshape(B, n + 1, heads * dim_head)
        x = mk((i, "out"), F.linear(out, g(p + "0.to_out.0.weight"), g(p + "0.to_out.0.bias"))) + x  # :45,60,73
        # FeedForward                                                              vit_3d.py:17-26,74
        a = F.layer_norm(x, x.shape[-1:], g(p + "1.net.0.weight"), g(p + "1.net.0.bias"))
        u = F.linear(a, g(p + "1.net.1.weight"), g(p + "1.net.1.bias"))
        x = mk((i, "down"), F.linear(mk((i, "gelu"), F.gelu(u)), g(p + "1.net.4.weight"), g(p + "1.net.4.bias"))) + x
    x = x.mean(dim=1) if pool == "mean" else x[:, 0]                                       # :123
    x = F.layer_norm(x, x.shape[-1:], g("mlp_head.0.weight"), g("mlp_head.0.bias"))
    return F.linear(x, g("mlp_head.1.weight"), g("mlp_head.1.bias"))                       # :108-110,126


def vit3d_loss_and_grads(sd: dict, video, labels, **kw):
    """Mean cross-entropy (reference src/Trainer.py:30,70) and d(loss)/d(param) for every tensor in sd."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    logits = vit3d_forward(leaf, video, **kw)
    loss = F.cross_entropy(logits, labels)
    used = [k for k in leaf]
    grads = torch.autograd.grad(loss, [leaf[k] for k in used], allow_unused=True)
    gd = {k: (torch.zeros_like(leaf[k]) if gr is None else gr) for k, gr in zip(used, grads)}
    return logits.detach(), loss.detach(), gd


# ------------------------------------------------------------------------------ 4D temporal head
def temporal_forward(sd: dict, x: torch.Tensor, *, prefix_t="temporal_transformer.transformer.layers.0.",
                     prefix_p="projection_head.projection_head."):
    """TemporalTransformer + mean over T + ProjectionHead (reference src/models/NeuroEncoder.py:63-66,
    207-230): nn.TransformerEncoderLayer(d_model=2, nhead=2, batch_first=True) defaults = post-norm, ReLU,
    LN eps 1e-5; dropout identity. x [B, T, 2] -> [B, 2]."""
    g = lambda k: sd[prefix_t + k]
    B, T, E = x.shape
    nhead = 2
    hd = E // nhead
    qkv = F.linear(x, g("self_attn.in_proj_weight"), g("self_attn.in_proj_bias"))
    q, k, v = (t.reshape(B, T, nhead, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    attn = (torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)).softmax(dim=-1)
    o = torch.matmul(attn, v).transpose(1, 2).reshape(B, T, E)
    a = F.linear(o, g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias"))
    x1 = F.layer_norm(x + a, (E,), g("norm1.weight"), g("norm1.bias"))
    f = F.linear(F.relu(F.linear(x1, g("linear1.weight"), g("linear1.bias"))), g("linear2.weight"), g("linear2.bias"))
    x2 = F.layer_norm(x1 + f, (E,), g("norm2.weight"), g("norm2.bias"))
    m = x2.mean(dim=1)                                                               # NeuroEncoder.py:64
    return F.linear(m, sd[prefix_p + "weight"], sd[prefix_p + "bias"])                # :66,223-228


def neuroencoder_forward(sd: dict, fmri: torch.Tensor, *, patch, heads=8, dim_head=64, training_dim=3):
    """NeuroEncoder.forward (reference src/models/NeuroEncoder.py:49-68). fmri [B,H,W,D] (3D) or
    [B,H,W,D,T] (4D). sd uses the NeuroEncoder key prefixes."""
    kw = dict(patch=(patch, patch, patch), heads=heads, dim_head=dim_head, prefix="volume_encoder.vit3d.")
    if training_dim == 3:
        return vit3d_forward(sd, neuro_view(fmri), **kw)
    fmri = fmri.permute(0, 4, 1, 2, 3)
    B, T, H, W, D = fmri.shape
    vols = fmri.reshape(B * T, H, W, D)
    enc = vit3d_forward(sd, neuro_view(vols), **kw).reshape(B, T, -1)
    return temporal_forward(sd, enc)


# ------------------------------------------------------------------------------ Grad-CAM fixture recipe
def perturb_for_cam(vit, seed=77):
    """Seeded in-place perturbation shared by oracle/gen_golden.py (applied to the reference model) and
    tests/test_gpu_model.py (applied to the drop-in model) before the Grad-CAM comparison. At initialisation the
    reference's Grad-CAM (NeuroEncoder.py:100-107) is pure round-off: the hooked LayerNorm has gamma = 1, beta = 0,
    so its output sums to zero over the features and cam = mean_k(grad) * sum_k(act) = 0. Trained-like affine
    parameters on that LayerNorm, per-row offsets on the last to_qkv (non-zero feature mean of the gradient) and a
    larger head make the map a well-conditioned function of the hot path (condition number ~15)."""
    gen = torch.Generator().manual_seed(seed)
    last = vit.transformer.layers[-1][0]
    with torch.no_grad():
        dev = last.to_qkv.weight.device
        last.to_qkv.weight.add_((0.03 * torch.randn(last.to_qkv.weight.shape[0], 1, generator=gen)).to(dev))
        vit.mlp_head[1].weight.mul_(20.0)
        last.norm.weight.add_((0.3 * torch.randn(last.norm.weight.shape[0], generator=gen)).to(dev))
        last.norm.bias.add_((0.1 * torch.randn(last.norm.bias.shape[0], generator=gen) + 0.05).to(dev))


def gradcam_from_hooks(gradients, activations, grid, patch, threshold):
    """Restatement of NeuroEncoder.get_attention_map steps 1-6 (NeuroEncoder.py:100-131) from the two hooked
    tensors [1, tokens, dim]: feature-mean gradient weights, weighted activation sum, drop cls, reshape to the
    patch grid, ReLU, min-max normalise, keep the top `threshold` percent, trilinear upsample to grid^3."""
    weights = gradients.mean(dim=2, keepdim=True)
    cam = (weights * activations).sum(dim=2)[:, 1:]
    cs = grid // patch
    cam = F.relu(cam.reshape(1, cs, cs, cs))
    cam = (cam - cam.min()) / (cam.max() - cam.min() + 1e-8)
    thr = np.percentile(cam.numpy(), 100 - threshold)
    kept = torch.from_numpy(np.where(cam.numpy() >= thr, cam.numpy(), 0)).unsqueeze(0)
    return F.interpolate(kept, size=(grid, grid, grid), mode="trilinear", align_corners=False).squeeze()
