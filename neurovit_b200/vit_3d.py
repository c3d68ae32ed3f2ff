"""Drop-in for the reference's src/models/vit_3d.py: same classes, constructor signatures, module tree
and state_dict keys; forward/backward run on the sm_100a C-ABI library through the
autograd Functions in neurovit_b200.functional. There is no torch-math or CPU fallback.

Module tree kept identical to the reference so checkpoints and hooks keep working:
  ViT.to_patch_embedding = Sequential(<rearrange>, LayerNorm, Linear, LayerNorm)     vit_3d.py:91-96
  ViT.pos_embedding, ViT.cls_token, ViT.dropout, ViT.transformer, ViT.to_latent, ViT.mlp_head   :98-110
  Transformer.layers = ModuleList([ModuleList([Attention, FeedForward])])             :62-69
  Attention.{norm, attend, dropout, to_qkv, to_out}                                   :28-46
  FeedForward.net = Sequential(LayerNorm, Linear, GELU, Dropout, Linear, Dropout)      :14-24
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import functional as Fn

_DEFAULT_PRECISION = os.environ.get("NEUROVIT_PRECISION", "bf16")


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


def _has_hooks(m: nn.Module) -> bool:
    return bool(m._forward_hooks or m._forward_pre_hooks or m._backward_hooks or m._backward_pre_hooks)


def _p(mod: nn.Module, drop: nn.Module) -> float:
    """Active dropout probability of an nn.Dropout child: its p in training mode, 0 in eval (vit_3d.py:21-23)."""
    return float(drop.p) if (mod.training and drop.training) else 0.0


class _PrecisionMixin:
    """`precision` is "bf16" (tensor-core path) or "fp32" (verification path); set it on any module of the
    tree with set_precision() — it is not a constructor argument, the reference signatures stay unchanged."""

    precision = _DEFAULT_PRECISION

    def set_precision(self, mode: str):
        Fn.check_mode(mode)
        for m in self.modules():
            if isinstance(m, _PrecisionMixin):
                m.precision = mode
        return self


class LayerNorm(nn.LayerNorm, _PrecisionMixin):
    """nn.LayerNorm whose forward is the CUDA kernel; output fp32. Forward/backward hooks registered on it
    (Grad-CAM, NeuroEncoder.py:70-82) observe the real output and its gradient."""

    def forward(self, x):
        return Fn.LayerNormFn.apply(x, self.weight, self.bias, self.eps, self.precision)


class PatchRearrange(nn.Module):
    """Place-holder for einops' Rearrange('b c (f pf) (h p1) (w p2) -> b (f h w) (p1 p2 pf c)') at index 0
    of to_patch_embedding (vit_3d.py:92): parameter-free, keeps the Sequential indices (and therefore the
    state_dict keys .1/.2/.3) identical. Called on its own it returns the gathered patches (CUDA kernel)."""

    def __init__(self, p1, p2, pf):
        super().__init__()
        self.p1, self.p2, self.pf = p1, p2, pf

    def extra_repr(self):
        return f"'b c (f pf) (h p1) (w p2) -> b (f h w) (p1 p2 pf c)', p1={self.p1}, p2={self.p2}, pf={self.pf}"

    def forward(self, video):
        from . import ops
        B, C, F_, H, W = video.shape
        n = (F_ // self.pf) * (H // self.p1) * (W // self.p2)
        P = C * self.pf * self.p1 * self.p2
        raw = torch.empty(B * n, P, device=video.device, dtype=torch.float32)
        ones = torch.ones(P, device=video.device)
        ops.patch_gather_ln(video.float(), (self.pf, self.p1, self.p2), ones, ones, None, raw=raw)
        return raw.view(B, n, P)


class FeedForward(nn.Module, _PrecisionMixin):
    def __init__(self, dim, hidden_dim, dropout=0.):
        super().__init__()
        self.net = nn.Sequential(
            LayerNorm(dim),
            nn.Linear(dim, hidden_dim),
            nn.GELU(),
            nn.Dropout(dropout),
            nn.Linear(hidden_dim, dim),
            nn.Dropout(dropout),
        )

    def drop_p(self):
        """(p after GELU, p after the down projection) active right now (0 in eval)."""
        return _p(self, self.net[3]), _p(self, self.net[5])

    def forward(self, x, residual=None, _drop=None):
        """Reference semantics: returns net(x) (vit_3d.py:25-26). With residual=x the add of vit_3d.py:74 is
        fused into the last GEMM's epilogue (used by Transformer.forward). _drop = (seed, stream base, prev):
        Transformer.forward shares one dropout seed per forward and tells each block which dropout site
        produced its input (prev = (p, stream)), so the backward can pre-mask that site's gradient."""
        ln, l1, l2 = self.net[0], self.net[1], self.net[4]
        p_gelu, p_down = self.drop_p()
        if _drop is not None:
            seed, sbase, prev = _drop
        else:
            seed, sbase, prev = (Fn.draw_seed() if (p_gelu > 0 or p_down > 0) else 0), 0, None
        if residual is None:
            residual_t = torch.zeros_like(x, dtype=torch.float32)
        else:
            residual_t = residual
        if residual is not None and residual is x and not _has_hooks(ln):
            return Fn.FFBlockFn.apply(x, ln.weight, ln.bias, l1.weight, l1.bias, l2.weight, l2.bias, ln.eps,
                                      self.precision, p_gelu, p_down, seed, sbase, prev)
        a = ln(x)  # module call: hooks on net[0] fire
        return Fn.FFCoreFn.apply(a, residual_t, l1.weight, l1.bias, l2.weight, l2.bias, self.precision, p_gelu,
                                 p_down, seed, sbase)


    def forward_cls(self, x_c, n_tok, _drop):
        """The residual block on the cls rows only, x_c [B, 1, D] (last layer under pool='cls'); masks are the dense
        block's (rows indexed b * n_tok). _drop = (seed, stream base)."""
        ln, l1, l2 = self.net[0], self.net[1], self.net[4]
        p_gelu, p_down = self.drop_p()
        return Fn.FFBlockClsFn.apply(x_c, ln.weight, ln.bias, l1.weight, l1.bias, l2.weight, l2.bias, ln.eps,
                                     self.precision, p_gelu, p_down, _drop[0], _drop[1], n_tok)


class Attention(nn.Module, _PrecisionMixin):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        project_out = not (heads == 1 and dim_head == dim)

        self.heads = heads
        self.scale = dim_head ** -0.5
        self.dim_head = dim_head

        self.norm = LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)

        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)

        self.to_out = nn.Sequential(
            nn.Linear(inner_dim, dim),
            nn.Dropout(dropout)
        ) if project_out else nn.Identity()

    def drop_p(self):
        """(p on the attention probabilities, p after to_out) active right now (0 in eval)."""
        if isinstance(self.to_out, nn.Identity):
            return _p(self, self.dropout), 0.0
        return _p(self, self.dropout), _p(self, self.to_out[1])

    def forward(self, x, residual=None, _drop=None):
        """Reference semantics: returns to_out(attention(norm(x))) (vit_3d.py:48-60). With residual=x the add
        of vit_3d.py:73 is fused into the to_out GEMM epilogue. _drop: see FeedForward.forward."""
        if isinstance(self.to_out, nn.Identity):
            raise NotImplementedError("project_out=False (heads == 1 and dim_head == dim) is not on the NeuroViT "
                                      "hot path (NeuroEncoder.py:181-195 uses heads=8, dim_head=64)")
        w_out, b_out = self.to_out[0].weight, self.to_out[0].bias
        p_attn, p_out = self.drop_p()
        if _drop is not None:
            seed, sbase, prev = _drop
        else:
            seed, sbase, prev = (Fn.draw_seed() if (p_attn > 0 or p_out > 0) else 0), 0, None
        if residual is not None and residual is x and not _has_hooks(self.norm):
            return Fn.AttnBlockFn.apply(x, self.norm.weight, self.norm.bias, self.to_qkv.weight, w_out, b_out,
                                        self.heads, self.dim_head, self.norm.eps, self.precision, p_attn, p_out, seed,
                                        sbase, prev)
        residual_t = torch.zeros_like(x, dtype=torch.float32) if residual is None else residual
        a = self.norm(x)  # real module call so Grad-CAM hooks on .norm observe output and grad_output
        return Fn.AttnCoreFn.apply(a, residual_t, self.to_qkv.weight, w_out, b_out, self.heads, self.dim_head,
                                   self.precision, p_attn, p_out, seed, sbase)


    def forward_cls(self, x, _drop):
        """Token 0 of attn(x) + x, [B, N, D] -> [B, 1, D] (last layer under pool='cls'). Hooks on .norm still see the
        whole LayerNorm output and its gradient. _drop = (seed, stream base, prev)."""
        w_out, b_out = self.to_out[0].weight, self.to_out[0].bias
        p_attn, p_out = self.drop_p()
        seed, sbase, prev = _drop
        if not _has_hooks(self.norm):
            return Fn.AttnBlockClsFn.apply(x, self.norm.weight, self.norm.bias, self.to_qkv.weight, w_out, b_out,
                                           self.heads, self.dim_head, self.norm.eps, self.precision, p_attn, p_out,
                                           seed, sbase, prev)
        a = self.norm(x)
        return Fn.AttnCoreClsFn.apply(a, x, self.to_qkv.weight, w_out, b_out, self.heads, self.dim_head,
                                      self.precision, p_attn, p_out, seed, sbase)


def _subtree_has_hooks(m: nn.Module, allow=()) -> bool:
    return any(_has_hooks(c) for c in m.modules() if not any(c is a for a in allow))


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout),
                FeedForward(dim, mlp_dim, dropout=dropout)
            ]))

    def _cls_last_ok(self, x) -> bool:
        """The last layer may run on the cls query / cls rows only: bf16 mode, a projecting attention, and no user hooks
        inside it other than on Attention.norm (which still sees every token) — a hook elsewhere would observe [B, 1, D]."""
        attn, ff = self.layers[-1]
        return (Fn.CLS_LAST and x.dim() == 3 and x.shape[1] > 1 and attn.precision == "bf16" and ff.precision == "bf16"
                and not isinstance(attn.to_out, nn.Identity) and attn.dim_head == 64
                and not _subtree_has_hooks(attn, allow=(attn.norm,)) and not _subtree_has_hooks(ff))

    def forward(self, x, _cls_last=False):
        """_cls_last (set by ViT.forward when pool == 'cls'): the caller uses token 0 of the result only, so the last
        layer is evaluated for that token alone and the result is [B, 1, D] (vit_3d.py:123 takes x[:, 0])."""
        # one dropout seed per forward; site streams = layer * 8 + {attn 0, to_out 1, gelu 2, down 3}
        drops = [(attn.drop_p(), ff.drop_p()) for attn, ff in self.layers]
        seed = Fn.draw_seed() if any(p > 0 for pair_ in drops for ps in pair_ for p in ps) else 0
        prev = None  # (p, stream) of the dropout site right before the residual add that produced x
        cls_last = bool(_cls_last) and len(self.layers) > 0 and self._cls_last_ok(x)
        for i, (attn, ff) in enumerate(self.layers):
            if cls_last and i == len(self.layers) - 1:
                n_tok = x.shape[1]
                x = attn.forward_cls(x, _drop=(seed, 8 * i, prev))
                return ff.forward_cls(x, n_tok, _drop=(seed, 8 * i))
            # x = attn(x) + x ; x = ff(x) + x   (vit_3d.py:72-74) with the adds fused into the GEMM epilogues;
            # modules that carry user hooks keep the reference's unfused call shape so the hooks see attn(x)
            if _has_hooks(attn):
                x = attn(x, _drop=(seed, 8 * i, None)) + x
            else:
                x = attn(x, residual=x, _drop=(seed, 8 * i, prev))
            prev = (drops[i][0][1], 8 * i + Fn.DROP_OUT) if not _has_hooks(attn) else None
            if _has_hooks(ff):
                x = ff(x, _drop=(seed, 8 * i, None)) + x
            else:
                x = ff(x, residual=x, _drop=(seed, 8 * i, prev))
            prev = (drops[i][1][1], 8 * i + Fn.DROP_DOWN) if not _has_hooks(ff) else None
        return x


class ViT(nn.Module, _PrecisionMixin):
    def __init__(self, *, image_size, image_patch_size, frames, frame_patch_size, num_classes, dim, depth, heads,
                 mlp_dim, pool='cls', channels=3, dim_head=64, dropout=0., emb_dropout=0.):
        super().__init__()
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(image_patch_size)

        assert image_height % patch_height == 0 and image_width % patch_width == 0, \
            'Image dimensions must be divisible by the patch size.'
        assert frames % frame_patch_size == 0, 'Frames must be divisible by frame patch size'

        num_patches = (image_height // patch_height) * (image_width // patch_width) * (frames // frame_patch_size)
        patch_dim = channels * patch_height * patch_width * frame_patch_size

        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'

        self.patch = (frame_patch_size, patch_height, patch_width)
        self.to_patch_embedding = nn.Sequential(
            PatchRearrange(p1=patch_height, p2=patch_width, pf=frame_patch_size),
            LayerNorm(patch_dim),
            nn.Linear(patch_dim, dim),
            LayerNorm(dim),
        )

        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)

        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)

        self.pool = pool
        self.to_latent = nn.Identity()

        self.mlp_head = nn.Sequential(
            LayerNorm(dim),
            nn.Linear(dim, num_classes)
        )

    def forward(self, video):
        """video [B, C, F, H, W] (any strides: the [B,1,D,H,W] view of ViT3DEncoder is gathered in place)
        -> logits [B, num_classes]. vit_3d.py:112-126."""
        if video.dim() != 5:
            raise ValueError(f"ViT expects a 5-D [B, C, F, H, W] tensor, got shape {tuple(video.shape)}")
        pf, p1, p2 = self.patch
        if video.shape[2] % pf or video.shape[3] % p1 or video.shape[4] % p2:
            raise ValueError(f"Shape mismatch: volume {tuple(video.shape[2:])} is not divisible by the patch size "
                             f"{(pf, p1, p2)}")
        if video.shape[0] == 0:  # empty batch: nothing to launch
            return video.new_zeros((0, self.mlp_head[1].out_features), dtype=torch.float32)
        pe = self.to_patch_embedding
        if pe[1].weight.numel() != video.shape[1] * pf * p1 * p2:
            raise ValueError(f"Shape mismatch: patch_dim {video.shape[1] * pf * p1 * p2} != LayerNorm dim "
                             f"{pe[1].weight.numel()} (channels differ from the constructor's)")
        x = Fn.PatchEmbedFn.apply(video, pe[1].weight, pe[1].bias, pe[2].weight, pe[2].bias, pe[3].weight,
                                  pe[3].bias, self.cls_token, self.pos_embedding, self.patch, pe[1].eps,
                                  self.precision)
        p_emb = _p(self, self.dropout)
        if p_emb > 0:  # x = dropout(x), vit_3d.py:119
            x = Fn.DropoutFn.apply(x, p_emb, Fn.draw_seed(), Fn.DROP_EMB)
        x = self.transformer(x, _cls_last=self.pool == 'cls')   # pool == 'cls': [B, 1, D], token 0 of the last layer
        h = self.mlp_head
        x = self.to_latent(x)
        return Fn.HeadFn.apply(x, h[0].weight, h[0].bias, h[1].weight, h[1].bias, self.pool, h[0].eps,
                               self.precision)
