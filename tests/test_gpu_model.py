"""Module-level parity on the GPU: the drop-in ViT / NeuroEncoder (CUDA kernels through the C ABI) against
the CPU oracle on the same seeded inputs and weights, and against the committed golden fixtures produced
by the unmodified reference. Tolerances from BASELINE.json north_star: 2e-2 relative (bf16 operands, fp32
accumulation), 1e-5 for the fp32 verification mode, logits and every gradient, measured as max|a-b| / max|b| per
tensor. Where a test holds a quantity to a different bound the reason is stated next to it."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from neurovit_b200.vit_3d import ViT  # noqa: E402
from neurovit_b200.NeuroEncoder import NeuroEncoder  # noqa: E402
from oracle import vit3d_oracle as O  # noqa: E402  (checker only)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"
TOL = {"bf16": dict(logits=2e-2, grad=2e-2), "fp32": dict(logits=1e-5, grad=1e-5)}


def rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def _golden_vit(name, ctor):
    g = np.load(os.path.join(GOLD, name))
    m = ViT(**ctor)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    m.load_state_dict(sd, strict=True)  # keys must match the reference 1:1
    return g, m.to(DEV).eval()


SMALL = dict(image_size=16, image_patch_size=8, frames=16, frame_patch_size=8, num_classes=2, dim=64, depth=2,
             heads=2, mlp_dim=128, channels=1, dim_head=64)
P4 = dict(image_size=(8, 12), image_patch_size=4, frames=4, frame_patch_size=4, num_classes=3, dim=64, depth=1,
          heads=2, mlp_dim=64, channels=2, dim_head=64, pool="mean")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name,ctor", [("vit3d_small.npz", SMALL), ("vit3d_p4.npz", P4)])
def test_vit_matches_reference_golden(name, ctor, mode):
    g, m = _golden_vit(name, ctor)
    m.set_precision(mode)
    video = torch.from_numpy(g["video"]).to(DEV)
    labels = torch.from_numpy(g["labels"]).to(DEV)
    logits = m(video)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    tol = TOL[mode]
    assert rel(logits, g["logits"]) < tol["logits"]
    worst = {}
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        worst[k] = rel(p.grad, g["grad." + k])
    bad = {k: v for k, v in worst.items() if v >= tol["grad"]}
    assert not bad, f"gradient mismatch ({mode}): {bad}"


def _cfg(dim, tmp, grid=16, patch=8, precision="bf16"):
    return {"DEVICE": DEV, "TRAINING_DIM": dim, "TRAINING_DROPOUT": 0.0, "TRAINING_VIT_INPUT_SIZE": grid,
            "GRADCAM_CUBE_SIZE": 8, "TRAINING_VIT_PATCH_SIZE": patch, "DATASET_NAME": "adni",
            "GLOBAL_BASE_PATH": tmp, "BEST_MODEL_PATH": "best.pth", "GRADCAM_THRESHOLD": 10, "GRADCAM_SLICE_DIM": 0,
            "GRADCAM_SLICE_IDX": 3, "PRECISION": precision}


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_neuroencoder_3d_golden_with_gradcam_hooks(tmp_path, mode):
    """Full-size model dims (1024/6/8/2048) through ViT3DEncoder's permuted view, with the Grad-CAM hooks of
    NeuroEncoder.py:70-82 live on layers[-1][0].norm."""
    g = np.load(os.path.join(GOLD, "neuro3d.npz"))
    torch.manual_seed(1234)  # same seed as oracle/gen_golden.py: identical initial weights (checksum pinned)
    m = NeuroEncoder(_cfg(3, str(tmp_path), precision=mode)).eval()
    checksum = float(sum(v.double().sum() for v in m.state_dict().values()))
    assert abs(checksum - float(g["sd_checksum"][0])) < 1e-4
    x = torch.from_numpy(g["x"]).to(DEV)
    logits = m(x)
    loss = torch.nn.functional.cross_entropy(logits, torch.from_numpy(g["labels"]).to(DEV))
    loss.backward()
    tol = TOL[mode]
    assert rel(logits, g["logits"]) < tol["logits"]
    params = dict(m.named_parameters())
    for k in [f[5:] for f in g.files if f.startswith("grad.")]:
        assert rel(params[k].grad, g["grad." + k]) < tol["grad"], k
    gn = float(sum((p.grad.double() ** 2).sum() for p in m.parameters()) ** 0.5)
    assert abs(gn - float(g["gradnorm"][0])) / float(g["gradnorm"][0]) < tol["grad"]
    # hooks observed the LayerNorm output and its gradient, on the host, like the reference
    assert m.activations.device.type == "cpu" and tuple(m.activations.shape) == (2, 9, 1024)
    assert rel(m.activations[:, :3, :8], g["act_hook"]) < tol["logits"]
    assert rel(m.gradients[:, :3, :8], g["grad_hook"]) < tol["grad"]


def test_neuroencoder_gradcam_map(tmp_path):
    torch.manual_seed(1234)
    m = NeuroEncoder(_cfg(3, str(tmp_path), precision="fp32")).eval()
    x = torch.randn(1, 16, 16, 16, device=DEV)
    cam, cls = m.get_attention_map(x)
    assert tuple(cam.shape) == (16, 16, 16) and cls.shape == (1,)
    assert torch.isfinite(cam).all() and cam.min() >= 0 and cam.max() <= 1 + 1e-6
    img, attn = m.visualize_slice(cam, x)
    assert img.shape == (16, 16) and tuple(attn.shape) == (16, 16)
    assert m.activations.shape[1] == 9 and m.gradients.shape == m.activations.shape


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_neuroencoder_gradcam_map_matches_reference(tmp_path, mode):
    """get_attention_map against the map the unmodified reference produced (tests/golden/neuro3d_cam.npz, 32^3 volume,
    4x4x4 patch tokens, weights regenerated from the seed + oracle.perturb_for_cam and pinned by checksum).
    The raw token scores have condition number ~15 (sum of 1024 signed products per token), so bf16 is held to
    15 x 0.3 % = 5e-2 on them; the fp32 verification mode to 1e-4 and to the reference's final thresholded map."""
    g = np.load(os.path.join(GOLD, "neuro3d_cam.npz"))
    torch.manual_seed(1234)
    cfg = {**_cfg(3, str(tmp_path), grid=32, precision=mode), "GRADCAM_THRESHOLD": 25}
    m = NeuroEncoder(cfg).eval()
    O.perturb_for_cam(m.volume_encoder.vit3d)
    checksum = float(sum(v.double().sum() for v in m.state_dict().values()))
    assert abs(checksum - float(g["sd_checksum"][0])) < 1e-3
    cam, cls = m.get_attention_map(torch.from_numpy(g["x"]).to(DEV))
    assert cls.cpu().tolist() == g["cls"].tolist()
    grads, acts = m.gradients.float().cpu(), m.activations.float().cpu()
    raw = (grads.mean(dim=2, keepdim=True) * acts).sum(dim=2)[:, 1:]
    assert rel(raw, g["raw_cam"]) < (1e-4 if mode == "fp32" else 5e-2)
    # the returned map is exactly the reference's post-processing (NeuroEncoder.py:100-131) of the hooked tensors
    assert torch.equal(cam.cpu(), O.gradcam_from_hooks(grads, acts, 32, 8, 25))
    if mode == "fp32":
        assert (cam.cpu() - torch.from_numpy(g["cam"])).abs().max().item() < 1e-4


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_neuroencoder_4d_golden(tmp_path, mode):
    g = np.load(os.path.join(GOLD, "neuro4d.npz"))
    torch.manual_seed(1234)
    m3 = NeuroEncoder(_cfg(3, str(tmp_path)))
    torch.save(m3.state_dict(), os.path.join(str(tmp_path), "best.pth"))
    torch.manual_seed(4321)
    m = NeuroEncoder(_cfg(4, str(tmp_path), precision=mode)).eval()
    for k, v in m.state_dict().items():
        if not k.startswith("volume_encoder."):
            assert np.array_equal(v.cpu().numpy(), g["sd." + k]), k
    assert not any(p.requires_grad for p in m.volume_encoder.parameters())
    out = m(torch.from_numpy(g["x"]).to(DEV))
    loss = torch.nn.functional.cross_entropy(out, torch.from_numpy(g["labels"]).to(DEV))
    loss.backward()
    # the frozen ViT's 2 logits feed LayerNorms over 2 elements; the bf16 ViT perturbs its inputs by ~1e-2
    assert rel(out, g["out"]) < (5e-2 if mode == "bf16" else 1e-4)
    params = dict(m.named_parameters())
    for k in [f[5:] for f in g.files if f.startswith("grad.")]:
        assert params[k].grad is not None, k
        if "norm2" in k or "projection_head" in k:  # the well-conditioned gradients (see tests/test_oracle.py)
            assert rel(params[k].grad, g["grad." + k]) < (5e-2 if mode == "bf16" else 1e-3), k


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_vit_vs_oracle_ragged_tokens(mode):
    """cfgA (BASELINE configs[1] geometry and model dims: 64x64x48 / patch 8 -> 385 tokens, not a multiple of
    any tile; dim 1024, depth 6, heads 8, mlp 2048) against the CPU oracle run on the same weights; direct ViT
    call with a contiguous [B,1,F,H,W] tensor (the permuted-K patch layout)."""
    torch.manual_seed(11)
    ctor = dict(image_size=64, image_patch_size=8, frames=48, frame_patch_size=8, num_classes=2, dim=1024, depth=6,
                heads=8, mlp_dim=2048, channels=1, dim_head=64)
    m = ViT(**ctor)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    # Labels are all the same class on purpose: at random init every sample produces nearly the same logits, so
    # with opposite labels the per-sample gradients of the batch-summed parameters (pos_embedding, cls_token,
    # patch-embedding biases) cancel to ~3% of their size and a 0.3% bf16 rounding error on each reads as 5-6%
    # of the (ill-conditioned) sum. Same-class labels keep the comparison well-conditioned.
    video = torch.randn(2, 1, 48, 64, 64)
    labels = torch.tensor([1, 1])
    ref_logits, ref_loss, ref_grads = O.vit3d_loss_and_grads(sd, video, labels, patch=(8, 8, 8), heads=8)
    m = m.to(DEV).eval().set_precision(mode)
    logits = m(video.to(DEV))
    torch.nn.functional.cross_entropy(logits, labels.to(DEV)).backward()
    tol = TOL[mode]
    assert rel(logits, ref_logits) < tol["logits"]
    bad = {k: rel(p.grad, ref_grads[k]) for k, p in m.named_parameters() if rel(p.grad, ref_grads[k]) >= tol["grad"]}
    assert not bad, bad


FULL = dict(num_classes=2, dim=1024, depth=6, heads=8, mlp_dim=2048, channels=1, dim_head=64)


def _perturbed_vit(ctor, seed):
    torch.manual_seed(seed)
    m = ViT(**ctor)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def _check_vs_oracle(m, sd, video, labels, patch, heads, mode, oracle_video=None):
    ref_logits, _, ref_grads = O.vit3d_loss_and_grads(sd, video if oracle_video is None else oracle_video, labels,
                                                      patch=(patch,) * 3, heads=heads)
    m = m.to(DEV).eval().set_precision(mode)
    logits = m(video.to(DEV))
    torch.nn.functional.cross_entropy(logits, labels.to(DEV)).backward()
    tol = TOL[mode]
    assert rel(logits, ref_logits) < tol["logits"]
    worst = {k: rel(p.grad, ref_grads[k]) for k, p in m.named_parameters()}
    bad = {k: v for k, v in worst.items() if v >= tol["grad"]}
    assert not bad, bad
    return worst


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_vit_vs_oracle_config4_long_sequence(mode):
    """BASELINE configs[3]: 1x96x96x96 volume, patch 8 -> 1728 patches + cls = 1729 tokens, full model dims, one
    volume, through the ViT3DEncoder view ([B,H,W,D] -> permute(0,3,1,2).unsqueeze(1), NeuroEncoder.py:200-202)."""
    ctor = dict(image_size=96, image_patch_size=8, frames=96, frame_patch_size=8, **FULL)
    m, sd = _perturbed_vit(ctor, 13)
    x = torch.randn(1, 96, 96, 96)
    _check_vs_oracle(m, sd, O.neuro_view(x), torch.tensor([1]), 8, 8, mode)


@pytest.mark.parametrize("grid", [18, 90])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_vit_vs_oracle_patch9(mode, grid):
    """The reference's shipped default geometry (configs/config.yaml:39-40: 90^3 volume, patch 9 -> 1000 patches,
    patch_dim 729 padded to 736 for the tensor-core path) and its 2x2x2-token miniature; strided-gather kernels
    (36-byte patch rows are not a TMA box)."""
    ctor = dict(image_size=grid, image_patch_size=9, frames=grid, frame_patch_size=9, **FULL)
    m, sd = _perturbed_vit(ctor, 14 + grid)
    B = 2 if grid == 18 else 1
    x = torch.randn(B, grid, grid, grid)
    _check_vs_oracle(m, sd, O.neuro_view(x), torch.ones(B, dtype=torch.int64), 9, 8, mode)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_vit_per_sample_gradients_opposite_labels(mode):
    """Companion of test_vit_vs_oracle_ragged_tokens, which uses same-class labels to keep the batch-summed
    gradients well conditioned: here the two samples carry OPPOSITE labels and each sample's gradient is checked
    on its own (batch of one), so no cancellation between samples hides or inflates an error."""
    ctor = dict(image_size=16, image_patch_size=8, frames=24, frame_patch_size=8, **{**FULL, "depth": 2})
    m, sd = _perturbed_vit(ctor, 15)
    video = torch.randn(2, 1, 24, 16, 16)
    for i, lab in enumerate((0, 1)):
        m.zero_grad(set_to_none=True)
        _check_vs_oracle(m, sd, video[i:i + 1], torch.tensor([lab]), 8, 8, mode)


def test_head_with_many_classes_gradcam_dataset():
    """DATASET_NAME == 'gradcam' builds num_classes = (grid // cube)^3 (NeuroEncoder.py:179): 512 classes at grid 64,
    beyond the 256 the fused one-CTA-per-sample head kernels hold; the generic LayerNorm + linear path takes over."""
    ctor = dict(image_size=16, image_patch_size=8, frames=16, frame_patch_size=8, num_classes=512, dim=128, depth=1,
                heads=2, mlp_dim=256, channels=1, dim_head=64)
    for mode in ("fp32", "bf16"):
        m, sd = _perturbed_vit(ctor, 16)
        video = torch.randn(3, 1, 16, 16, 16)
        _check_vs_oracle(m, sd, video, torch.tensor([5, 300, 511]), 8, 2, mode)


def test_errors_and_contract():
    m = ViT(**SMALL).to(DEV)
    with pytest.raises(ValueError):
        m(torch.randn(1, 1, 16, 16, 12, device=DEV))          # not divisible by the patch size
    with pytest.raises(Exception):
        m(torch.randn(1, 1, 16, 16, 16))                       # CPU tensor: no CPU fallback
    with pytest.raises(AssertionError):
        ViT(**{**SMALL, "pool": "max"})
    # smaller volume than configured: pos_embedding is sliced to n+1 (vit_3d.py:118)
    out = m.eval()(torch.randn(1, 1, 8, 16, 16, device=DEV))
    assert out.shape == (1, 2)
    # B = 0 (empty batch) is a no-op that keeps shapes
    assert m(torch.randn(0, 1, 16, 16, 16, device=DEV)).shape == (0, 2)


def test_trainer_grad_sinks_match_autograd_accumulation():
    """DataParallelTrainer.step lets the wgrad / LayerNorm-backward kernels add straight into the flat
    gradient buffer (functional.SINKS); the result must equal plain autograd accumulation into .grad."""
    from neurovit_b200.trainer import DataParallelTrainer
    torch.manual_seed(3)
    ctor = dict(image_size=16, image_patch_size=8, frames=16, frame_patch_size=8, num_classes=2, dim=128, depth=2,
                heads=2, mlp_dim=256, channels=1, dim_head=64)
    m = ViT(**ctor).to(DEV)
    x = torch.randn(6, 1, 16, 16, 16, device=DEV)
    y = torch.randint(0, 2, (6,), device=DEV)
    torch.nn.functional.cross_entropy(m(x), y).backward()
    want = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    from neurovit_b200.functional import SINKS
    tr = DataParallelTrainer(m, optimizer=torch.optim.SGD(m.parameters(), lr=0.0))
    sunk0 = SINKS.sunk
    for _ in range(2):  # second step: the flat buffer is re-zeroed, nothing carries over
        tr.step(x, y)
    torch.cuda.synchronize()
    assert SINKS.sunk - sunk0 == 2 * (2 * 9 + 4 + 1), "per layer 9 tensors (2 LayerNorm pairs, w_qkv, w_out, w1, b1, w2), 4 in the patch embedding, the head weight"
    for k, p in m.named_parameters():
        assert p.grad.data_ptr() % 128 == 0
        assert rel(p.grad, want[k]) < 1e-4, k


def _replay_masks(trace, B, N, heads, ks_of):
    """Rebuild the oracle's mask dict from functional.DROPOUT_TRACE records (forward order: emb, then per layer
    attn, out, gelu, down)."""
    from neurovit_b200 import ops
    masks, layer = {}, 0
    for site, p, seed, stream, info in trace:
        ks = ks_of(p)
        if site == "emb":
            masks["emb"] = ops.dropout_keep_mask(info[0], info[1], p=p, seed=seed, stream=stream).cpu() * ks
        elif site == "attn":      # saved bit mask [B*H, N, ceil(N/32)] of the flash kernel
            # bit position p of a row = key token p + 1 (p < N - 1) or key token 0 (p = N - 1): attention_tc.cu
            w = info.view(B, heads, N, -1).to(torch.int64) & 0xFFFFFFFF
            bits = ((w.unsqueeze(-1) >> torch.arange(32, device=w.device)) & 1).reshape(B, heads, N, -1)[..., :N]
            bits = torch.cat([bits[..., N - 1:N], bits[..., :N - 1]], dim=-1)
            masks[(layer, "attn")] = bits.float().cpu() * ks
        elif site == "attn_flat":  # fp32 mode: flat element index over [B, H, N, N]
            L = info[0] * info[1] * info[2] * info[3]
            body = L // 8 * 8
            m = torch.ones(L)
            m[:body] = ops.dropout_keep_mask(body // 8, 8, p=p, seed=seed, stream=stream).cpu().view(-1)
            if L - body:
                m[body:] = ops.dropout_keep_mask(1, 8, p=p, seed=seed, stream=stream + 1000).cpu().view(-1)[:L - body]
            masks[(layer, "attn")] = m.view(info) * ks
        else:
            masks[(layer, site)] = ops.dropout_keep_mask(info[0], info[1], p=p, seed=seed, stream=stream).cpu() * ks
            if site == "down":
                layer += 1
    return masks


@pytest.mark.parametrize("batch", [6, 64])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_vit_training_dropout_matches_oracle_with_replayed_masks(mode, batch):
    """Dropout p > 0 in training mode (the reference's default TRAINING_DROPOUT 0.1 at all 25 sites): the masks
    the kernels drew are replayed into the CPU oracle; logits and every gradient must then agree."""
    from neurovit_b200 import functional as Fn
    torch.manual_seed(21)
    ctor = dict(image_size=16, image_patch_size=8, frames=24, frame_patch_size=8, num_classes=2, dim=64, depth=2,
                heads=2, mlp_dim=128, channels=1, dim_head=64, dropout=0.2, emb_dropout=0.1)
    m = ViT(**ctor)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    # 6 x 13 tokens = 78 rows >= 64: the LayerNorm-backward side-car masking path; batch 64 also takes the pipelined
    # LayerNorm backward on the 64 cls rows of the last block (row maps + side-car mask with the original row index)
    x = torch.randn(batch, 1, 24, 16, 16)
    y = torch.randint(0, 2, (batch,))
    m = m.to(DEV).train().set_precision(mode)
    Fn.DROPOUT_TRACE.record = []
    try:
        logits = m(x.to(DEV))
        trace = Fn.DROPOUT_TRACE.record
    finally:
        Fn.DROPOUT_TRACE.record = None
    loss = torch.nn.functional.cross_entropy(logits, y.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    n_tok = (24 // 8) * (16 // 8) * (16 // 8) + 1
    assert len(trace) == 1 + 2 * 4, [t[0] for t in trace]
    ks_of = lambda p: 65536.0 / (65536 - int(p * 65536 + 0.5))
    masks = _replay_masks(trace, batch, n_tok, 2, ks_of)
    for k_, v_ in masks.items():
        assert 0.5 < (v_ != 0).float().mean().item() < 0.98, k_   # really dropping, at roughly the asked rate
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_logits = O.vit3d_forward(leaf, x, patch=(8, 8, 8), heads=2, masks=masks)
    ref_loss = torch.nn.functional.cross_entropy(ref_logits, y)
    grads = torch.autograd.grad(ref_loss, list(leaf.values()), allow_unused=True)
    tol = TOL[mode]
    assert rel(logits, ref_logits) < tol["logits"]
    for (k, p_), gr in zip(m.named_parameters(), grads):
        gr = torch.zeros_like(leaf[k]) if gr is None else gr
        assert rel(p_.grad, gr) < tol["grad"], k
    # eval mode is deterministic and dropout-free
    m.eval()
    a, b = m(x.to(DEV)), m(x.to(DEV))
    assert torch.equal(a, b)


@pytest.mark.parametrize("p_drop", [0.0, 0.2])
@pytest.mark.parametrize("hooked,depth,batch", [(False, 2, 64), (True, 2, 64), (False, 1, 64), (False, 3, 3)])
def test_cls_only_last_layer_equals_dense_last_layer(p_drop, hooked, depth, batch):
    """pool='cls' (vit_3d.py:123): the last layer evaluated for the cls query / cls rows only (AttnBlockClsFn,
    FFBlockClsFn; AttnCoreClsFn when Attention.norm of that layer carries hooks, as under NeuroEncoder) gives the dense
    layer's logits and gradients: same dropout masks (indexed by the original rows), same arithmetic per row. What a
    hook on that LayerNorm sees (all tokens, output and gradient) is the same too."""
    from neurovit_b200 import functional as Fn
    torch.manual_seed(5)
    ctor = dict(image_size=16, image_patch_size=8, frames=24, frame_patch_size=8, num_classes=2, dim=128, depth=depth,
                heads=2, mlp_dim=256, channels=1, dim_head=64, dropout=p_drop, emb_dropout=p_drop / 2)
    m = ViT(**ctor).to(DEV).train()
    seen = {}
    if hooked:
        norm = m.transformer.layers[-1][0].norm
        norm.register_forward_hook(lambda mod, inp, out: seen.__setitem__("act", out.detach().clone()))
        norm.register_full_backward_hook(lambda mod, gi, go: seen.__setitem__("grad", go[0].detach().clone()))
    x = torch.randn(batch, 1, 24, 16, 16, device=DEV)
    y = torch.randint(0, 2, (batch,), device=DEV)

    def run(flag):
        Fn.CLS_LAST = flag
        torch.manual_seed(77)   # the dropout seeds come from torch's CPU generator: same masks in both runs
        m.zero_grad(set_to_none=True)
        seen.clear()
        logits = m(x)
        torch.nn.functional.cross_entropy(logits, y).backward()
        torch.cuda.synchronize()
        return logits.detach().clone(), {k: p.grad.detach().clone() for k, p in m.named_parameters()}, dict(seen)

    try:
        got, want = run(True), run(False)
    finally:
        Fn.CLS_LAST = True
    assert got[0].shape == want[0].shape == (batch, 2)
    assert rel(got[0], want[0]) < 1e-5                      # same kernels per row; nothing but reduction order may differ
    for k in want[1]:
        assert rel(got[1][k], want[1][k]) < 2e-3, k         # split-K / column-sum atomics order, bf16 side-car rounding
    if hooked:
        assert torch.equal(got[2]["act"], want[2]["act"]) and got[2]["act"].shape[1] == 13
        assert rel(got[2]["grad"], want[2]["grad"]) < 2e-3


def test_flat_adamw_matches_torch_adamw():
    """trainer.FlatAdamW (one kernel over flat params / grads / moments + bf16 weight-copy refresh) against
    torch.optim.AdamW on the same model, data and hyper-parameters, several steps."""
    import copy
    from neurovit_b200.trainer import DataParallelTrainer, FlatAdamW
    torch.manual_seed(31)
    ctor = dict(image_size=16, image_patch_size=8, frames=16, frame_patch_size=8, num_classes=2, dim=128, depth=2,
                heads=2, mlp_dim=256, channels=1, dim_head=64)
    ma = ViT(**ctor).to(DEV)
    mb = copy.deepcopy(ma)
    x = torch.randn(6, 1, 16, 16, 16, device=DEV)
    y = torch.randint(0, 2, (6,), device=DEV)
    ta = DataParallelTrainer(ma, lr=1e-3, weight_decay=0.05)
    assert isinstance(ta.optimizer, FlatAdamW)
    tb = DataParallelTrainer(mb, optimizer=torch.optim.AdamW(mb.parameters(), lr=1e-3, weight_decay=0.05))
    for _ in range(4):
        la, lb = ta.step(x, y), tb.step(x, y)
    torch.cuda.synchronize()
    assert abs(la.item() - lb.item()) < 2e-3
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert (pa - pb).abs().max().item() < 1e-3 + 2e-3 * pb.abs().max().item(), k   # atomics-order noise only
    # state_dict still has the reference's 78 tensors and round-trips
    sd = ma.state_dict()
    mc = ViT(**ctor).to(DEV)
    mc.load_state_dict(sd, strict=True)
    ma.eval(); mc.eval()
    assert rel(mc(x), ma(x)) < 1e-3


def test_graphed_step_matches_eager_and_redraws_dropout():
    """DataParallelTrainer(graph=True): one CUDA graph per step. Without dropout the trajectory must follow the eager
    trainer; with dropout every replay must draw new masks (device-side epoch) — the losses on the same batch differ."""
    import copy
    from neurovit_b200.trainer import DataParallelTrainer
    torch.manual_seed(41)
    ctor = dict(image_size=16, image_patch_size=8, frames=16, frame_patch_size=8, num_classes=2, dim=128, depth=2,
                heads=2, mlp_dim=256, channels=1, dim_head=64)
    ma = ViT(**ctor).to(DEV)
    mb = copy.deepcopy(ma)
    xs = [torch.randn(8, 1, 16, 16, 16, device=DEV) for _ in range(3)]
    ys = [torch.randint(0, 2, (8,), device=DEV) for _ in range(3)]
    ta = DataParallelTrainer(ma, lr=1e-3, graph=True)
    tb = DataParallelTrainer(mb, lr=1e-3, graph=False)
    for i in range(5):
        la = ta.step(xs[i % 3], ys[i % 3]).item()
        lb = tb.step(xs[i % 3], ys[i % 3]).item()
        assert abs(la - lb) < 5e-3, (i, la, lb)
    assert ta.optimizer.t == 5 and tb.optimizer.t == 5   # the capture's warm-up steps were rolled back
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert (pa - pb).abs().max().item() < 2e-3 + 2e-3 * pb.abs().max().item(), k
    # dropout: same batch, consecutive replays -> different masks -> different losses
    mc = ViT(**{**ctor, "dropout": 0.3, "emb_dropout": 0.3}).to(DEV)
    tc = DataParallelTrainer(mc, lr=0.0, weight_decay=0.0, graph=True)
    losses = [tc.step(xs[0], ys[0]).item() for _ in range(4)]
    assert len({round(v, 6) for v in losses}) == 4, losses


def test_predrawn_and_inline_dropout_bits_give_identical_results():
    """Keep bits drawn ahead on the side stream (functional.MaskGen / nv_dropout_bits) are the same Philox bits the
    kernels draw inline: with equal seeds the two paths must agree bit for bit in forward and closely in backward."""
    from neurovit_b200 import functional as Fn
    ctor = dict(image_size=16, image_patch_size=8, frames=24, frame_patch_size=8, num_classes=2, dim=64, depth=2,
                heads=2, mlp_dim=128, channels=1, dim_head=64, dropout=0.25, emb_dropout=0.1)
    torch.manual_seed(5)
    m = ViT(**ctor).to(DEV).train()
    x = torch.randn(6, 1, 24, 16, 16, device=DEV)
    y = torch.tensor([0, 1, 1, 0, 1, 0], device=DEV)
    saved_sites = set(Fn.MASKS.sites)
    results = []
    try:
        for sites in ({"attn", "gemm"}, set()):
            Fn.MASKS.sites = set(sites)
            torch.manual_seed(77)       # same dropout seeds in both runs
            m.zero_grad(set_to_none=True)
            logits = m(x)
            torch.nn.functional.cross_entropy(logits, y).backward()
            torch.cuda.synchronize()
            results.append((logits.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()}))
    finally:
        Fn.MASKS.sites = saved_sites
    assert torch.equal(results[0][0], results[1][0])
    for k in results[0][1]:
        assert rel(results[0][1][k], results[1][1][k]) < 1e-4, k   # split-K / column-sum atomics reorder


def test_trainer_load_state_dict_refreshes_bf16_copies_and_recaptures():
    """A checkpoint loaded into a live graphed trainer (trainer.load_state_dict): the next step must run on the loaded
    weights — bf16 operand copies re-cast, the captured graph dropped, optimizer hyper-parameters and moments
    restored — i.e. continue exactly like a fresh trainer built from that checkpoint."""
    import copy
    from neurovit_b200.trainer import DataParallelTrainer
    torch.manual_seed(51)
    ctor = dict(image_size=16, image_patch_size=8, frames=16, frame_patch_size=8, num_classes=2, dim=128, depth=2,
                heads=2, mlp_dim=256, channels=1, dim_head=64)
    x = torch.randn(8, 1, 16, 16, 16, device=DEV)
    y = torch.randint(0, 2, (8,), device=DEV)
    # trainer A: two steps, then checkpoint
    ma = ViT(**ctor).to(DEV)
    ta = DataParallelTrainer(ma, lr=1e-3, graph=True)
    for _ in range(2):
        ta.step(x, y)
    torch.cuda.synchronize()
    model_sd = copy.deepcopy(ma.state_dict())
    opt_sd = copy.deepcopy(ta.optimizer.state_dict())
    la = ta.step(x, y).item()                      # what the third step gives from that checkpoint
    # trainer B: different weights and hyper-parameters, already captured; then the checkpoint is loaded into it
    torch.manual_seed(52)
    mb = ViT(**ctor).to(DEV)
    tb = DataParallelTrainer(mb, lr=5e-2, weight_decay=0.3, graph=True)
    tb.step(x, y)
    tb.load_state_dict(model_sd, opt_sd)
    assert tb.optimizer.lr == 1e-3 and tb.optimizer.t == 2
    lb = tb.step(x, y).item()
    torch.cuda.synchronize()
    assert abs(la - lb) < 2e-3, (la, lb)
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert (pa - pb).abs().max().item() < 1e-3 + 2e-3 * pa.abs().max().item(), k
