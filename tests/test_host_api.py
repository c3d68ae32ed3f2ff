"""Host-side mirror of the reference interface (no GPU): constructor checks and messages (vit_3d.py:83-89), module
tree / state_dict keys as saved by the unmodified reference (tests/golden/*.npz, oracle/gen_golden.py), argument
validation in forward, and the rule that nothing computes without the CUDA library's device (no CPU fallback)."""
import os

import numpy as np
import pytest
import torch

from neurovit_b200._lib import NeuroViTLibraryError
from neurovit_b200.vit_3d import ViT

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _vit(**kw):
    args = dict(image_size=16, image_patch_size=8, frames=16, frame_patch_size=8, num_classes=2, dim=64, depth=2,
                heads=2, mlp_dim=128, channels=1, dim_head=64, dropout=0.0, emb_dropout=0.0)   # = vit3d_small.npz
    args.update(kw)
    return ViT(**args)


def test_state_dict_keys_and_shapes_match_the_reference_checkpoint():
    g = np.load(os.path.join(GOLD, "vit3d_small.npz"))
    ref = {k[3:]: g[k].shape for k in g.files if k.startswith("sd.")}
    sd = _vit().state_dict()
    assert list(sd.keys()) == list(ref.keys())          # same names in the same order: load_state_dict both ways
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref[k]), k
    m = _vit()
    m.load_state_dict({k: torch.from_numpy(g["sd." + k]) for k in ref}, strict=True)


def test_module_tree_keeps_the_hook_points_of_the_reference():
    m = _vit()
    # explainability hooks attach to transformer.layers[i][0].attend (gradcam3DViT_fmris.py:25-26)
    attn, ff = m.transformer.layers[-1]
    assert isinstance(attn.attend, torch.nn.Softmax) and isinstance(attn.dropout, torch.nn.Dropout)
    assert isinstance(attn.to_qkv, torch.nn.Linear) and attn.to_qkv.bias is None
    assert isinstance(m.to_patch_embedding[2], torch.nn.Linear) and isinstance(m.mlp_head[1], torch.nn.Linear)
    assert m.pos_embedding.shape == (1, 2 * 2 * 2 + 1, 64) and m.cls_token.shape == (1, 1, 64)


@pytest.mark.parametrize("kw,msg", [
    (dict(image_size=20), "Image dimensions must be divisible by the patch size"),
    (dict(frames=12), "Frames must be divisible by frame patch size"),
    (dict(pool="max"), "pool type must be either cls"),
])
def test_constructor_rejects_what_the_reference_rejects(kw, msg):
    with pytest.raises(AssertionError, match=msg):
        _vit(**kw)


def test_forward_validates_shapes_before_touching_the_device():
    m = _vit()
    with pytest.raises(ValueError, match="5-D"):
        m(torch.zeros(2, 16, 16, 16))
    with pytest.raises(ValueError, match="not divisible by the patch size"):
        m(torch.zeros(2, 1, 16, 16, 12))
    with pytest.raises(ValueError, match="channels differ"):
        m(torch.zeros(2, 3, 16, 16, 16))
    assert m(torch.zeros(0, 1, 16, 16, 16)).shape == (0, 2)      # empty batch: nothing to launch


def test_cpu_tensors_fail_loudly_instead_of_falling_back():
    m = _vit()
    with pytest.raises(NeuroViTLibraryError, match="no CPU fallback"):
        m(torch.zeros(2, 1, 16, 16, 16))


# ------------------------------------------------------------------------------------ NeuroEncoder (config dict)
def _cfg(tmp, **kw):
    cfg = {"DEVICE": "cpu", "TRAINING_DIM": 3, "TRAINING_DROPOUT": 0.0, "TRAINING_VIT_INPUT_SIZE": 16,
           "GRADCAM_CUBE_SIZE": 8, "TRAINING_VIT_PATCH_SIZE": 8, "DATASET_NAME": "adni", "GLOBAL_BASE_PATH": str(tmp),
           "BEST_MODEL_PATH": "best.pth", "GRADCAM_THRESHOLD": 10, "GRADCAM_SLICE_DIM": 0, "GRADCAM_SLICE_IDX": 0}
    cfg.update(kw)
    return cfg


def test_neuroencoder_takes_the_reference_config_and_registers_the_gradcam_hooks(tmp_path):
    from neurovit_b200.NeuroEncoder import NeuroEncoder
    m = NeuroEncoder(_cfg(tmp_path))
    assert m.forward_handle is not None and m.backward_handle is not None       # NeuroEncoder.py:70-82
    keys = set(m.state_dict().keys())
    g = np.load(os.path.join(GOLD, "neuro3d.npz"))       # holds gradients of a sample of the reference's parameters
    assert len(keys) == 78 and {k[5:] for k in g.files if k.startswith("grad.")} <= keys
    off = NeuroEncoder(_cfg(tmp_path, GRADCAM_CAPTURE="off"))
    assert off.forward_handle is None and off.backward_handle is None
    with pytest.raises(ValueError, match="GRADCAM_CAPTURE"):
        NeuroEncoder(_cfg(tmp_path, GRADCAM_CAPTURE="disk"))


def test_neuroencoder_rejects_an_unknown_training_dim(tmp_path):
    from neurovit_b200.NeuroEncoder import NeuroEncoder
    m = NeuroEncoder(_cfg(tmp_path))
    m.config["TRAINING_DIM"] = 5
    with pytest.raises(ValueError, match="TRAINING_DIM must be 3 or 4"):
        m(torch.zeros(1, 16, 16, 16))


def test_zero_arena_hands_out_disjoint_aligned_regions():
    """functional.ZEROS: small fp32 accumulators carved from one zero-filled chunk — every region handed out once,
    128-byte slots, a fresh chunk after reset() and when the request does not fit."""
    from neurovit_b200 import functional as Fn
    Fn.ZEROS.reset()
    a = Fn.zeros_f32(10, "cpu")
    b = Fn.zeros_f32(1024, "cpu")
    c = Fn.zeros_f32(3, "cpu")
    assert a.shape == (10,) and b.shape == (1024,) and c.shape == (3,)
    assert float(a.abs().sum() + b.abs().sum() + c.abs().sum()) == 0.0
    assert b.data_ptr() - a.data_ptr() == 128 and c.data_ptr() - b.data_ptr() == 4096   # 32-float slots, in order
    a.fill_(1.0)
    b.fill_(2.0)
    assert float(c.sum()) == 0.0 and float(Fn.zeros_f32(10, "cpu").sum()) == 0.0          # neighbours untouched
    big = Fn.zeros_f32(Fn.ZEROS.CHUNK, "cpu")                                             # larger than a chunk share: own tensor
    assert big.numel() == Fn.ZEROS.CHUNK and float(big.sum()) == 0.0
    Fn.ZEROS.reset()
    d = Fn.zeros_f32(10, "cpu")
    assert d.data_ptr() != a.data_ptr() and float(d.sum()) == 0.0                         # a new chunk after a step boundary
    # exhaust a chunk: the next request starts a new one instead of wrapping around
    n = Fn.ZEROS.CHUNK // 4
    ptrs = {Fn.zeros_f32(n, "cpu").data_ptr() for _ in range(6)}
    assert len(ptrs) == 6


def test_sparse_grad_zeros_is_cached_and_recleared_after_an_in_place_write():
    from neurovit_b200 import functional as Fn
    t = Fn.sparse_grad_zeros("test", (4, 3, 8), torch.float32, "cpu")
    assert Fn.sparse_grad_zeros("test", (4, 3, 8), torch.float32, "cpu") is t           # same buffer, no new fill
    t.add_(1.0)                                                                           # a torch-side write bumps the version
    u = Fn.sparse_grad_zeros("test", (4, 3, 8), torch.float32, "cpu")
    assert u is not t and float(u.abs().sum()) == 0.0


def test_cls_only_last_layer_is_taken_only_when_nothing_could_observe_the_difference():
    """Transformer._cls_last_ok (host logic, no device): with pool='cls' the last layer may run for token 0 alone only in
    bf16 mode and only when no user hook sits inside it other than on Attention.norm (Grad-CAM, NeuroEncoder.py:47: that
    LayerNorm still sees every token); a hook anywhere else would observe [B, 1, D] instead of [B, N, D]."""
    from neurovit_b200 import functional as Fn
    x = torch.zeros(2, 9, 64)
    m = _vit()
    tr = m.transformer
    attn, ff = tr.layers[-1]
    assert tr._cls_last_ok(x)
    assert not tr._cls_last_ok(torch.zeros(2, 1, 64))             # a single token: nothing to skip
    h = attn.norm.register_forward_hook(lambda *a: None)
    assert tr._cls_last_ok(x)                                      # the Grad-CAM hook point is allowed
    h.remove()
    for mod in (attn, attn.to_qkv, attn.to_out[0], ff, ff.net[0], ff.net[4]):
        h = mod.register_forward_hook(lambda *a: None)
        assert not tr._cls_last_ok(x), type(mod).__name__
        h.remove()
        h = mod.register_full_backward_hook(lambda *a: None)
        assert not tr._cls_last_ok(x), type(mod).__name__
        h.remove()
    assert tr._cls_last_ok(x)
    hook_first = tr.layers[0][1].register_forward_hook(lambda *a: None)   # hooks in OTHER layers do not matter
    assert tr._cls_last_ok(x)
    hook_first.remove()
    m.set_precision("fp32")
    assert not tr._cls_last_ok(x)                                  # the verification mode keeps the dense layer
    m.set_precision("bf16")
    try:
        Fn.CLS_LAST = False
        assert not tr._cls_last_ok(x)
    finally:
        Fn.CLS_LAST = True
    assert not _vit(pool="mean").pool == "cls"                     # ViT.forward asks for it with pool == 'cls' only
