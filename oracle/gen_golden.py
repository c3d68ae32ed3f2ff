"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only at /root/reference) on seeded
inputs. Run in the build container only (the reference does not travel to the GPU box):

    python oracle/gen_golden.py

Fixtures (all fp32 unless noted):
  patch_index.npz   int64 arange volumes pushed through the reference's einops Rearrange, both input layouts
  vit3d_small.npz   ViT(dim=64, depth=2, heads=2, dim_head=64, mlp=128, 16x16x16 / patch 8, 1 channel):
                    state_dict, input, labels, logits, CE loss, every parameter gradient
  vit3d_p4.npz      ViT(channels=2, 8x8x12 / patch 4, pool='mean') — permuted-K layout + mean pooling
  neuro3d.npz       NeuroEncoder 3D through ViT3DEncoder's permuted view (grid 16, patch 8; dims hard-coded 1024/6/8/2048)
  neuro4d.npz       NeuroEncoder 4D (T=6) incl. TemporalTransformer/ProjectionHead gradients
  neuro3d_cam.npz   NeuroEncoder.get_attention_map on a 32^3 volume: the Grad-CAM map, the class and the raw token scores
"""
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

REF = os.environ.get("NEUROVIT_REF", "/root/reference")
sys.path.insert(0, REF)
from src.models.vit_3d import ViT  # noqa: E402
from src.models.NeuroEncoder import NeuroEncoder  # noqa: E402
from einops import rearrange  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
warnings.filterwarnings("ignore")


def to_np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def gen_patch_index():
    arrs = {}
    for tag, (B, C, Fr, H, W, p) in {"c1_p8": (2, 1, 16, 16, 8, 8), "c2_p4": (1, 2, 8, 12, 4, 4),
                                      "c1_p9": (1, 1, 18, 18, 9, 9)}.items():
        v = torch.arange(B * C * Fr * H * W, dtype=torch.int64).reshape(B, C, Fr, H, W)
        arrs[tag + "_contig"] = rearrange(v, 'b c (f pf) (h p1) (w p2) -> b (f h w) (p1 p2 pf c)', p1=p, p2=p, pf=p).numpy()
        arrs[tag + "_shape"] = np.array([B, C, Fr, H, W, p])
        if C == 1:  # the ViT3DEncoder view: x[B,H,W,D] -> [B,1,D,H,W] (NeuroEncoder.py:200-202)
            x = torch.arange(B * H * W * Fr, dtype=torch.int64).reshape(B, H, W, Fr)
            view = x.permute(0, 3, 1, 2).unsqueeze(1)
            arrs[tag + "_view"] = rearrange(view, 'b c (f pf) (h p1) (w p2) -> b (f h w) (p1 p2 pf c)', p1=p, p2=p, pf=p).numpy()
    save("patch_index.npz", **arrs)


def run_vit(name, seed, B, ctor, shape):
    torch.manual_seed(seed)
    m = ViT(**ctor).eval()
    # non-trivial LayerNorm affine params / biases so their gradients are exercised
    with torch.no_grad():
        for k, p in m.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    video = torch.randn(B, *shape)
    labels = torch.randint(0, ctor["num_classes"], (B,))
    logits = m(video)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    arrs = {"sd." + k: v for k, v in to_np(m.state_dict()).items()}
    arrs.update({"grad." + k: p.grad.numpy() for k, p in m.named_parameters()})
    arrs.update(video=video.numpy(), labels=labels.numpy(), logits=logits.detach().numpy(), loss=loss.detach().numpy())
    save(name, **arrs)


def neuro_config(tmp, dim):
    return {"DEVICE": "cpu", "TRAINING_DIM": dim, "TRAINING_DROPOUT": 0.0, "TRAINING_VIT_INPUT_SIZE": 16,
            "GRADCAM_CUBE_SIZE": 8, "TRAINING_VIT_PATCH_SIZE": 8, "DATASET_NAME": "adni", "GLOBAL_BASE_PATH": tmp,
            "BEST_MODEL_PATH": "best.pth", "GRADCAM_THRESHOLD": 10, "GRADCAM_SLICE_DIM": 0, "GRADCAM_SLICE_IDX": 0}


def gen_neuro():
    """The 38 M-parameter state_dict is NOT stored: it is regenerated from the seed by the consumer, and the
    fixture pins a checksum of it plus logits / selected gradients."""
    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(1234)
        m3 = NeuroEncoder(neuro_config(tmp, 3)).eval()
        x = torch.randn(2, 16, 16, 16)
        labels = torch.tensor([0, 1])
        logits = m3(x)
        loss = torch.nn.functional.cross_entropy(logits, labels)
        loss.backward()
        sd = m3.state_dict()
        torch.save(sd, os.path.join(tmp, "best.pth"))
        keep = ["volume_encoder.vit3d.mlp_head.1.weight", "volume_encoder.vit3d.mlp_head.0.weight",
                "volume_encoder.vit3d.cls_token", "volume_encoder.vit3d.transformer.layers.5.0.norm.weight",
                "volume_encoder.vit3d.transformer.layers.0.0.to_out.0.bias",
                "volume_encoder.vit3d.to_patch_embedding.1.weight", "volume_encoder.vit3d.to_patch_embedding.3.bias"]
        grads = dict(m3.named_parameters())
        arrs = {"grad." + k: grads[k].grad.numpy() for k in keep}
        arrs.update(x=x.numpy(), labels=labels.numpy(), logits=logits.detach().numpy(), loss=loss.detach().numpy(),
                    sd_checksum=np.array([float(sum(v.double().sum() for v in sd.values()))]),
                    gradnorm=np.array([float(sum((p.grad.double() ** 2).sum() for p in m3.parameters()) ** 0.5)]),
                    act_hook=m3.activations.numpy()[:, :3, :8], grad_hook=m3.gradients.numpy()[:, :3, :8])
        save("neuro3d.npz", **arrs)

        torch.manual_seed(4321)
        m4 = NeuroEncoder(neuro_config(tmp, 4)).eval()
        x4 = torch.randn(2, 16, 16, 16, 6)
        out = m4(x4)
        loss4 = torch.nn.functional.cross_entropy(out, labels)
        loss4.backward()
        arrs = {"sd." + k: v for k, v in to_np(m4.state_dict()).items() if not k.startswith("volume_encoder.")}
        arrs.update({"grad." + k: p.grad.numpy() for k, p in m4.named_parameters() if p.grad is not None})
        arrs.update(x=x4.numpy(), labels=labels.numpy(), out=out.detach().numpy(), loss=loss4.detach().numpy())
        save("neuro4d.npz", **arrs)


def gen_cam():
    """Grad-CAM map of the unmodified reference (NeuroEncoder.get_attention_map, NeuroEncoder.py:84-133) on a 32^3
    volume (4x4x4 patch tokens), same seeded weights as neuro3d.npz; the consumer regenerates them from the seed."""
    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(1234)
        cfg = {**neuro_config(tmp, 3), "TRAINING_VIT_INPUT_SIZE": 32, "GRADCAM_THRESHOLD": 25}
        m = NeuroEncoder(cfg).eval()
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle.vit3d_oracle import perturb_for_cam, gradcam_from_hooks
        perturb_for_cam(m.volume_encoder.vit3d)   # the raw map is round-off at initialisation: see its docstring
        g = torch.Generator().manual_seed(99)
        x = torch.randn(1, 32, 32, 32, generator=g)
        cam, cls = m.get_attention_map(x)
        raw = (m.gradients.mean(dim=2, keepdim=True) * m.activations).sum(dim=2)[:, 1:]
        again = gradcam_from_hooks(m.gradients, m.activations, 32, 8, 25)
        assert torch.equal(again, cam), "oracle restatement of the Grad-CAM post-processing differs from the reference"
        save("neuro3d_cam.npz", x=x.numpy(), cam=cam.numpy(), cls=cls.numpy(), raw_cam=raw.detach().numpy(),
             sd_checksum=np.array([float(sum(v.double().sum() for v in m.state_dict().values()))]))


if __name__ == "__main__":
    if "--cam-only" in sys.argv:
        gen_cam()
        sys.exit(0)
    gen_patch_index()
    run_vit("vit3d_small.npz", 42, 3, dict(image_size=16, image_patch_size=8, frames=16, frame_patch_size=8,
                                           num_classes=2, dim=64, depth=2, heads=2, mlp_dim=128, channels=1,
                                           dim_head=64), (1, 16, 16, 16))
    run_vit("vit3d_p4.npz", 7, 2, dict(image_size=(8, 12), image_patch_size=4, frames=4, frame_patch_size=4,
                                       num_classes=3, dim=64, depth=1, heads=2, mlp_dim=64, channels=2, dim_head=64,
                                       pool="mean"), (2, 4, 8, 12))
    gen_neuro()
    gen_cam()
