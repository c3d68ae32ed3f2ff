"""Where does bf16 error come from? Runs the full-size ViT3D (cfgA dims) in fp32 verification mode and in
bf16 mode on the same weights/inputs (GPU) and prints, per parameter, max-norm and L2 relative gradient
error of bf16 vs fp32, in backward order. Usage: python tools/precision_probe.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200.vit_3d import ViT  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(11)
ctor = dict(image_size=64, image_patch_size=8, frames=48, frame_patch_size=8, num_classes=2, dim=1024, depth=6,
            heads=8, mlp_dim=2048, channels=1, dim_head=64)
m = ViT(**ctor)
with torch.no_grad():
    for p in m.parameters():
        if p.dim() == 1:
            p.add_(0.1 * torch.randn_like(p))
m = m.cuda().eval()
video = torch.randn(B, 1, 48, 64, 64, device="cuda")
labels = torch.randint(0, 2, (B,), device="cuda")
res = {}
for mode in ("fp32", "bf16"):
    m.set_precision(mode)
    m.zero_grad(set_to_none=True)
    logits = m(video)
    torch.nn.functional.cross_entropy(logits, labels).backward()
    res[mode] = (logits.detach().double(), {k: p.grad.detach().double() for k, p in m.named_parameters()})
lf, gf = res["fp32"]
lb, gb = res["bf16"]
print("logits max-rel", ((lb - lf).abs().max() / lf.abs().max()).item())
worst = 0.0
for k in reversed(list(gf)):
    a, b = gb[k], gf[k]
    mx = ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
    l2 = ((a - b).norm() / (b.norm() + 1e-30)).item()
    worst = max(worst, mx)
    print(f"{mx:9.3e} {l2:9.3e}  {k}")
print("worst max-rel", worst)
