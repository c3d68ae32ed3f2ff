"""Small instances of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): a tiny ViT training
step with dropout through the drop-in module (patch gather + LN fold, tcgen05 GEMMs with every epilogue, LayerNorm
fwd / bwd, flash attention fwd / dQ / dK/dV, the cls-only last layer, head, mask generator, fused AdamW through the
trainer) and the 4D de-interleave. Usage (GPU box):
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py
(compute-sanitizer is closed on the gpurun pool of this build — the call answers "closed on this pool" — so the script has
only been run plain there; out-of-range accesses are guarded by the NaN-prefilled outputs and exact-shape assertions of
tests/test_gpu_kernels.py instead.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200 import functional as Fn, ops  # noqa: E402
from neurovit_b200.trainer import DataParallelTrainer  # noqa: E402
from neurovit_b200.vit_3d import ViT  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
m = ViT(image_size=16, image_patch_size=8, frames=24, frame_patch_size=8, num_classes=2, dim=128, depth=2, heads=2,
        mlp_dim=256, channels=1, dim_head=64, dropout=0.1, emb_dropout=0.1).to(dev).train()
x = torch.randn(6, 1, 24, 16, 16, device=dev)
y = torch.randint(0, 2, (6,), device=dev)
for cls_last in (True, False):
    Fn.CLS_LAST = cls_last
    m.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(m(x), y)
    loss.backward()
    torch.cuda.synchronize()
    print(f"ViT step (cls-only last layer {cls_last}): loss {loss.item():.4f}", flush=True)
Fn.CLS_LAST = True
tr = DataParallelTrainer(m, lr=1e-3, weight_decay=0.01)      # flat buffers, gradient sinks, fused AdamW (eager launches)
for _ in range(2):
    l = tr.step(x, y)
torch.cuda.synchronize()
print(f"trainer steps: loss {l.item():.4f}", flush=True)
# attention kernels on a ragged shape with dropout, dense and cls-only
B, N, H, hd = 2, 161, 2, 64
qkv = torch.randn(B * N, 3 * H * hd, device=dev).to(torch.bfloat16)
o = torch.empty(B * N, H * hd, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device=dev)
mask = torch.zeros(B * H, N, (N + 31) // 32, device=dev, dtype=torch.int32)
kw = dict(B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=0.2)
ops.attention_fwd(qkv, o, lse, seed=3, drop_mask=mask, **kw)
dO = torch.randn(B * N, H * hd, device=dev).to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
ops.attention_bwd(qkv, o, dO, lse, torch.empty(B * H * N, device=dev), dqkv, drop_mask=mask, **kw)
o_c = torch.empty(B, H * hd, device=dev, dtype=torch.bfloat16)
ops.attention_cls_fwd(qkv, o_c, lse, seed=3, drop_mask=mask, **kw)
ops.attention_cls_bwd(qkv, o_c, dO[:B].contiguous(), lse, dqkv, drop_mask=mask, o_bs=o_c.stride(0), **kw)
# 4D de-interleave
f = torch.randn(1, 8, 8, 4, 7, device=dev)
v = Fn.fmri_to_volumes(f, zscore=True)
torch.cuda.synchronize()
print("attention / 4D kernels done", flush=True)
