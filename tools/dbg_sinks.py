"""Which parameters report their gradient how many times per backward (bucket countdown of trainer.FlatGradBuckets)?"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200.trainer import DataParallelTrainer  # noqa: E402
from neurovit_b200.vit_3d import ViT  # noqa: E402

dev = "cuda"
ctor = dict(image_size=16, image_patch_size=8, frames=24, frame_patch_size=8, num_classes=2, dim=128, depth=2,
            heads=2, mlp_dim=256, channels=1, dim_head=64, dropout=float(os.environ.get("P", "0")), emb_dropout=0.0)
torch.manual_seed(0)
m = ViT(**ctor).to(dev).train()
tr = DataParallelTrainer(m, lr=0.0, weight_decay=0.0, bucket_mb=0)
names = {id(p): k for k, p in m.named_parameters()}
b = tr.buckets
b.world = 2
b.hooks = []
counts = collections.Counter()
order = []
def _rec(p):
    counts.update([names[id(p)]])
    order.append(names[id(p)])


b._on_grad = _rec
for h in b.hooks:
    h.remove()
b.hooks = [p.register_post_accumulate_grad_hook(b._on_hook) for p in b.params]
b.finish = lambda: None
B = int(os.environ.get("B", "6"))
x = torch.randn(B, 1, 24, 16, 16, device=dev)
y = torch.randint(0, 2, (B,), device=dev)
tr.step(x, y)
torch.cuda.synchronize()
for k, p in m.named_parameters():
    if counts[k] != 1:
        print(f"{k}: reported {counts[k]} times")
print("total params", len(names), "reports", sum(counts.values()))
print("order:", order)
