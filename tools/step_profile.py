"""Per-kernel device time of the benchmark training step measured with torch.profiler (CUPTI activity
records: real, overlapped execution — unlike ncu's serialised cold-cache replays), plus GPU busy/idle
time over the profiled steps. Usage: python tools/step_profile.py [--steps 3] [--batch 64] [--out FILE]"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from neurovit_b200.trainer import DataParallelTrainer  # noqa: E402
from neurovit_b200.vit_3d import ViT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--config", default="cfgA")
ap.add_argument("--out", default=None)
ap.add_argument("--dropout", type=float, default=bench.DROPOUT)
args = ap.parse_args()
bench.DROPOUT = args.dropout

cfg = bench.CONFIGS[args.config] if hasattr(bench, "CONFIGS") else bench.CFG
dev = torch.device("cuda", 0)
torch.manual_seed(42)


class Enc(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.vit3d = ViT(**bench.vit_ctor(cfg))

    def forward(self, x):
        return self.vit3d(x.permute(0, 3, 1, 2).unsqueeze(1))


enc = Enc().to(dev).train()
tr = DataParallelTrainer(enc, lr=1e-4, weight_decay=0.01)
H, W, D = cfg["vol"]
xs = [torch.randn(args.batch, H, W, D, device=dev) for _ in range(3)]
ys = [torch.randint(0, 2, (args.batch,), device=dev) for _ in range(3)]
for i in range(4):
    tr.step(xs[i % 3], ys[i % 3])
torch.cuda.synchronize()

import time  # noqa: E402

e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(10):
    tr.step(xs[i % 3], ys[i % 3])
e1.record()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"un-profiled: {e0.elapsed_time(e1) / 10:.3f} ms/step on the device, host issue time {t_issue * 100:.3f} ms/step "
      f"(host-bound if the two are equal)")

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        tr.step(xs[i % 3], ys[i % 3])
    e1.record()
    torch.cuda.synchronize()
wall_ms = e0.elapsed_time(e1)

tot = collections.defaultdict(float)
cnt = collections.Counter()
spans = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name
        dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        tot[name] += dur
        cnt[name] += 1
        spans.append((ev.time_range.start, ev.time_range.end))
spans.sort()
busy, cur_s, cur_e = 0.0, None, None
for s, e in spans:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
if cur_e is not None:
    busy += cur_e - cur_s
lines = []
total = sum(tot.values())
lines.append(f"steps {args.steps}  wall (CUDA events, under the profiler) {wall_ms / args.steps:.3f} ms/step  "
             f"sum of kernel time {total / 1e3 / args.steps:.3f} ms/step  GPU busy (union) {busy / 1e3 / args.steps:.3f} ms/step")
lines.append("| kernel | launches/step | us/step | share % | avg us |")
lines.append("|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    lines.append(f"| `{k[:90]}` | {cnt[k] / args.steps:.1f} | {v / args.steps:.1f} | {100 * v / total:.1f} | {v / cnt[k]:.1f} |")
text = "\n".join(lines)
print(text)
if args.out:
    with open(args.out, "w") as f:
        f.write(text + "\n")
