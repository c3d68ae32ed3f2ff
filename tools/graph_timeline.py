"""Timeline of the REPLAYED training-step graph (CUPTI activity records through torch.profiler): per-kernel time,
the idle gaps between consecutive kernels on the main stream, and how much of the side-stream work (dropout bit
generation, NCCL) runs under main-stream kernels. Answers "where does step time minus kernel time go".
Usage: python tools/graph_timeline.py [--steps 3] [--dropout 0.1] [--out FILE]"""
import argparse
import collections
import json
import os
import sys
import tempfile

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
ap.add_argument("--dropout", type=float, default=bench.DROPOUT)
ap.add_argument("--out", default=None)
ap.add_argument("--dump", default=None, help="CSV of every kernel of the profiled steps: start us, duration us, stream, name")
args = ap.parse_args()
bench.DROPOUT = args.dropout
cfg = bench.CONFIGS[args.config]
dev = torch.device("cuda", 0)
torch.manual_seed(42)


class Enc(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.vit3d = ViT(**bench.vit_ctor(cfg))

    def forward(self, x):
        return self.vit3d(x.permute(0, 3, 1, 2).unsqueeze(1))


enc = Enc().to(dev).train()
tr = DataParallelTrainer(enc, lr=1e-4, weight_decay=0.01, graph=True)
H, W, D = cfg["vol"]
xs = [torch.randn(args.batch, H, W, D, device=dev) for _ in range(3)]
ys = [torch.randint(0, 2, (args.batch,), device=dev) for _ in range(3)]
for i in range(6):
    tr.step(xs[i % 3], ys[i % 3])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    tr.step(xs[i % 3], ys[i % 3])
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / 10

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(args.steps):
        tr.step(xs[i % 3], ys[i % 3])
    torch.cuda.synchronize()
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "trace.json")
    prof.export_chrome_trace(path)
    trace = json.load(open(path))
ks = [e for e in trace["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ks.sort(key=lambda e: e["ts"])
by_stream = collections.Counter(e["args"].get("stream") for e in ks)
main = by_stream.most_common(1)[0][0]
mk = [e for e in ks if e["args"].get("stream") == main]
sk = [e for e in ks if e["args"].get("stream") != main]
t_first, t_last = ks[0]["ts"], max(e["ts"] + e["dur"] for e in ks)
span = (t_last - t_first) / args.steps
gaps = []
for a, b in zip(mk, mk[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    gaps.append((g, a["name"][:60], b["name"][:60]))
step_gap = sorted(gaps, key=lambda x: -x[0])[:args.steps - 1]   # the gaps between replays themselves
inner = [g for g in gaps if g not in step_gap]
pos = [g for g in inner if g[0] > 0]
lines = [f"un-profiled graph replay: {plain_ms:.3f} ms/step; profiled span {span / 1e3:.3f} ms/step",
         f"main stream {main}: {len(mk) / args.steps:.0f} kernels/step, kernel time {sum(e['dur'] for e in mk) / args.steps / 1e3:.3f} ms/step, "
         f"positive gaps {sum(g[0] for g in pos) / args.steps / 1e3:.3f} ms/step over {len(pos) / args.steps:.0f} boundaries "
         f"(mean {sum(g[0] for g in pos) / max(len(pos), 1):.2f} us); overlapping (negative) {sum(g[0] for g in inner if g[0] <= 0) / args.steps / 1e3:.3f} ms",
         f"side streams: {len(sk) / args.steps:.0f} kernels/step, {sum(e['dur'] for e in sk) / args.steps / 1e3:.3f} ms/step"]
lines.append("largest gaps (us): " + "; ".join(f"{g:.0f} [{a[:28]} -> {b[:28]}]" for g, a, b in sorted(gaps, key=lambda x: -x[0])[:6]))
# gap by following-kernel name
bynext = collections.defaultdict(lambda: [0.0, 0])
for g, a, b in inner:
    bynext[b][0] += g
    bynext[b][1] += 1
lines.append("| gap before kernel | count/step | total us/step | mean us |")
lines.append("|---|---|---|---|")
for k, (t, n) in sorted(bynext.items(), key=lambda kv: -kv[1][0])[:25]:
    lines.append(f"| `{k}` | {n / args.steps:.1f} | {t / args.steps:.1f} | {t / n:.2f} |")
# per-kernel durations in the replay
tot = collections.defaultdict(lambda: [0.0, 0])
for e in ks:
    tot[e["name"][:90]][0] += e["dur"]
    tot[e["name"][:90]][1] += 1
lines.append("")
lines.append("| kernel (graph replay) | launches/step | us/step | avg us |")
lines.append("|---|---|---|---|")
for k, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
    lines.append(f"| `{k}` | {n / args.steps:.1f} | {t / args.steps:.1f} | {t / n:.1f} |")
# busy / idle / concurrency over the whole profiled span (all streams). Idle time next to the input copies (the only
# memcpys of a step: they sit between two replays) is the step boundary — graph launch latency, and once per trace the
# profiler's own start-up — and is reported apart from idle time inside a step.
ev = sorted([(e["ts"], 1, e) for e in ks] + [(e["ts"] + e["dur"], -1, e) for e in ks], key=lambda t: (t[0], t[1]))
busy = multi = idle_in = 0.0
boundary = []
depth, last, last_e = 0, ev[0][0], None
for t, d, e in ev:
    if depth >= 1:
        busy += t - last
    elif d == 1 and last_e is not None and t > last:
        if "memcpy" in e.get("cat", "") or "memcpy" in last_e.get("cat", ""):
            boundary.append(t - last)
        else:
            idle_in += t - last
    if depth >= 2:
        multi += t - last
    depth += d
    last = t
    if d == -1:
        last_e = e
ksum = sum(e["dur"] for e in ks)
boundary.sort()
startup = boundary.pop() if boundary else 0.0   # the largest one is the profiler starting up under the first replay
lines.insert(1, f"all streams: sum of kernel durations {ksum / args.steps / 1e3:.3f} ms/step, GPU busy (union) {busy / args.steps / 1e3:.3f}, "
                f"two or more kernels resident {multi / args.steps / 1e3:.3f}, idle inside a step {idle_in / args.steps / 1e3:.3f}, "
                f"idle at a step boundary (input copies -> first kernel of the next replay) {sum(boundary) / max(args.steps - 1, 1) / 1e3:.3f} ms/step "
                f"(+ {startup / 1e3:.2f} ms once: profiler start-up)")
if args.dump:   # every kernel / copy of the profiled span, times relative to the first one
    with open(args.dump, "w") as f:
        f.write("start_us,dur_us,stream,name\n")
        for e in ks:
            f.write(f"{e['ts'] - t_first:.1f},{e['dur']:.1f},{e['args'].get('stream')},\"{e['name'][:110]}\"\n")
text = "\n".join(lines)
print(text)
if args.out:
    open(args.out, "w").write(text + "\n")
