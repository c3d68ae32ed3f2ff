"""Where does (step time - sum of kernel time) go in the replayed training-step graph? Times, with CUDA events:
  A  replay only (no input copies)          B  the trainer's step (copies + replay)
  C  two steps captured into ONE graph      D  per-replay device time (event pair around every replay)
Usage: python tools/graph_gap_probe.py [--dropout 0.1]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from neurovit_b200.trainer import DataParallelTrainer  # noqa: E402
from neurovit_b200.vit_3d import ViT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dropout", type=float, default=bench.DROPOUT)
ap.add_argument("--batch", type=int, default=64)
args = ap.parse_args()
bench.DROPOUT = args.dropout
cfg = bench.CONFIGS["cfgA"]
dev = torch.device("cuda", 0)
torch.manual_seed(42)


class Enc(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.vit3d = ViT(**bench.vit_ctor(cfg))

    def forward(self, x):
        return self.vit3d(x.permute(0, 3, 1, 2).unsqueeze(1))


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


enc = Enc().to(dev).train()
os.environ["NEUROVIT_GRAPH_PINGPONG"] = "1"
tr = DataParallelTrainer(enc, lr=1e-4, weight_decay=0.01, graph=True)
H, W, D = cfg["vol"]
x = torch.randn(args.batch, H, W, D, device=dev)
y = torch.randint(0, 2, (args.batch,), device=dev)
for i in range(5):
    tr.step(x, y)
print(f"B  trainer.step (copies + ping-pong replay): {timed(lambda i: tr.step(x, y), 20):.3f} ms/step")
print(f"A  one executable replayed back to back:     {timed(lambda i: tr._graph.replay(), 20):.3f} ms/step")
print(f"A2 two executables alternating, no copies:   {timed(lambda i: (tr._graph_alt if i & 1 else tr._graph).replay(), 20):.3f} ms/step")
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]


def one(i):
    evs[i][0].record()
    tr._graph.replay()
    evs[i][1].record()


loop = timed(one, 20)
per = sorted(a.elapsed_time(b) for a, b in evs)
print(f"D  event pair around each replay: median {per[10]:.3f} ms, loop {loop:.3f} ms/step")
# C: two steps in one graph
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2, pool=tr._graph.pool()):
    tr._eager_step(tr._sx, tr._sy)
    tr._eager_step(tr._sx, tr._sy)
print(f"C  two steps captured in one graph:          {timed(lambda i: g2.replay(), 10) / 2:.3f} ms/step")
tr.use_graph = False
for i in range(3):
    tr.step(x, y)
print(f"E  eager (kernel by kernel):                 {timed(lambda i: tr.step(x, y), 10):.3f} ms/step")
