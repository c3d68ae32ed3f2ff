"""Mask generator (nv_dropout_bits) timing at the four site sizes of a cfgA block, L2 flushed between launches.
Usage: python tools/bits_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200 import ops  # noqa: E402

B, N, H, M = 64, 385, 8, 64 * 385
sites = {"attn [B*H, N, 13 words]": B * H * N * 13 * 4, "gelu [M, 2048]": M * 2048 // 8, "out / down [M, 1024]": M * 1024 // 8}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
total = 0.0
for name, nbytes in sites.items():
    buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ops.dropout_bits(buf, p=0.1, seed=1, stream=0)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.dropout_bits(buf, p=0.1, seed=1, stream=0)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    total += med * (2 if "out" in name else 1)
    print(f"  {name}: {med:.1f} us  ({nbytes * 8 / med / 1e3:.1f} Gbit/s)  keep rate {torch.mean((buf.view(torch.int32) & 1).float()).item():.4f}")
print(f"  per transformer block (attn + out + gelu + down): {total:.1f} us; x6 = {6 * total / 1e3:.3f} ms per step")
