"""Attention kernel timing at the benchmark shapes. Usage: python tools/attn_probe.py [B N H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200 import ops  # noqa: E402

B, N, H = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 385, 8)
hd = 64
torch.manual_seed(0)
qkv = torch.randn(B * N, 3 * H * hd, device="cuda").to(torch.bfloat16)
o = torch.empty(B * N, H * hd, device="cuda", dtype=torch.bfloat16)
dO = torch.randn(B * N, H * hd, device="cuda").to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
lse = torch.empty(B, H, N, device="cuda")
ws = torch.empty(B * H * N, device="cuda")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


fwd = timeit(lambda: ops.attention_fwd(qkv, o, lse, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5))
bwd = timeit(lambda: ops.attention_bwd(qkv, o, dO, lse, ws, dqkv, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5))
flops = 4.0 * B * H * N * N * hd
print(f"attention B={B} N={N} H={H}: fwd {fwd * 1e3:.1f} us ({flops / fwd / 1e9:.0f} TFLOP/s)  "
      f"bwd {bwd * 1e3:.1f} us ({2.5 * flops / bwd / 1e9:.0f} TFLOP/s)")
