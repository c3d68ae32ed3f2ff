"""Latency of the tcgen05 GEMM on the 64-row problems of the cls-only last layer (and a trivial one: the fixed cost of a
launch), per (block_n, cta_group), L2 warm and flushed. Usage: python tools/small_gemm_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200 import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, cold, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def empty_kernel_floor():
    x = torch.zeros(1, device="cuda")
    return timeit(lambda: x.add_(1), False)


print(f"  event pair around one tiny torch kernel: {empty_kernel_floor():.1f} us")
torch.manual_seed(0)
for (M, N, K, what) in [(64, 128, 64, "trivial"), (64, 1024, 512, "out-proj cls"), (64, 2048, 1024, "mlp-up cls"),
                        (64, 1024, 2048, "mlp-down cls"), (24640, 1024, 512, "out-proj dense")]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda")
    for bn in (128, 256):
        for cg in (1, 2):
            try:
                f = lambda: ops.gemm_bf16(a, w, out_f32=out, block_n=bn, cta_group=cg)
                print(f"  {what:15s} M={M:5d} N={N:4d} K={K:4d} block_n={bn} cta_group={cg}: warm {timeit(f, False):6.1f} us  cold {timeit(f, True):6.1f} us")
            except Exception as e:  # noqa: BLE001
                print(f"  {what} bn={bn} cg={cg}: {type(e).__name__} {str(e)[:80]}")
