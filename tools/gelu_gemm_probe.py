"""The two GELU-epilogue GEMMs of a cfgA block (mlp-up forward, mlp-down dgrad x gelu'), with dropout bits pre-drawn,
timed alone with the L2 flushed. Usage: python tools/gelu_gemm_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200 import ops  # noqa: E402

M, D, F = 64 * 385, 1024, 2048
torch.manual_seed(0)
a = torch.randn(M, D, device="cuda").to(torch.bfloat16)
w1 = (torch.randn(F, D, device="cuda") * 0.03).to(torch.bfloat16)
b1 = torch.randn(F, device="cuda") * 0.1
pre = torch.empty(M, F, device="cuda", dtype=torch.bfloat16)
act = torch.empty(M, F, device="cuda", dtype=torch.bfloat16)
bits = torch.empty(M * F // 8, device="cuda", dtype=torch.uint8)
ops.dropout_bits(bits, p=0.1, seed=5, stream=2)
dy = torch.randn(M, D, device="cuda").to(torch.bfloat16)
w2 = (torch.randn(D, F, device="cuda") * 0.03).to(torch.bfloat16)
du = torch.empty(M, F, device="cuda", dtype=torch.bfloat16)
cs = torch.zeros(F, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


fl = 2.0 * M * D * F
t = timeit(lambda: ops.gemm_bf16(a, w1, bias=b1, out_bf16=act, out_pre=pre, apply_gelu=True, dropout=(0.1, 5, 2, bits)))
print(f"  mlp-up forward  (bias + erf-GELU + dropout, writes u and g): {t:.1f} us  {fl / t / 1e6:.0f} TFLOP/s")
t = timeit(lambda: ops.gemm_bf16(dy, w2, b_mn=True, gelu_u=pre, out_bf16=du, colsum=cs, dropout=(0.1, 5, 2, bits)))
print(f"  mlp-down dgrad  (x gelu'(u) x mask, column sums):             {t:.1f} us  {fl / t / 1e6:.0f} TFLOP/s")
x = torch.linspace(-8, 8, M * F, device="cuda").view(M, F)
ref = torch.nn.functional.gelu(pre.double())
print(f"  gelu max abs err vs torch erf-GELU on the bf16 pre-activation: {(act.double()[act != 0] / (65536 / (65536 - 6554)) - ref[act != 0]).abs().max().item():.3e}")
