"""LayerNorm forward/backward timing at the benchmark shape (M = 64*385 rows, D = 1024) with achieved HBM GB/s.
Usage: python tools/ln_probe.py [M D]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200 import ops  # noqa: E402

M, D = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64 * 385, 1024)
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(M, D, device=dev)
g = torch.randn(D, device=dev)
b = torch.randn(D, device=dev)
mean = torch.empty(M, device=dev)
rstd = torch.empty(M, device=dev)
y = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
dres = torch.randn(M, D, device=dev)
dx = torch.empty(M, D, device=dev)
dxb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
acc = torch.zeros(3, D, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


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
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


t = timeit(lambda: ops.layernorm_fwd(x, g, b, y, M=M, D=D, mean=mean, rstd=rstd))
print(f"ln_fwd  fp32->bf16  {t * 1e3:7.1f} us  {M * D * 6 / t / 1e6:7.0f} GB/s")
for dt in (torch.float32, torch.bfloat16):
    dy = torch.randn(M, D, device=dev).to(dt)
    nbytes = M * D * (dy.element_size() + 4 + 4 + 4 + 2)
    t = timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, M=M, D=D, dres=dres, dx=dx, dx_bf16=dxb, dgamma=acc[0],
                                         dbeta=acc[1], colsum=acc[2]))
    print(f"ln_bwd  dy {str(dt)[6:]:9s} {t * 1e3:7.1f} us  {nbytes / t / 1e6:7.0f} GB/s")
c = torch.empty_like(dres)
t = timeit(lambda: c.copy_(dres))
print(f"torch copy fp32     {t * 1e3:7.1f} us  {M * D * 8 / t / 1e6:7.0f} GB/s")
