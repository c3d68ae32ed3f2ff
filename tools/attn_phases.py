"""Phase clocks of the attention kernels (library built with NV_PROFILE=1): prints clock64 cycles per phase for
a few sampled threads. Usage: NV_PROFILE=1 python neurovit_b200/_build.py --force; python tools/attn_phases.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200 import _lib, ops  # noqa: E402

B, N, H, hd = 64, 385, 8, 64
torch.manual_seed(0)
qkv = torch.randn(B * N, 3 * H * hd, device="cuda").to(torch.bfloat16)
o = torch.empty(B * N, H * hd, device="cuda", dtype=torch.bfloat16)
dO = torch.randn(B * N, H * hd, device="cuda").to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
lse = torch.empty(B, H, N, device="cuda")
ws = torch.empty(B * H * N, device="cuda")
kw = dict(B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5)
for _ in range(3):
    ops.attention_fwd(qkv, o, lse, **kw)
    ops.attention_bwd(qkv, o, dO, lse, ws, dqkv, **kw)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 512)()
lib.nv_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.nv_debug_read(buf, 512) == 0
names = sys.argv[1:] or [f"p{i}" for i in range(12)]
for slot in range(12):
    v = [buf[slot * 12 + i] for i in range(12)]
    if any(v):
        print(f"slot {slot:2d}: total {sum(v):7d} | " + " ".join(f"{n}={x}" for n, x in zip(names, v) if x))
