"""Phase clocks of the attention kernels (library built with NV_PROFILE=1): prints clock64 cycles per phase for
a few sampled threads. Usage: NV_PROFILE=1 python neurovit_b200/_build.py; NEUROVIT_LIB=neurovit_b200/libneurovit_b200_prof.so python tools/attn_phases.py"""
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
P_DROP = float(os.environ.get("ATTN_DROPOUT", "0"))
mask = torch.zeros(B * H, N, (N + 31) // 32, device="cuda", dtype=torch.int32) if P_DROP > 0 else None
if mask is not None:
    ops.dropout_bits(mask, p=P_DROP, seed=1, stream=0)
kw = dict(B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=P_DROP, drop_mask=mask)
for _ in range(3):
    ops.attention_fwd(qkv, o, lse, seed=1, mask_ready=mask is not None, **kw)
    ops.attention_bwd(qkv, o, dO, lse, ws, dqkv, **kw)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 512)()
lib.nv_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.nv_debug_read(buf, 512) == 0
names = sys.argv[1:] or [f"p{i}" for i in range(12)]
for slot in range(15):
    v = [buf[slot * 12 + i] for i in range(12)]
    if any(v):
        print(f"slot {slot:2d}: total {sum(v):7d} | " + " ".join(f"{n}={x}" for n, x in zip(names, v) if x))
