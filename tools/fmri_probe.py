"""4D input pipeline kernel (csrc/fmri4d.cu): time at BASELINE configs[4]'s sample shape against torch's
permute().reshape() copy, plain and with the fused z-score. Usage: python tools/fmri_probe.py [B H W D T]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200 import functional as Fn  # noqa: E402

B, H, W, D, T = (int(v) for v in sys.argv[1:6]) if len(sys.argv) >= 6 else (2, 64, 64, 48, 140)
x = torch.randn(B, H, W, D, T, device="cuda")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def timeit(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


ref = lambda: x.permute(0, 4, 1, 2, 3).reshape(B * T, H, W, D)
assert torch.equal(Fn.fmri_to_volumes(x), ref())
nbytes = 2 * x.numel() * 4
for name, fn in (("torch permute+reshape", ref), ("nv_fmri_deinterleave", lambda: Fn.fmri_to_volumes(x)),
                 ("nv_fmri_deinterleave + z-score", lambda: Fn.fmri_to_volumes(x, zscore=True))):
    ms = timeit(fn)
    extra = x.numel() * 4 if "z-score" in name else 0
    print(f"{name:32s} {ms * 1e3:8.1f} us   {(nbytes + extra) / ms / 1e6:7.0f} GB/s ({(nbytes + extra) / 1e6:.0f} MB algorithmic)")
