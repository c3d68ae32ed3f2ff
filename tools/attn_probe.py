"""Attention kernels: correctness against a float64 torch reference and timing at the benchmark shape.
Usage: python tools/attn_probe.py [--impl 0|1|both] [--check-only] [--time-only] [--dropout P] [B N H]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="both")
ap.add_argument("--check-only", action="store_true")
ap.add_argument("--time-only", action="store_true")
ap.add_argument("--dropout", type=float, default=0.0)
ap.add_argument("shape", nargs="*", type=int)
args = ap.parse_args()
hd = 64
NAMES = {0: "tcgen05", 1: "mma.sync"}


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def ref(qkv, B, N, H, keep=None, ks=1.0):
    q, k, v = qkv.view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    p = s.softmax(-1)
    if keep is not None:
        p = p * keep * ks
    return (p @ v).permute(0, 2, 1, 3).reshape(B * N, H * hd), torch.logsumexp(s, -1)


def unpack_mask(mask, B, H, N):
    w = mask.view(B, H, N, -1).to(torch.int64) & 0xFFFFFFFF
    bits = (w.unsqueeze(-1) >> torch.arange(32, device=w.device)) & 1
    return bits.reshape(B, H, N, -1)[..., :N].double()


def check(impl, B, N, H, p_drop=0.0):
    torch.manual_seed(7)
    qkv = torch.randn(B * N, 3 * H * hd, device="cuda").to(torch.bfloat16)
    o = torch.full((B * N, H * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((B, H, N), float("nan"), device="cuda")
    mask = torch.zeros(B * H, N, (N + 31) // 32, device="cuda", dtype=torch.int32) if p_drop > 0 else None
    ops.attention_fwd(qkv, o, lse, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=p_drop, seed=1234,
                      drop_mask=mask)
    torch.cuda.synchronize()
    keep, ks = None, 1.0
    if p_drop > 0:
        keep = unpack_mask(mask, B, H, N)
        thr = int(p_drop * 65536 + 0.5)
        ks = 65536.0 / (65536 - thr)
        print(f"    keep rate {keep.mean().item():.4f} (expect {1 - thr / 65536:.4f})")
    qd = qkv.double().requires_grad_(True)
    oref, lref = ref(qd, B, N, H, keep, ks)
    dO = torch.randn(B * N, H * hd, device="cuda").to(torch.bfloat16)
    gref, = torch.autograd.grad(oref, qd, dO.double())
    dqkv = torch.full_like(qkv, float("nan"))
    ws = torch.empty(B * H * N, device="cuda")
    ops.attention_bwd(qkv, o, dO, lse, ws, dqkv, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=p_drop,
                      drop_mask=mask)
    torch.cuda.synchronize()
    inner = H * hd
    errs = {"o": rel(o, oref), "lse": rel(lse, lref), "dq": rel(dqkv[:, :inner], gref[:, :inner]),
            "dk": rel(dqkv[:, inner:2 * inner], gref[:, inner:2 * inner]),
            "dv": rel(dqkv[:, 2 * inner:], gref[:, 2 * inner:])}
    ok = errs["o"] < 1e-2 and errs["lse"] < 1e-3 and all(errs[k] < 2e-2 for k in ("dq", "dk", "dv"))
    print(f"  [{NAMES[impl]}] B={B} N={N} H={H} p={p_drop}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()) +
          ("  OK" if ok else "  FAIL"))
    return ok


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


def bench(impl, B, N, H, p_drop=0.0):
    torch.manual_seed(0)
    qkv = torch.randn(B * N, 3 * H * hd, device="cuda").to(torch.bfloat16)
    o = torch.empty(B * N, H * hd, device="cuda", dtype=torch.bfloat16)
    dO = torch.randn(B * N, H * hd, device="cuda").to(torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    lse = torch.empty(B, H, N, device="cuda")
    ws = torch.empty(B * H * N, device="cuda")
    mask = torch.zeros(B * H, N, (N + 31) // 32, device="cuda", dtype=torch.int32) if p_drop > 0 else None
    kw = dict(B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=p_drop, drop_mask=mask)
    fwd = timeit(lambda: ops.attention_fwd(qkv, o, lse, seed=1, **kw))
    bwd = timeit(lambda: ops.attention_bwd(qkv, o, dO, lse, ws, dqkv, **kw))
    flops = 4.0 * B * H * N * N * hd
    print(f"  [{NAMES[impl]}] B={B} N={N} H={H} p={p_drop}: fwd {fwd * 1e3:.1f} us ({flops / fwd / 1e9:.0f} TFLOP/s)  "
          f"bwd {bwd * 1e3:.1f} us ({2.5 * flops / bwd / 1e9:.0f} TFLOP/s)")


impls = [0, 1] if args.impl == "both" else [int(args.impl)]
all_ok = True
for impl in impls:
    ops.set_attention_impl(impl)
    p_drop = args.dropout if impl == 0 else 0.0
    if not args.time_only:
        shapes = [tuple(args.shape)] if args.shape else [(1, 64, 2), (1, 128, 1), (3, 9, 2), (2, 200, 2), (2, 385, 8),
                                                         (1, 1729, 1)]
        for (B, N, H) in shapes:
            all_ok &= check(impl, B, N, H)
            if p_drop > 0:
                all_ok &= check(impl, B, N, H, p_drop)
    if not args.check_only:
        B, N, H = tuple(args.shape) if args.shape else (64, 385, 8)
        bench(impl, B, N, H)
        if p_drop > 0:
            bench(impl, B, N, H, p_drop)
ops.set_attention_impl(0)
print("attn_probe:", "all ok" if all_ok else "FAILURES")
sys.exit(0 if all_ok else 1)
