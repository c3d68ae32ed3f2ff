"""Attention kernels: correctness against a float64 torch reference (with the kernels' own dropout bits replayed)
and timing at the benchmark shape.
Usage: python tools/attn_probe.py [--check-only] [--time-only] [--dropout P] [B N H]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neurovit_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--check-only", action="store_true")
ap.add_argument("--time-only", action="store_true")
ap.add_argument("--dropout", type=float, default=0.0)
ap.add_argument("shape", nargs="*", type=int)
args = ap.parse_args()
hd = 64


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
    """bit position p of a row = key token p + 1 (p < N - 1) or key token 0 (p = N - 1)"""
    w = mask.view(B, H, N, -1).to(torch.int64) & 0xFFFFFFFF
    bits = (w.unsqueeze(-1) >> torch.arange(32, device=w.device)) & 1
    bits = bits.reshape(B, H, N, -1)[..., :N]
    return torch.cat([bits[..., N - 1:N], bits[..., :N - 1]], dim=-1).double()


def worst_rows(got, want, B, N, what):
    """Which tokens carry the error (token 0 is handled outside the tiles)."""
    g, w = got.double().view(B, N, -1), want.double().view(B, N, -1)
    per_tok = (g - w).abs().amax(dim=(0, 2)) / (w.abs().max() + 1e-30)
    top = torch.topk(per_tok, min(4, N))
    nan_rows = torch.isnan(g).any(dim=2).any(dim=0).nonzero().flatten()[:8].tolist()
    print(f"      {what}: worst tokens {top.indices.tolist()} errs {[f'{v:.1e}' for v in top.values.tolist()]} "
          f"token0 {per_tok[0].item():.1e} nan-tokens {nan_rows}")


def check(B, N, H, p_drop=0.0, predrawn=False):
    torch.manual_seed(7)
    qkv = torch.randn(B * N, 3 * H * hd, device="cuda").to(torch.bfloat16)
    o = torch.full((B * N, H * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((B, H, N), float("nan"), device="cuda")
    mask = torch.zeros(B * H, N, (N + 31) // 32, device="cuda", dtype=torch.int32) if p_drop > 0 else None
    if predrawn and p_drop > 0:
        ops.dropout_bits(mask, p=p_drop, seed=1234, stream=0)
    ops.attention_fwd(qkv, o, lse, B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=p_drop, seed=1234,
                      drop_mask=mask, mask_ready=predrawn)
    torch.cuda.synchronize()
    keep, ks = None, 1.0
    if p_drop > 0:
        keep = unpack_mask(mask, B, H, N)
        thr = int(p_drop * 65536 + 0.5)
        ks = 65536.0 / (65536 - thr)
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
    parts = {"o": (o, oref), "dq": (dqkv[:, :inner], gref[:, :inner]), "dk": (dqkv[:, inner:2 * inner], gref[:, inner:2 * inner]),
             "dv": (dqkv[:, 2 * inner:], gref[:, 2 * inner:])}
    errs = {k: rel(*v) for k, v in parts.items()}
    errs["lse"] = rel(lse, lref)
    ok = errs["o"] < 1e-2 and errs["lse"] < 1e-3 and all(errs[k] < 2e-2 for k in (("dv",) if N == 1 else ("dq", "dk", "dv")))
    print(f"  B={B} N={N} H={H} p={p_drop}{' predrawn' if predrawn else ''}: " +
          " ".join(f"{k}={v:.2e}" for k, v in errs.items()) + ("  OK" if ok else "  FAIL"), flush=True)
    if not ok:
        for k, (g, w) in parts.items():
            worst_rows(g.contiguous(), w.contiguous(), B, N, k)
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


def bench(B, N, H, p_drop=0.0):
    torch.manual_seed(0)
    qkv = torch.randn(B * N, 3 * H * hd, device="cuda").to(torch.bfloat16)
    o = torch.empty(B * N, H * hd, device="cuda", dtype=torch.bfloat16)
    dO = torch.randn(B * N, H * hd, device="cuda").to(torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    lse = torch.empty(B, H, N, device="cuda")
    ws = torch.empty(B * H * N, device="cuda")
    mask = torch.zeros(B * H, N, (N + 31) // 32, device="cuda", dtype=torch.int32) if p_drop > 0 else None
    if mask is not None:
        ops.dropout_bits(mask, p=p_drop, seed=1, stream=0)
    kw = dict(B=B, N=N, H=H, head_dim=hd, scale=hd ** -0.5, dropout_p=p_drop, drop_mask=mask)
    fwd = timeit(lambda: ops.attention_fwd(qkv, o, lse, seed=1, mask_ready=mask is not None, **kw))
    bwd = timeit(lambda: ops.attention_bwd(qkv, o, dO, lse, ws, dqkv, **kw))
    flops = 4.0 * B * H * N * N * hd
    print(f"  B={B} N={N} H={H} p={p_drop}: fwd {fwd * 1e3:.1f} us ({flops / fwd / 1e9:.0f} TFLOP/s)  "
          f"bwd {bwd * 1e3:.1f} us ({2.5 * flops / bwd / 1e9:.0f} TFLOP/s)", flush=True)


all_ok = True
if not args.time_only:
    shapes = [tuple(args.shape)] if args.shape else [(1, 64, 2), (1, 129, 1), (3, 9, 2), (2, 1, 2), (2, 200, 2), (2, 385, 8),
                                                     (1, 1001, 2), (1, 1729, 1)]
    for (B, N, H) in shapes:
        all_ok &= check(B, N, H)
        if args.dropout > 0:
            all_ok &= check(B, N, H, args.dropout)
            all_ok &= check(B, N, H, args.dropout, predrawn=True)
if not args.check_only:
    B, N, H = tuple(args.shape) if args.shape else (64, 385, 8)
    bench(B, N, H)
    if args.dropout > 0:
        bench(B, N, H, args.dropout)
    if not args.shape:
        bench(16, 1729, 8)
print("attn_probe:", "all ok" if all_ok else "FAILURES")
sys.exit(0 if all_ok else 1)
