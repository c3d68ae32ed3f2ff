"""Bring-up + performance probe for the tcgen05 GEMM: every operand-major variant / tile / CTA-group /
epilogue in its own subprocess (a trapped kernel poisons the CUDA context); one line per case.
Usage (GPU box): python tools/gemm_probe.py [--quick] [--perf]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = []
for cg in (1, 2):
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            for block_n in (128, 256):
                CASES.append(dict(cg=cg, a_mn=a_mn, b_mn=b_mn, block_n=block_n, M=256, N=256, K=128, epi="plain"))
                CASES.append(dict(cg=cg, a_mn=a_mn, b_mn=b_mn, block_n=block_n, M=392, N=1536, K=1024, epi="plain"))
    CASES += [
        dict(cg=cg, a_mn=0, b_mn=0, block_n=256, M=24640, N=1024, K=512, epi="bias_res"),
        dict(cg=cg, a_mn=0, b_mn=0, block_n=256, M=24640, N=2048, K=1024, epi="gelu"),
        dict(cg=cg, a_mn=0, b_mn=1, block_n=256, M=1000, N=1024, K=2048, epi="gelu_grad"),
        dict(cg=cg, a_mn=1, b_mn=1, block_n=256, M=1024, N=2048, K=24640, epi="splitk"),
        dict(cg=cg, a_mn=1, b_mn=1, block_n=128, M=1536, N=1024, K=3000, epi="splitk"),
        dict(cg=cg, a_mn=0, b_mn=0, block_n=128, M=72, N=64, K=64, epi="bias_res"),
    ]

# the twelve GEMM shapes of one transformer block at B=64 (M = 64*385 = 24640), SURVEY §8a
PERF = []
for cg in (1, 2):
    for (M, N, K, a_mn, b_mn, epi, name) in [
        (24640, 1536, 1024, 0, 0, "plain_bf16", "qkv fwd"), (24640, 1024, 512, 0, 0, "bias_res", "out fwd"),
        (24640, 2048, 1024, 0, 0, "gelu", "mlp-up fwd"), (24640, 1024, 2048, 0, 0, "bias_res", "mlp-down fwd"),
        (24640, 1024, 1536, 0, 1, "plain", "qkv dgrad"), (24640, 512, 1024, 0, 1, "plain_bf16", "out dgrad"),
        (24640, 2048, 1024, 0, 1, "gelu_grad", "mlp-down dgrad"), (24640, 1024, 2048, 0, 1, "plain", "mlp-up dgrad"),
        (1536, 1024, 24640, 1, 1, "splitk", "qkv wgrad"), (1024, 512, 24640, 1, 1, "splitk", "out wgrad"),
        (2048, 1024, 24640, 1, 1, "splitk", "mlp-up wgrad"), (1024, 2048, 24640, 1, 1, "splitk", "mlp-down wgrad"),
    ]:
        PERF.append(dict(cg=cg, a_mn=a_mn, b_mn=b_mn, block_n=256, M=M, N=N, K=K, epi=epi, name=name, perf=1))


def run_case(c):
    import math
    import torch
    from neurovit_b200 import ops
    from neurovit_b200.functional import _splitk
    torch.manual_seed(0)
    dev = "cuda"
    M, N, K = c["M"], c["N"], c["K"]
    a = torch.randn((K, M) if c["a_mn"] else (M, K), device=dev).to(torch.bfloat16)
    b = torch.randn((K, N) if c["b_mn"] else (N, K), device=dev).to(torch.bfloat16)
    perf = bool(c.get("perf"))
    out = torch.zeros(M, N, device=dev)
    epi = c["epi"]
    kw = dict(out_f32=out)
    extra = {}
    ref = None
    if not perf:
        af = a.float().t() if c["a_mn"] else a.float()
        bf = b.float().t() if c["b_mn"] else b.float()
        ref = af.double() @ bf.double().t()
    if epi == "plain_bf16":
        kw = dict(out_bf16=torch.empty(M, N, device=dev, dtype=torch.bfloat16))
    elif epi == "bias_res":
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        kw.update(bias=bias, residual=res)
        if ref is not None:
            ref = ref + bias.double() + res.double()
    elif epi == "gelu":
        bias = torch.randn(N, device=dev)
        pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        act = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw = dict(bias=bias, out_pre=pre, out_bf16=act, apply_gelu=True)
        if ref is not None:
            u = ref + bias.double()
            ref = torch.nn.functional.gelu(u)
            extra = dict(pre=(pre, u), act=(act, ref))
            out = act
    elif epi == "gelu_grad":
        u = torch.randn(M, N, device=dev).to(torch.bfloat16)
        cs = torch.zeros(N, device=dev)
        kw = dict(gelu_u=u, out_bf16=torch.empty(M, N, device=dev, dtype=torch.bfloat16), colsum=cs)
        if ref is not None:
            ud = u.double().requires_grad_(True)
            g, = torch.autograd.grad(torch.nn.functional.gelu(ud).sum(), ud)
            ref = ref * g
            out = kw["out_bf16"]
            extra = dict(colsum=(cs, ref.sum(0)))
    elif epi == "splitk":
        tiles = math.ceil(M / (128 * c["cg"])) * math.ceil(N / c["block_n"])
        kw.update(accumulate=True, k_splits=_splitk(tiles, math.ceil(K / 64), 148 // c["cg"]) if perf else 8)
        out.fill_(1.0)
        if ref is not None:
            ref = ref + 1.0
    call = lambda: ops.gemm_bf16(a, b, a_mn=bool(c["a_mn"]), b_mn=bool(c["b_mn"]), block_n=c["block_n"],
                                 cta_group=c["cg"], **kw)
    call()
    torch.cuda.synchronize()
    res = dict(case=c)
    if perf:
        for _ in range(3):
            call()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for _ in range(10):
            flush.zero_()  # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        res.update(ms_med=ts[len(ts) // 2], ms_min=ts[0], tflops=2.0 * M * N * K / (ts[len(ts) // 2] * 1e-3) / 1e12,
                   ok=True, k_splits=kw.get("k_splits", 1))
    else:
        scale = ref.abs().max().item()
        err = (out.double() - ref).abs().max().item()
        tol = 1e-2 if out.dtype == torch.bfloat16 else 2e-3
        res.update(max_abs_err=err, ref_max=scale, rel=err / scale)
        for nm, (t, r) in extra.items():
            res[nm + "_rel"] = ((t.double() - r).abs().max() / r.abs().max()).item()
        res["ok"] = bool(err / scale < tol and all(v < 1e-2 for k_, v in res.items() if k_.endswith("_rel")))
        if not res["ok"]:
            bad = ((out.double() - ref).abs() > 1e-2 * scale)
            res["bad_frac"] = bad.float().mean().item()
            res["bad_rows"] = bad.any(1).nonzero().flatten()[:8].tolist()
            res["bad_cols"] = bad.any(0).nonzero().flatten()[:8].tolist()
            res["sample"] = [out[0, :4].tolist(), ref[0, :4].tolist()]
    print("RESULT " + json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        run_case(json.loads(sys.argv[2]))
        sys.exit(0)
    cases = PERF if "--perf" in sys.argv else (CASES[:4] if "--quick" in sys.argv else CASES)
    n_ok = 0
    for c in cases:
        try:
            r = subprocess.run([sys.executable, __file__, "--case", json.dumps(c)], capture_output=True, text=True,
                               timeout=180)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                res = json.loads(lines[-1][7:])
                n_ok += int(res["ok"])
                if c.get("perf"):
                    print(f"PERF cg={c['cg']} {c['name']:15s} M={c['M']:5d} N={c['N']:4d} K={c['K']:5d} "
                          f"splits={res['k_splits']:2d} {res['ms_med'] * 1e3:7.1f} us  {res['tflops']:7.1f} TFLOP/s")
                else:
                    print(("PASS " if res["ok"] else "FAIL ") + json.dumps(res))
            else:
                print("CRASH " + json.dumps(c) + " rc=%d\n%s\n%s" % (r.returncode, r.stdout[-600:], r.stderr[-1500:]))
        except subprocess.TimeoutExpired:
            print("TIMEOUT " + json.dumps(c))
        sys.stdout.flush()
    print(f"gemm_probe: {n_ok}/{len(cases)} ok")
