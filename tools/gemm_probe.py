"""Bring-up probe for the tcgen05 GEMM: runs every operand-major variant / tile / epilogue in its own
subprocess (a trapped kernel poisons the CUDA context) and prints one line per case.
Usage (GPU box): python tools/gemm_probe.py [--quick]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = []
for a_mn in (0, 1):
    for b_mn in (0, 1):
        for block_n in (128, 256):
            CASES.append(dict(a_mn=a_mn, b_mn=b_mn, block_n=block_n, M=256, N=256, K=128, epi="plain"))
            CASES.append(dict(a_mn=a_mn, b_mn=b_mn, block_n=block_n, M=385, N=1536, K=1024, epi="plain"))
CASES += [
    dict(a_mn=0, b_mn=0, block_n=256, M=24640, N=1024, K=512, epi="bias_res"),
    dict(a_mn=0, b_mn=0, block_n=256, M=24640, N=2048, K=1024, epi="gelu"),
    dict(a_mn=0, b_mn=0, block_n=256, M=1000, N=1024, K=2048, epi="gelu_grad"),
    dict(a_mn=1, b_mn=1, block_n=256, M=1024, N=2048, K=24640, epi="splitk"),
    dict(a_mn=1, b_mn=1, block_n=128, M=1536, N=1024, K=3000, epi="splitk"),
    dict(a_mn=0, b_mn=0, block_n=128, M=72, N=64, K=64, epi="bias_res"),
]


def run_case(c):
    import torch
    from neurovit_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    M, N, K = c["M"], c["N"], c["K"]
    a = torch.randn((K, M) if c["a_mn"] else (M, K), device=dev).to(torch.bfloat16)
    b = torch.randn((K, N) if c["b_mn"] else (N, K), device=dev).to(torch.bfloat16)
    af = a.float().t() if c["a_mn"] else a.float()
    bf = b.float().t() if c["b_mn"] else b.float()
    ref = af.double() @ bf.double().t()
    out = torch.zeros(M, N, device=dev)
    epi = c["epi"]
    kw = {}
    extra = {}
    if epi == "bias_res":
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        kw = dict(bias=bias, residual=res)
        ref = ref + bias.double() + res.double()
    elif epi == "gelu":
        bias = torch.randn(N, device=dev)
        pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        act = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw = dict(bias=bias, out_pre=pre, out_bf16=act, apply_gelu=True)
        u = ref + bias.double()
        ref = torch.nn.functional.gelu(u)
        extra = dict(pre=(pre, u), act=(act, ref))
    elif epi == "gelu_grad":
        u = torch.randn(M, N, device=dev).to(torch.bfloat16)
        kw = dict(gelu_u=u)
        ud = u.double().requires_grad_(True)
        g, = torch.autograd.grad(torch.nn.functional.gelu(ud).sum(), ud)
        ref = ref * g
    elif epi == "splitk":
        kw = dict(accumulate=True, k_splits=8)
        out.fill_(1.0)
        ref = ref + 1.0
    ops.gemm_bf16(a, b, a_mn=bool(c["a_mn"]), b_mn=bool(c["b_mn"]), out_f32=out, block_n=c["block_n"], **kw)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (out.double() - ref).abs().max().item()
    res = dict(case=c, max_abs_err=err, ref_max=scale, rel=err / scale)
    for nm, (t, r) in extra.items():
        res[nm + "_rel"] = ((t.double() - r).abs().max() / r.abs().max()).item()
    res["ok"] = bool(err / scale < 2e-3 and all(v < 1e-2 for k_, v in res.items() if k_.endswith("_rel")))
    if not res["ok"]:
        # localise: which rows/cols are wrong?
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
    cases = CASES[:4] if "--quick" in sys.argv else CASES
    n_ok = 0
    for c in cases:
        try:
            r = subprocess.run([sys.executable, __file__, "--case", json.dumps(c)], capture_output=True, text=True,
                               timeout=120)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                res = json.loads(lines[-1][7:])
                n_ok += int(res["ok"])
                print(("PASS " if res["ok"] else "FAIL ") + json.dumps(res))
            else:
                print("CRASH " + json.dumps(c) + " rc=%d\n%s\n%s" % (r.returncode, r.stdout[-600:], r.stderr[-1200:]))
        except subprocess.TimeoutExpired:
            print("TIMEOUT " + json.dumps(c))
        sys.stdout.flush()
    print(f"gemm_probe: {n_ok}/{len(cases)} passed")
