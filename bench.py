#!/usr/bin/env python
"""Headline benchmark: ViT3D training-step throughput (volumes/sec) on synthetic 1x64x64x48 volumes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference] [--config cfgA|cfgB]

A "step" = forward + CrossEntropy + backward (+ bucketed gradient all-reduce for N>1) + AdamW step on one
batch of `--batch` volumes per GPU (BASELINE.json configs[1]: batch 64, patch 8, 385 tokens, model dims of
NeuroEncoder.py:181-195). N>1 is launched by torchrun, one rank per GPU (weak scaling: per-GPU batch fixed).
Rank 0 prints ONE JSON line (see the driver contract in the task statement).

--impl reference times the reference's CPU path (the oracle port of src/models/vit_3d.py + the step of
src/Trainer.py:65-76, fp32, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# ---- workload definitions (BASELINE.json configs) ------------------------------------------------------
CONFIGS = {
    # volume [H, W, D] as delivered by the datasets (Trainer.py:66), model dims hard-coded in NeuroEncoder.py:181-195
    "cfgA": dict(vol=(64, 64, 48), patch=8, tokens=385, gflop_fwd_bwd=93.469),
    "cfgB": dict(vol=(96, 96, 96), patch=8, tokens=1729, gflop_fwd_bwd=505.432),
    # BASELINE config 5: 4D NeuroEncoder — frozen ViT3D forward over T timepoints + temporal head fwd+bwd+AdamW
    "cfg5": dict(vol=(64, 64, 48), patch=8, tokens=385, T=140, gflop_fwd=31.291),
}
MODEL = dict(dim=1024, depth=6, heads=8, dim_head=64, mlp_dim=2048, num_classes=2)
DROPOUT = 0.1  # TRAINING_DROPOUT of the reference's configs/config.yaml:38 (all 25 sites); --dropout overrides, both arms


def vit_ctor(cfg):
    H, W, D = cfg["vol"]
    return dict(channels=1, image_size=(H, W), image_patch_size=cfg["patch"], frames=D,
                frame_patch_size=cfg["patch"], num_classes=MODEL["num_classes"], dim=MODEL["dim"],
                depth=MODEL["depth"], heads=MODEL["heads"], mlp_dim=MODEL["mlp_dim"], dim_head=MODEL["dim_head"],
                dropout=DROPOUT, emb_dropout=DROPOUT, pool="cls")


def gemm_flops_per_volume(cfg):
    """Algorithmic FLOPs of the linear-layer GEMMs only (2*M*N*K; bwd = 2x fwd, patch embed 1x), per volume —
    the work the dominant kernel (gemm_tc_kernel) does. SURVEY §8d / BASELINE.md §2."""
    N = cfg["tokens"]
    n = N - 1
    p3 = cfg["patch"] ** 3
    D, inner, mlp = MODEL["dim"], MODEL["heads"] * MODEL["dim_head"], MODEL["mlp_dim"]
    per_layer = 2 * N * (D * 3 * inner + inner * D + 2 * D * mlp)
    fwd = MODEL["depth"] * per_layer
    patch = 2 * n * p3 * D
    return 3 * fwd + 2 * patch  # patch: fwd + wgrad + the dgrad used for the patch-LN parameter grads is extra work


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=float(p["bf16_tflops_sustained"]), burst=float(p["bf16_tflops"]), hbm=float(p["hbm_gbs"]),
                    src="measured")
    except Exception:
        return dict(tflops=1400.0, burst=1675.0, hbm=6650.0, src="fallback")  # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 50 ms during the timed region (a 20-step region is < 200 ms)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- the reference arm / cpu_baseline: oracle port on the host cores ------------------------------------
def cpu_reference_step_fn(cfg, batch, seed=42):
    """Returns (step_fn, n_volumes): one fp32 CPU training step of the oracle port (forward, CE, backward, AdamW)."""
    from oracle import vit3d_oracle as O
    from neurovit_b200.vit_3d import ViT  # parameter container only (same init as the reference under the seed)
    torch.manual_seed(seed)
    m = ViT(**vit_ctor(cfg))
    params = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-4, weight_decay=0.01)
    H, W, D = cfg["vol"]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, H, W, D, generator=g)
    y = torch.randint(0, 2, (batch,), generator=g)
    p = cfg["patch"]

    def step():
        opt.zero_grad(set_to_none=True)
        logits = O.vit3d_forward(params, O.neuro_view(x), patch=(p, p, p), heads=MODEL["heads"],
                                 dim_head=MODEL["dim_head"], dropout_p=DROPOUT)
        loss = torch.nn.functional.cross_entropy(logits, y)
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, batch


# ---- stock-PyTorch GPU arm: the kernels to beat (BASELINE.md §3 a-c) -----------------------------------------
def torch_gpu_step_fn(cfg, batch, dev, sdpa, seed=42):
    """One training step of the reference model's math on the GPU through stock PyTorch only: the oracle port
    (same ATen ops the reference dispatches to) under bf16 autocast, eager, fused torch AdamW; sdpa=True swaps the
    explicit softmax attention for F.scaled_dot_product_attention. None of this repo's kernels are involved."""
    from oracle import vit3d_oracle as O
    from neurovit_b200.vit_3d import ViT  # parameter container only
    torch.manual_seed(seed)
    m = ViT(**vit_ctor(cfg))
    params = {k: v.detach().to(dev).requires_grad_(True) for k, v in m.state_dict().items()}
    del m
    opt = torch.optim.AdamW(list(params.values()), lr=1e-4, weight_decay=0.01, fused=True)
    H, W, D = cfg["vol"]
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn(batch, H, W, D, generator=g).to(dev) for _ in range(3)]
    ys = [torch.randint(0, 2, (batch,), generator=g).to(dev) for _ in range(3)]
    p = cfg["patch"]

    def step(i):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = O.vit3d_forward(params, O.neuro_view(xs[i % 3]), patch=(p, p, p), heads=MODEL["heads"],
                                     dim_head=MODEL["dim_head"], dropout_p=DROPOUT, sdpa=sdpa)
            loss = torch.nn.functional.cross_entropy(logits.float(), ys[i % 3])
        loss.backward()
        opt.step()
        return loss

    return step


def time_torch_gpu(cfg, batch, dev, steps, warmup):
    """{variant: volumes/s} for the explicit-softmax and the SDPA variant, CUDA events around `steps` steps."""
    out = {}
    for name, sdpa in (("eager_bf16_autocast", False), ("eager_bf16_autocast_sdpa", True)):
        step = torch_gpu_step_fn(cfg, batch, dev, sdpa)
        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        out[name] = batch * steps / (e0.elapsed_time(e1) * 1e-3)
        del step
        torch.cuda.empty_cache()
    return out


def run_torch_gpu(args, cfg):
    """--impl torch_gpu: rank 0 only, one GPU (a per-GPU number; the data-parallel arm is ours)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local)
    res = time_torch_gpu(cfg, args.batch, dev, args.steps, max(args.warmup, 3))
    clocks = sampler.stop()
    best = max(res, key=res.get)
    peaks = load_peaks()
    tfl = res[best] * cfg["gflop_fwd_bwd"] / 1e3
    print(json.dumps({"impl": "torch_gpu", "metric": "ViT3D training volumes/sec (fwd+bwd+AdamW)", "value": res[best],
                      "unit": "volumes/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
                      "ms_per_step": 1e3 * args.batch / res[best], "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": workload_name(args, cfg, args.batch), "variant": best, "variants": res,
                                 "dropout": DROPOUT, "torch": torch.__version__, "model_tflops_per_gpu": tfl,
                                 "model_frac_of_peak": tfl / peaks["tflops"],
                                 "note": "stock PyTorch (cuBLAS + ATen + SDPA) on the same GPU, none of this repo's kernels"},
                      "clocks": clocks}), flush=True)


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = args.ref_batch
    step, nvol = cpu_reference_step_fn(cfg, batch)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = nvol * args.steps / dt
    sample = f"{nvol} volume(s)/step x {args.steps} steps, fp32, dropout {DROPOUT}, oracle port of vit_3d.py + Trainer.py:65-76"
    line = {"impl": "reference", "metric": "ViT3D training volumes/sec (fwd+bwd+AdamW)", "value": v,
            "unit": "volumes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, cfg, batch), "batch_per_step": batch, "dropout": DROPOUT},
            "cpu_baseline": {"value": v, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(args, cfg, batch):
    H, W, D = cfg["vol"]
    return (f"ViT3D training step, batch {batch} synthetic 1x{H}x{W}x{D} volumes, patch {cfg['patch']} "
            f"({cfg['tokens'] - 1} patches + cls), dim 1024 depth 6 heads 8 mlp 2048")


# ---- secondary workload: BASELINE config 5 (4D NeuroEncoder) -------------------------------------------
def run_4d(args, cfg):
    """One step = NeuroEncoder(TRAINING_DIM=4) forward on `--batch` fMRI sequences [H,W,D,T] (frozen ViT3D over
    B*T volumes, then TemporalTransformer + ProjectionHead), CrossEntropy, backward through the temporal head
    and AdamW on its 10 280 parameters; sequences are sharded over the ranks, the all-reduce is 41 KB."""
    import tempfile
    import torch.distributed as dist
    from neurovit_b200 import _lib
    from neurovit_b200.NeuroEncoder import NeuroEncoder
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device(local)
    H, W, D = cfg["vol"]
    T, B = cfg["T"], args.batch
    base = dict(DEVICE=dev, TRAINING_DROPOUT=DROPOUT, TRAINING_VIT_INPUT_SIZE=H, GRADCAM_CUBE_SIZE=8,
                TRAINING_VIT_PATCH_SIZE=cfg["patch"], DATASET_NAME="adni", GRADCAM_THRESHOLD=0.5, GRADCAM_SLICE_DIM=0,
                GRADCAM_SLICE_IDX=0, GRADCAM_CAPTURE=args.gradcam)
    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(42)
        m3 = NeuroEncoder({**base, "TRAINING_DIM": 3, "GLOBAL_BASE_PATH": tmp, "BEST_MODEL_PATH": "vit3d.pth"})
        torch.save(m3.state_dict(), os.path.join(tmp, "vit3d.pth"))   # the 3D checkpoint the 4D model freezes
        del m3
        model = NeuroEncoder({**base, "TRAINING_DIM": 4, "GLOBAL_BASE_PATH": tmp, "BEST_MODEL_PATH": "vit3d.pth"})
    model.train()
    model.volume_encoder.eval()  # frozen ViT3D (NeuroEncoder.py:33-36)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01)
    g = torch.Generator().manual_seed(42 + rank)
    x = [torch.randn(B, H, W, D, T, generator=g).to(dev) for _ in range(2)]
    y = [torch.randint(0, 2, (B,), generator=g).to(dev) for _ in range(2)]

    def step(i):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(x[i & 1]), y[i & 1])
        loss.backward()
        if world > 1:
            for p_ in params:
                dist.all_reduce(p_.grad, op=dist.ReduceOp.AVG)
        opt.step()
        return loss

    for i in range(max(args.warmup, 3)):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        peaks = load_peaks()
        sps = world * B * args.steps / (ms * 1e-3)
        tfl = sps / world * T * cfg["gflop_fwd"] / 1e3
        print(json.dumps({
            "metric": "4D NeuroEncoder sequences/sec (frozen ViT3D fwd over T volumes + temporal head fwd+bwd+AdamW)",
            "value": sps, "unit": "sequences/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"NeuroEncoder TRAINING_DIM=4, {B} sequences/GPU of {T} x 1x{H}x{W}x{D} volumes, patch "
                                   f"{cfg['patch']}", "volumes_per_s": sps * T, "model_tflops_per_gpu": tfl,
                       "model_frac_of_peak": tfl / peaks["tflops"], "dropout_temporal": 0.1,
                       "gradcam_capture": args.gradcam,
                       "parallelism": f"dp{world} (sequences sharded, 41 KB gradient all-reduce)"}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def roofline_traffic(workload_key):
    """dram bytes of ONE launch of the dominant kernel from the committed ncu --set full capture of this round
    (profiles/roofline_traffic.json, written by tools/ncu_summary.py); None when no capture is recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            rec = json.load(f)[workload_key]
        return float(rec["dram_bytes"]), rec["source"]
    except Exception:
        return None, None


def secondary_vit(cfg_name, batch, dev, steps=6, warmup=3):
    """A short graphed training-step measurement of another ViT3D geometry (same step definition as the headline),
    reported under config.secondary of the default line so the driver-run bench carries it."""
    from neurovit_b200.trainer import DataParallelTrainer
    from neurovit_b200.vit_3d import ViT
    cfg = CONFIGS[cfg_name]

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.vit3d = ViT(**vit_ctor(cfg))

        def forward(self, x):
            return self.vit3d(x.permute(0, 3, 1, 2).unsqueeze(1))

    torch.manual_seed(42)
    enc = Enc().to(dev).train()
    tr = DataParallelTrainer(enc, lr=1e-4, weight_decay=0.01, graph=True)
    H, W, D = cfg["vol"]
    xs = [torch.randn(batch, H, W, D, device=dev) for _ in range(2)]
    ys = [torch.randint(0, 2, (batch,), device=dev) for _ in range(2)]
    for i in range(warmup):
        tr.step(xs[i & 1], ys[i & 1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        tr.step(xs[i & 1], ys[i & 1])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    peaks = load_peaks()
    tfl = batch / (ms * 1e-3) * cfg["gflop_fwd_bwd"] / 1e3
    del tr, enc
    torch.cuda.empty_cache()
    return {"workload": f"ViT3D training step, batch {batch} x 1x{H}x{W}x{D}, patch {cfg['patch']} ({cfg['tokens']} tokens)",
            "value": batch / (ms * 1e-3), "unit": "volumes/s", "ms_per_step": ms, "steps": steps,
            "model_tflops_per_gpu": tfl, "model_frac_of_peak_sustained": tfl / peaks["tflops"],
            "model_frac_of_peak_burst": tfl / peaks["burst"]}


def secondary_neuro3d(dev, batch=64, steps=5, warmup=3):
    """The module the reference Trainer actually trains (src/Trainer.py:65-76 on NeuroEncoder, TRAINING_DIM=3): the
    Grad-CAM hooks of NeuroEncoder.py:70-82 sit on the last block's attention LayerNorm, which takes that block off the
    fused path, and in the reference-default 'host' mode each step pays a device sync and a [B, N, 1024] fp32 D2H copy
    per forward and per backward. 64^3 volumes (ViT3DEncoder is cubic): 513 tokens. One number per capture mode."""
    import tempfile
    from neurovit_b200.NeuroEncoder import NeuroEncoder
    from neurovit_b200.trainer import DataParallelTrainer
    out = {}
    xs = [torch.randn(batch, 64, 64, 64, device=dev) for _ in range(2)]
    ys = [torch.randint(0, 2, (batch,), device=dev) for _ in range(2)]
    for mode in ("off", "device", "host"):
        with tempfile.TemporaryDirectory() as tmp:
            torch.manual_seed(42)
            m = NeuroEncoder(dict(DEVICE=dev, TRAINING_DIM=3, TRAINING_DROPOUT=DROPOUT, TRAINING_VIT_INPUT_SIZE=64,
                                  GRADCAM_CUBE_SIZE=8, TRAINING_VIT_PATCH_SIZE=8, DATASET_NAME="adni",
                                  GRADCAM_THRESHOLD=0.5, GRADCAM_SLICE_DIM=0, GRADCAM_SLICE_IDX=0,
                                  GLOBAL_BASE_PATH=tmp, BEST_MODEL_PATH="x.pth", GRADCAM_CAPTURE=mode)).train()
        tr = DataParallelTrainer(m, lr=1e-4, weight_decay=0.01, graph=(mode != "host"))  # .cpu() cannot be captured
        for i in range(warmup):
            tr.step(xs[i & 1], ys[i & 1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            tr.step(xs[i & 1], ys[i & 1])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[mode] = {"value": batch / (ms * 1e-3), "unit": "volumes/s", "ms_per_step": ms, "cuda_graph": mode != "host"}
        del tr, m
        torch.cuda.empty_cache()
    out["workload"] = (f"NeuroEncoder TRAINING_DIM=3 training step, batch {batch} x 64x64x64 volumes, patch 8 (513 tokens), "
                       "per GRADCAM_CAPTURE mode ('host' = the reference's hooks: sync + D2H copy per forward and backward)")
    return out


def secondary_4d(dev, batch=2, steps=4, warmup=3, frozen_train_mode=False):
    """BASELINE configs[4] on one GPU: frozen ViT3D over batch*T volumes + temporal head fwd+bwd+AdamW.
    frozen_train_mode: leave the frozen ViT3D in train mode as the reference loop does (Trainer.train() calls
    model.train() on the whole NeuroEncoder, so the frozen encoder's 25 dropout sites are live: src/Trainer.py:53);
    False = the encoder in eval mode (deterministic embeddings, what NeuroEncoder.__init__ sets up at :36)."""
    import tempfile
    from neurovit_b200.NeuroEncoder import NeuroEncoder
    cfg = CONFIGS["cfg5"]
    H, W, D = cfg["vol"]
    T = cfg["T"]
    base = dict(DEVICE=dev, TRAINING_DROPOUT=DROPOUT, TRAINING_VIT_INPUT_SIZE=H, GRADCAM_CUBE_SIZE=8,
                TRAINING_VIT_PATCH_SIZE=cfg["patch"], DATASET_NAME="adni", GRADCAM_THRESHOLD=0.5, GRADCAM_SLICE_DIM=0,
                GRADCAM_SLICE_IDX=0, GRADCAM_CAPTURE="device")
    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(42)
        m3 = NeuroEncoder({**base, "TRAINING_DIM": 3, "GLOBAL_BASE_PATH": tmp, "BEST_MODEL_PATH": "vit3d.pth"})
        torch.save(m3.state_dict(), os.path.join(tmp, "vit3d.pth"))
        del m3
        model = NeuroEncoder({**base, "TRAINING_DIM": 4, "GLOBAL_BASE_PATH": tmp, "BEST_MODEL_PATH": "vit3d.pth"})
    model.train()
    if not frozen_train_mode:
        model.volume_encoder.eval()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01)
    xs = [torch.randn(batch, H, W, D, T, device=dev) for _ in range(2)]
    ys = [torch.randint(0, 2, (batch,), device=dev) for _ in range(2)]

    def step(i):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(xs[i & 1]), ys[i & 1])
        loss.backward()
        opt.step()

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    peaks = load_peaks()
    tfl = batch / (ms * 1e-3) * T * cfg["gflop_fwd"] / 1e3
    del model, opt
    torch.cuda.empty_cache()
    return {"workload": f"NeuroEncoder TRAINING_DIM=4, {batch} sequences of {T} x 1x{H}x{W}x{D} volumes "
                        f"(frozen ViT3D forward + temporal head fwd+bwd+AdamW), Grad-CAM capture on device",
            "value": batch / (ms * 1e-3), "unit": "sequences/s", "volumes_per_s": batch * T / (ms * 1e-3),
            "ms_per_step": ms, "steps": steps, "model_tflops_per_gpu": tfl,
            "model_frac_of_peak_sustained": tfl / peaks["tflops"],
            "frozen_encoder_mode": "train (dropout live, as under the reference's Trainer.train())" if frozen_train_mode
            else "eval"}


# ---- our arm ------------------------------------------------------------------------------------------
def run_ours(args, cfg):
    import torch.distributed as dist
    from neurovit_b200 import _lib, ops
    from neurovit_b200.NeuroEncoder import ViT3DEncoder
    from neurovit_b200.trainer import DataParallelTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; neurovit_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device(local)

    H, W, D = cfg["vol"]
    B = args.batch
    torch.manual_seed(42)
    enc = ViT3DEncoder({"DEVICE": dev, "TRAINING_DROPOUT": DROPOUT, "TRAINING_VIT_INPUT_SIZE": H,
                        "GRADCAM_CUBE_SIZE": 8, "TRAINING_VIT_PATCH_SIZE": cfg["patch"], "DATASET_NAME": "adni"}) \
        if H == W == D else None
    if enc is None:  # non-cubic volume (64x64x48): ViT3DEncoder hard-codes frames=image_size, so build ViT directly
        from neurovit_b200.vit_3d import ViT

        class Enc(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.vit3d = ViT(**vit_ctor(cfg))

            def forward(self, x):  # same layout adapter as ViT3DEncoder.forward (NeuroEncoder.py:200-204)
                return self.vit3d(x.permute(0, 3, 1, 2).unsqueeze(1))

        enc = Enc().to(dev)
    enc.train()
    trainer = DataParallelTrainer(enc, lr=1e-4, weight_decay=0.01, graph=not args.no_graph)

    g = torch.Generator().manual_seed(42 + rank)
    n_buf = 3  # rotate host/device input buffers; one batch (50 MB) + activations (GBs) far exceed the 126 MB L2
    host_x = [torch.randn(B, H, W, D, generator=g).pin_memory() for _ in range(n_buf)]
    host_y = [torch.randint(0, 2, (B,), generator=g).pin_memory() for _ in range(n_buf)]
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(loop_fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop_fn(steps)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def resident_loop(steps):
        for i in range(steps):
            trainer.step(dev_x[i % n_buf], dev_y[i % n_buf])

    losses = []

    copy_stream = torch.cuda.Stream(device=dev)
    stage_x = [torch.empty(B, H, W, D, device=dev) for _ in range(2)]
    stage_y = [torch.empty(B, dtype=torch.int64, device=dev) for _ in range(2)]

    def e2e_loop(steps):
        """Public-API loop with HOST inputs: every step's batch is copied from pinned host memory inside the timed
        region (on a copy stream, one step ahead — what a pinned-memory DataLoader with non_blocking copies does)
        and the loss is read back to the host."""
        main = torch.cuda.current_stream()
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def stage(i):
            k = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[k])
                stage_x[k].copy_(host_x[i % n_buf], non_blocking=True)
                stage_y[k].copy_(host_y[i % n_buf], non_blocking=True)
                ready[k].record(copy_stream)

        for ev in consumed:
            ev.record(main)
        stage(0)
        for i in range(steps):
            k = i & 1
            if i + 1 < steps:
                stage(i + 1)
            main.wait_event(ready[k])
            loss = trainer.step(stage_x[k], stage_y[k])
            consumed[k].record(main)
            losses.append(loss.item())  # device -> host read of the step's result

    resident_loop(args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.LAUNCHES.reset()
    ms = timed(resident_loop, args.steps)           # the headline: no per-kernel events inside
    launches = _lib.LAUNCHES.count
    clocks = sampler.stop() if sampler else None
    # second pass of the same loop with a CUDA-event pair around every GEMM launch (roofline numbers); kept out
    # of the headline pass because ~75 event pairs per step cost ~5 % of the step
    ops.PROFILE.reset()
    ops.PROFILE.enabled = not args.no_kernel_events
    prof_steps = min(args.steps, 10)
    graphed, trainer.use_graph = trainer.use_graph, False   # events cannot be recorded inside a replayed graph
    ops.PROFILE.enabled = False
    resident_loop(2)                                        # the eager path's own allocations, outside the graph pool
    ops.PROFILE.reset()
    ops.PROFILE.enabled = not args.no_kernel_events
    ms_prof = timed(resident_loop, prof_steps)
    trainer.use_graph = graphed
    ops.PROFILE.enabled = False
    gemm_ms, gemm_flops, gemm_calls = ops.PROFILE.summary()
    e2e_loop(min(2, args.warmup))
    ms_e2e = timed(e2e_loop, args.steps)

    # BASELINE configs[2] as stated — global batch 512 — next to the weak-scaling headline: 256 / 128 per GPU at
    # N = 2 / 4 (at N = 8 the headline itself is global 512). Same trainer, graph re-captured for the new shape.
    strong = None
    if world in (2, 4) and args.config == "cfgA" and B == 64 and not args.no_secondary:
        Bs = 512 // world
        sx = [torch.randn(Bs, H, W, D, generator=g).to(dev) for _ in range(2)]
        sy = [torch.randint(0, 2, (Bs,), generator=g).to(dev) for _ in range(2)]
        trainer.reset_graph()

        def strong_loop(steps):
            for i in range(steps):
                trainer.step(sx[i & 1], sy[i & 1])

        strong_loop(3)
        k = max(4, args.steps // 2)
        ms_s = timed(strong_loop, k)
        strong = {"workload": f"global batch 512 = {Bs} volumes per GPU x {world} GPUs (BASELINE configs[2])",
                  "value": 512 * k / (ms_s * 1e-3), "unit": "volumes/s", "ms_per_step": ms_s / k, "steps": k,
                  "scaling": "strong"}

    if rank == 0:
        peaks = load_peaks()
        vps = world * B * args.steps / (ms * 1e-3)
        vps_e2e = world * B * args.steps / (ms_e2e * 1e-3)
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        model_tflops = vps / world * cfg["gflop_fwd_bwd"] / 1e3
        traffic, traffic_src = roofline_traffic(f"{args.config}_b{B}")
        line = {
            "metric": "ViT3D training volumes/sec (fwd+bwd+AdamW)", "value": vps, "unit": "volumes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args, cfg, B), "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world}", "dropout": DROPOUT, "optimizer": "AdamW (FlatAdamW, one kernel)",
                       "cuda_graph": bool(trainer.use_graph),
                       "l2_policy": "inputs+activations per step (>3 GB) exceed the 126 MB L2; 3 rotating input buffers",
                       "model_tflops_per_gpu": model_tflops,
                       "model_frac_of_peak": model_tflops / peaks["tflops"],
                       "model_frac_of_peak_sustained": model_tflops / peaks["tflops"],
                       "model_frac_of_peak_burst": model_tflops / peaks["burst"],
                       "peaks_tflops": {"sustained": peaks["tflops"], "burst": peaks["burst"], "source": peaks["src"]}},
            "clocks": clocks,
            "e2e": {"value": vps_e2e, "unit": "volumes/s", "h2d_bytes_per_step": B * H * W * D * 4 + B * 8,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 linear fwd/dgrad/wgrad)",
                         "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["tflops"]) if achieved else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the qkv-forward instance, read
                         # from the committed capture record of this round (null when there is none)
                         "traffic": traffic, "traffic_source": traffic_src,
                         "frac_of_burst": (achieved / peaks["burst"]) if achieved else None,
                         "peak_source": f"{peaks['src']} bf16_tflops_sustained",
                         "launches_timed": gemm_calls, "avg_launch_ms": gemm_ms / max(gemm_calls, 1),
                         # GEMM device time per step over the HEADLINE step time: the second pass launches kernel
                         # by kernel with events in between, its own step time says nothing about the graph replay
                         "share_of_step": (gemm_ms / prof_steps) / (ms / args.steps) if ms > 0 else None,
                         "share_of_profiled_pass": gemm_ms / ms_prof if ms_prof > 0 else None,
                         "measured": f"CUDA events around each GEMM launch in a second pass of {prof_steps} steps "
                                     f"({ms_prof / prof_steps:.3f} ms/step with the events in)"},
        }
        if strong is not None:
            line["config"]["config3_global512"] = strong
        if world == 1 and not args.no_secondary:
            # the kernels to beat: stock PyTorch (cuBLAS + ATen, with and without SDPA) on this GPU, same step
            del trainer
            torch.cuda.empty_cache()
            try:
                tg = time_torch_gpu(cfg, B, dev, steps=5, warmup=3)
                best = max(tg, key=tg.get)
                line["config"]["torch_gpu_baseline"] = {
                    "value": tg[best], "unit": "volumes/s", "variant": best, "variants": tg,
                    "ours_over_best": vps / tg[best],
                    "note": "oracle port on CUDA, bf16 autocast, eager, fused torch AdamW (stock cuBLAS / ATen / SDPA)"}
            except Exception as e:  # noqa: BLE001
                line["config"]["torch_gpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
            if args.config == "cfgA" and B == 64:
                def guarded(fn, *a, **k):   # a failing side measurement must never cost the headline line
                    try:
                        return fn(*a, **k)
                    except Exception as e:  # noqa: BLE001
                        torch.cuda.empty_cache()
                        return {"error": f"{type(e).__name__}: {e}"[:300]}

                line["config"]["secondary"] = {
                    "cfgB": guarded(secondary_vit, "cfgB", 16, dev), "cfg5": guarded(secondary_4d, dev),
                    "cfg5_frozen_encoder_in_train_mode": guarded(secondary_4d, dev, frozen_train_mode=True),
                    "neuroencoder3d_gradcam": guarded(secondary_neuro3d, dev)}
        if not args.skip_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            step, nvol = cpu_reference_step_fn(cfg, 1)
            step()
            t0 = time.perf_counter()
            n = 0
            while n < 3 or (time.perf_counter() - t0 < 10 and n < 40):
                step()
                n += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nvol * n / dt, "unit": "volumes/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} steps of 1 volume (fwd+CE+bwd+AdamW), fp32, torch CPU, oracle port"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    global DROPOUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="volumes per GPU per step")
    ap.add_argument("--config", default="cfgA", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--ref-batch", type=int, default=2, help="volumes per CPU reference step (bounded sample)")
    ap.add_argument("--dropout", type=float, default=DROPOUT, help="dropout p at all sites, training mode (both arms)")
    ap.add_argument("--gradcam", default="device", choices=["host", "device", "off"],
                    help="cfg5 only: NeuroEncoder Grad-CAM capture ('host' = the reference's per-step D2H copies)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of one CUDA graph")
    ap.add_argument("--no-kernel-events", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the torch-GPU baseline, the cfgB / cfg5 secondary workloads and the global-512 line")
    args = ap.parse_args()
    DROPOUT = args.dropout
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        if args.steps > 6:
            args.steps = 6  # bounded sample: the whole run must end within minutes on the host cores
        if args.warmup > 2:
            args.warmup = 2
        run_reference(args, cfg)
    elif args.impl == "torch_gpu":
        run_torch_gpu(args, cfg)
    elif args.config == "cfg5":
        if args.batch == 64:
            args.batch = 2  # 2 sequences x 140 timepoints = 280 volumes per GPU per step
        run_4d(args, cfg)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
