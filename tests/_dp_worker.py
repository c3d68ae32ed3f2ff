"""Worker of tests/test_gpu_multirank.py (launched under torchrun, one rank per GPU): data-parallel training step of
the drop-in ViT through DataParallelTrainer on REAL NCCL, checked against a single-process full-batch gradient.
Prints 'DP_OK <mode>' from rank 0 when every check passed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from neurovit_b200.trainer import DataParallelTrainer, shard_batch  # noqa: E402
from neurovit_b200.vit_3d import ViT  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    graph = os.environ.get("DP_TEST_GRAPH", "1") == "1"
    ctor = dict(image_size=16, image_patch_size=8, frames=24, frame_patch_size=8, num_classes=2, dim=128, depth=2,
                heads=2, mlp_dim=256, channels=1, dim_head=64)
    torch.manual_seed(100 + rank)          # DIFFERENT initial weights per rank: the trainer must broadcast rank 0's
    model = ViT(**ctor).to(dev).train()
    tr = DataParallelTrainer(model, lr=0.0, weight_decay=0.0, graph=graph, bucket_mb=0)   # bucket_mb=0: one bucket per parameter
    # (1) replicas identical after construction
    flat = tr.buckets.flat_params
    ref0 = flat.clone()
    dist.broadcast(ref0, src=0)
    assert torch.equal(flat, ref0), "parameters differ between ranks after DataParallelTrainer construction"
    assert len(tr.buckets.buckets) >= 3, "want several buckets in flight"
    print(f"rank {rank}: replicas identical, {len(tr.buckets.buckets)} buckets", flush=True)
    # (2) N-rank averaged gradients == single-process full-batch gradients
    g = torch.Generator().manual_seed(5)
    per = 6
    X = torch.randn(world * per, 1, 24, 16, 16, generator=g).to(dev)
    Y = torch.randint(0, 2, (world * per,), generator=g).to(dev)
    single = ViT(**ctor).to(dev).train()
    single.load_state_dict(model.state_dict())
    torch.nn.functional.cross_entropy(single(X), Y).backward()
    want = {k: p.grad.clone() for k, p in single.named_parameters()}
    losses = []
    for it in range(4):                    # eager warm-up + capture on the first call, then replays
        losses.append(tr.step(shard_batch(X, rank, world), shard_batch(Y, rank, world)).item())
        torch.cuda.synchronize()
        worst = max(rel(p.grad, want[k]) for k, p in model.named_parameters())
        assert worst < 2e-3, f"rank {rank} step {it}: averaged gradient differs from the full-batch gradient ({worst:.2e})"
        print(f"rank {rank}: step {it} ok ({worst:.1e})", flush=True)
    # every rank holds the same reduced buffer
    red = tr.buckets.flat.clone()
    dist.broadcast(red, src=0)
    assert rel(tr.buckets.flat, red) < 1e-6, "reduced gradient buffers differ between ranks"
    # (3) lr > 0: parameters stay in lock-step over several steps
    tr.optimizer.lr = 1e-3
    tr.reset_graph()                       # hyper-parameters are baked into a captured graph: re-capture
    for it in range(3):
        tr.step(shard_batch(X, rank, world), shard_batch(Y, rank, world))
    torch.cuda.synchronize()
    print(f"rank {rank}: optimizer steps done", flush=True)
    cur = tr.buckets.flat_params.clone()
    dist.broadcast(cur, src=0)
    assert torch.equal(tr.buckets.flat_params, cur), "parameters diverged between ranks after optimizer steps"
    assert not torch.equal(cur, ref0), "optimizer did not move the parameters"
    dist.barrier()
    if rank == 0:
        mode = f"graph={int(tr.use_graph)} two_graphs={int(tr._two_graphs)} own_nccl={int(tr.buckets.comm is not None)}"
        print(f"DP_OK {mode} buckets={len(tr.buckets.buckets)} losses={losses}", flush=True)
    print(f"rank {rank}: closing", flush=True)
    tr.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
