"""Data-parallel training step for the ViT3D hot path: one process per GPU, batch sharding at the step
boundary and a bucketed gradient all-reduce overlapped with backward.

Reference: src/Trainer.py:65-76 is the single-GPU step (forward, CrossEntropy, zero_grad, backward, AdamW
step; fp16 autocast + GradScaler there, bf16 operands with fp32 master weights here — no loss scaling
needed). The reference has no distributed code; data parallelism is the new work BASELINE.json
asks for: rank r takes batch[r*B/g:(r+1)*B/g]; gradients live in ONE flat fp32 buffer laid out in reverse
parameter order (the order backward produces them), cut into buckets; each bucket is all-reduced (NCCL,
average) as soon as autograd has accumulated its last gradient, while the rest of backward still runs.
Inference shards the batch with no communication.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .functional import SINKS


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rows [rank*B/world, (rank+1)*B/world) of the global batch (global batch must divide evenly)."""
    B = t.shape[0]
    if B % world:
        raise ValueError(f"global batch {B} is not divisible by world size {world}")
    per = B // world
    return t[rank * per:(rank + 1) * per]


class FlatGradBuckets:
    """All trainable parameters' .grad are views into one flat fp32 buffer (reverse parameter order, every
    slot 128-byte aligned so the wgrad / LayerNorm-backward epilogues can red.add vectors straight into it:
    functional.SINKS). Post-accumulate hooks — or the sink notification, for gradients the kernels wrote
    in place — launch an async all-reduce per bucket as it completes."""

    ALIGN = 32  # elements (128 B)

    def __init__(self, params, bucket_bytes: int = 32 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        order = list(reversed(self.params))
        pad = lambda n: (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        total = sum(pad(p.numel()) for p in order)
        dev = order[0].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.bucket_of = {}
        self.buckets = []  # [start, end, n_params]
        self.sink_views = {}   # parameter storage address -> its .grad view
        self._by_key = {}
        off, b_start, b_count = 0, 0, 0
        for p in order:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self.bucket_of[p] = len(self.buckets)
            if p.is_contiguous():
                self.sink_views[p.data_ptr()] = p.grad
                self._by_key[p.data_ptr()] = p
            off += pad(n)
            b_count += 1
            if (off - b_start) * 4 >= bucket_bytes:
                self.buckets.append([b_start, off, b_count])
                b_start, b_count = off, 0
        if b_count:
            self.buckets.append([b_start, off, b_count])
        self.pending = [b[2] for b in self.buckets]
        self.handles = []
        self.hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params] \
            if self.world > 1 else []

    def sink_notify(self, key):
        """A backward kernel accumulated this parameter's gradient in place (no autograd hook will fire)."""
        if self.world > 1:
            self._on_grad(self._by_key[key])

    def _on_grad(self, p):
        b = self.bucket_of[p]
        self.pending[b] -= 1
        if self.pending[b] == 0:
            s, e, _ = self.buckets[b]
            self.handles.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.AVG, group=self.group,
                                                async_op=True))

    def zero(self):
        """Start of a step: clear the flat buffer (grads stay views into it; never set_to_none)."""
        self.flat.zero_()
        self.pending = [b[2] for b in self.buckets]
        self.handles = []

    def finish(self):
        """After backward: wait for the in-flight bucket all-reduces (the current stream waits, not the host)."""
        if self.world > 1 and any(n != 0 for n in self.pending):
            # parameters that received no gradient this step never fired their hook: reduce what is left
            for b, n in enumerate(self.pending):
                if n != 0:
                    s, e, _ = self.buckets[b]
                    self.handles.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.AVG, group=self.group,
                                                        async_op=True))
        for h in self.handles:
            h.wait()
        self.handles = []

    def close(self):
        for h in self.hooks:
            h.remove()


class DataParallelTrainer:
    """model: any nn.Module (the drop-in ViT / NeuroEncoder on GPU; a plain torch module in the CPU gloo
    tests). step(inputs, labels) runs forward, CrossEntropy, backward with overlapped bucketed all-reduce
    and the optimizer step on this rank's shard, and returns the (local) loss tensor without syncing."""

    def __init__(self, model, optimizer=None, lr=1e-4, weight_decay=0.01, bucket_mb=32, group=None):
        self.model = model
        self.buckets = FlatGradBuckets(list(model.parameters()), bucket_mb << 20, group)
        params = self.buckets.params
        if optimizer is None:
            fused = params[0].is_cuda
            optimizer = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, fused=fused)
        self.optimizer = optimizer
        self.criterion = torch.nn.CrossEntropyLoss()

    def step(self, inputs, labels):
        self.buckets.zero()
        out = self.model(inputs)
        loss = self.criterion(out, labels)
        with SINKS.active(self.buckets.sink_views, self.buckets.sink_notify):
            loss.backward()
        self.buckets.finish()
        self.optimizer.step()
        return loss

    @torch.no_grad()
    def predict(self, inputs):
        """Inference: each rank evaluates its own shard, no communication."""
        return self.model(inputs)
