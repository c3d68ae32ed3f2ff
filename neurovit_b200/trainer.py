"""Data-parallel training step for the ViT3D hot path: one process per GPU, batch sharding at the step
boundary and a bucketed gradient all-reduce overlapped with backward.

Reference: src/Trainer.py:65-76 is the single-GPU step (forward, CrossEntropy, zero_grad, backward, AdamW
step; fp16 autocast + GradScaler there, bf16 operands with fp32 master weights here — no loss scaling
needed). The reference has no distributed code; data parallelism is the new work BASELINE.json
asks for: rank r takes batch[r*B/g:(r+1)*B/g]; gradients live in ONE flat fp32 buffer laid out in reverse
parameter order (the order backward produces them), cut into buckets; each bucket is all-reduced (NCCL,
average) as soon as its last gradient has landed, on a side stream, while the rest of backward still runs.
On CUDA the all-reduce is issued through the library's own communicator (csrc/nv_dp.cu, neurovit_b200/dp.py): a
plain kernel launch on the side stream, so the whole multi-rank step — forward, backward, the per-bucket
all-reduces and FlatAdamW — is ONE CUDA graph. CPU tensors (the gloo tests) and NEUROVIT_DP_NCCL=torch go through
torch.distributed's all_reduce instead. Inference shards the batch with no communication.
"""
from __future__ import annotations

import os

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

    def __init__(self, params, bucket_bytes: int = 32 << 20, group=None, flatten_params: bool = False, comm=None):
        """flatten_params: also move the parameters themselves into one flat fp32 buffer with the same slot layout
        (p.data becomes a view of it; values, names and state_dict are unchanged) so FlatAdamW can update
        everything with one kernel."""
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.comm = comm            # dp.NcclComm: all-reduces are kernel launches on self.comm_stream (graph-capturable)
        self.comm_stream = torch.cuda.Stream() if comm is not None else None
        self._comm_used = False
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        order = list(reversed(self.params))
        pad = lambda n: (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        total = sum(pad(p.numel()) for p in order)
        dev = order[0].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_params = None
        if flatten_params:
            self.flat_params = torch.zeros(total, device=dev, dtype=torch.float32)
            off = 0
            with torch.no_grad():
                for p in order:
                    n = p.numel()
                    self.flat_params[off:off + n].copy_(p.detach().reshape(-1))
                    p.data = self.flat_params[off:off + n].view_as(p)
                    off += pad(n)
        self.slot_of = {}      # parameter -> (offset, numel) in the flat buffers
        self.bucket_of = {}
        self.buckets = []  # [start, end, n_params]
        self.sink_views = {}   # parameter storage address -> its .grad view
        self._by_key = {}
        off, b_start, b_count = 0, 0, 0
        for p in order:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self.slot_of[p] = (off, n)
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
        self.seen = [set() for _ in self.buckets]   # parameters of each bucket whose gradient has landed this step
        self.fired = [False] * len(self.buckets)
        self.handles = []
        self._echo = {}
        self.hooks = [p.register_post_accumulate_grad_hook(self._on_hook) for p in self.params] \
            if self.world > 1 else []
        if self.world > 1:
            self.broadcast_from_rank0()

    def broadcast_from_rank0(self):
        """Replicas must start identical: rank 0's parameters (and buffers, by the caller) win, as DDP does at
        construction. Cheap insurance against per-rank seeds or a checkpoint loaded on one rank only."""
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        with torch.no_grad():
            if self.flat_params is not None:
                dist.broadcast(self.flat_params, src=src, group=self.group)
            else:
                for p in self.params:
                    dist.broadcast(p.data, src=src, group=self.group)

    def sink_notify(self, key):
        """A backward kernel accumulated this parameter's gradient in place. Autograd still runs the parameter's
        AccumulateGrad node afterwards (with an undefined gradient: the Function returned None) and with it the
        post-accumulate hook — that echo carries no new gradient and is swallowed in _on_hook."""
        if self.world > 1:
            p = self._by_key[key]
            self._echo[id(p)] = self._echo.get(id(p), 0) + 1
            self._on_grad(p)

    def _on_hook(self, p):
        n = self._echo.get(id(p), 0)
        if n > 0:
            self._echo[id(p)] = n - 1
            return
        self._on_grad(p)

    defer = False  # True: no all-reduce is launched from the backward thread; finish() reduces the whole buffer

    def _reduce_bucket(self, b):
        s, e, _ = self.buckets[b]
        self.fired[b] = True
        if self.comm is not None:
            # fork: the side stream waits for everything the backward stream has enqueued so far (inside a capture
            # this event becomes a graph edge), then runs ncclAllReduce as an ordinary kernel launch
            ev = torch.cuda.Event()
            ev.record()
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                self.comm.all_reduce_(self.flat[s:e], average=True)
            self._comm_used = True
        else:
            self.handles.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.AVG, group=self.group,
                                                async_op=True))

    def _on_grad(self, p):
        """One parameter's gradient is complete. A bucket is reduced when ALL its parameters have reported — tracked
        as a set, so a parameter that reports twice in one step (shared weights, two forwards before one backward)
        cannot release the bucket early; its later contributions would miss the reduction, so that case raises."""
        if self.defer:
            return
        b = self.bucket_of[p]
        if self.fired[b]:
            raise RuntimeError("a gradient arrived for a bucket that was already all-reduced in this step (parameter "
                               "used twice before one backward?); set buckets.defer = True to reduce once after backward")
        self.seen[b].add(id(p))
        if len(self.seen[b]) == self.buckets[b][2]:
            self._reduce_bucket(b)

    def zero(self):
        """Start of a step: clear the flat buffer (grads stay views into it; never set_to_none)."""
        self.flat.zero_()
        for sset in self.seen:
            sset.clear()
        self.fired = [False] * len(self.buckets)
        self.handles = []
        self._echo = {}

    def finish(self):
        """After backward: wait for the in-flight bucket all-reduces (the current stream waits, not the host)."""
        if self.world == 1:
            return
        if self.defer:
            if self.comm is not None:
                self.comm.all_reduce_(self.flat, average=True)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            return
        # parameters that received no gradient this step never reported: reduce what is left
        for b in range(len(self.buckets)):
            if not self.fired[b]:
                self._reduce_bucket(b)
        if self._comm_used:
            torch.cuda.current_stream().wait_stream(self.comm_stream)   # join (closes the fork inside a capture)
            self._comm_used = False
        for h in self.handles:
            h.wait()
        self.handles = []

    def close(self):
        for h in self.hooks:
            h.remove()
        if self.comm is not None:
            self.comm.close()


class FlatAdamW:
    """torch.optim.AdamW semantics (src/Trainer.py:31: lr, weight_decay; default betas / eps) as ONE kernel over the
    trainer's flat parameter / gradient / moment buffers (nv_adamw_flat), which also rewrites the bf16 weight copies
    the next forward's tensor-core GEMMs read. CUDA only."""

    def __init__(self, buckets: FlatGradBuckets, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        from . import ops
        from .functional import engine
        assert buckets.flat_params is not None, "FlatAdamW needs FlatGradBuckets(flatten_params=True)"
        self.buckets = buckets
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.m = torch.zeros_like(buckets.flat)
        self.v = torch.zeros_like(buckets.flat)
        self.t_dev = torch.zeros(1, device=buckets.flat.device, dtype=torch.float32)  # step count, device-side:
        # the kernel reads it, so a CUDA graph that replays step() keeps the bias corrections moving
        self._ops = ops
        self.shadow = torch.empty(buckets.flat.numel(), device=buckets.flat.device, dtype=torch.bfloat16)
        ops.cast_bf16(buckets.flat_params, out=self.shadow)
        self._wc = engine("bf16").wc
        for p, (off, n) in buckets.slot_of.items():
            if p.dim() == 2:  # GEMM weights: their bf16 operand copy lives in the shadow buffer from now on
                self._wc.pin(p, self.shadow[off:off + n].view_as(p))

    @property
    def t(self):
        return int(self.t_dev.item())

    def step(self):
        b = self.buckets
        self._ops.counter_add(self.t_dev, 1.0)
        self._ops.adamw_flat(b.flat_params, b.flat, self.m, self.v, self.shadow, lr=self.lr, beta1=self.betas[0],
                             beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay, step=1,
                             step_dev=self.t_dev)
        self._wc.epoch += 1  # derived copies that are not pinned (K-padded patch weights) must be re-cast

    def zero_grad(self, set_to_none=False):
        self.buckets.zero()

    def state_dict(self):
        return {"t": self.t, "m": self.m, "v": self.v, "lr": self.lr, "betas": self.betas, "eps": self.eps,
                "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        """Restores step count, moments AND hyper-parameters (torch.optim semantics), then re-casts the bf16 weight
        copies from the current fp32 masters. A CUDA graph captured earlier baked the old hyper-parameters in as
        kernel arguments: DataParallelTrainer.load_state_dict drops it so the next step re-captures."""
        self.t_dev.fill_(float(sd["t"]))
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.lr = sd.get("lr", self.lr)
        self.betas = tuple(sd.get("betas", self.betas))
        self.eps = sd.get("eps", self.eps)
        self.weight_decay = sd.get("weight_decay", self.weight_decay)
        self.refresh_shadow()

    def refresh_shadow(self):
        """fp32 masters changed behind the optimizer's back (load_state_dict, manual edit): re-cast the bf16 copies."""
        self._ops.cast_bf16(self.buckets.flat_params, out=self.shadow)
        self._wc.epoch += 1


class DataParallelTrainer:
    """model: any nn.Module (the drop-in ViT / NeuroEncoder on GPU; a plain torch module in the CPU gloo
    tests). step(inputs, labels) runs forward, CrossEntropy, backward with overlapped bucketed all-reduce
    and the optimizer step on this rank's shard, and returns the (local) loss tensor without syncing."""

    def __init__(self, model, optimizer=None, lr=1e-4, weight_decay=0.01, bucket_mb=None, group=None, graph=False):
        """graph=True: the whole step (forward, loss, backward, all-reduce, optimizer) is captured into ONE CUDA graph
        on its first call and replayed afterwards — the ~300 kernel launches of a step then cost one launch on the
        host and ~1 us instead of ~3 us of idle GPU time each. Needs static shapes (same batch size every step), the
        built-in FlatAdamW optimizer and CUDA tensors; inputs are copied into static buffers. Dropout masks still
        change every step (device-side epoch counter, nv_rng_epoch_advance). Ignored (eager launches) when the
        process group has more than one rank."""
        if bucket_mb is None:
            # ~ half a transformer layer: the tail of backward (layer 0's attention, the patch embedding) then leaves
            # only a few MB to reduce after the last kernel
            bucket_mb = int(os.environ.get("NEUROVIT_BUCKET_MB", "12"))
        self.model = model
        plist = list(model.parameters())
        own_adamw = optimizer is None and plist[0].is_cuda and os.environ.get("NEUROVIT_TORCH_ADAMW") != "1"
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        comm = None
        # Which all-reduce: "own" = the library's communicator, buckets reduced under backward inside the step's graph;
        # "torch" = torch.distributed, one all-reduce of the whole buffer between two graphs. Measured (profiles/
        # r02_summary.md §6): at 2 ranks the overlapped path wins (9.34 vs 9.45 ms/step); at 8 ranks NCCL confined to the
        # few CTAs that backward can spare is too slow to finish inside backward and the exposed full-speed all-reduce
        # wins (9.54 vs 9.87). "auto" (default) picks accordingly.
        which = os.environ.get("NEUROVIT_DP_NCCL", "auto")
        if which == "auto":
            which = "own" if world == 2 else "torch"
        if world > 1 and plist[0].is_cuda and which == "own":
            from . import dp, ops
            comm = dp.NcclComm(group)
        self.buckets = FlatGradBuckets(plist, bucket_mb << 20, group, flatten_params=own_adamw, comm=comm)
        if comm is not None:
            comm.register(self.buckets.flat)
        if world > 1:
            for buf in model.buffers():
                dist.broadcast(buf.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        params = self.buckets.params
        if own_adamw:
            optimizer = FlatAdamW(self.buckets, lr=lr, weight_decay=weight_decay)
        elif optimizer is None:
            fused = params[0].is_cuda
            optimizer = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, fused=fused)
        self.optimizer = optimizer
        self.criterion = torch.nn.CrossEntropyLoss()
        # Several ranks: with the library's own communicator the bucketed all-reduces are captured into the step's
        # graph. Through torch.distributed (NEUROVIT_DP_NCCL=torch) NCCL stays out of the capture (ProcessGroupNCCL
        # inside a capture hung at world_size 2): two graphs around ONE eager all-reduce of the whole buffer.
        # NEUROVIT_GRAPH_DP=0 keeps multi-rank steps eager.
        self.use_graph = bool(graph) and (self.buckets.world == 1 or os.environ.get("NEUROVIT_GRAPH_DP", "1") == "1")
        self._two_graphs = self.use_graph and self.buckets.world > 1 and comm is None
        if self._two_graphs or os.environ.get("NEUROVIT_DP_DEFER") == "1":
            self.buckets.defer = True
        if self.use_graph and not isinstance(optimizer, FlatAdamW):
            raise ValueError("graph=True needs the built-in FlatAdamW optimizer (CUDA parameters, optimizer=None)")
        self._graph = self._graph2 = self._graph_alt = None
        self._flip = False
        self._cuda = plist[0].is_cuda
        # NCCL's CTAs need SMs of their own while the bucket all-reduces run under backward: the persistent kernels
        # (GEMM, LayerNorm backward) launched during backward leave that many SMs free — a persistent grid that asked
        # for every SM would have some of its CTAs queued behind the NCCL kernel for a whole extra wave.
        self._sm_reserve = 0
        if comm is not None and not self.buckets.defer:
            from . import dp
            self._sm_reserve = int(os.environ.get("NEUROVIT_SM_RESERVE", str(dp.NcclComm.MAX_CTAS)))

    def _fwd_bwd(self, inputs, labels):
        if self._cuda:
            from .functional import ZEROS
            ZEROS.reset()   # accumulators of this step come from chunks filled inside this step (or this capture)
        self.buckets.zero()
        out = self.model(inputs)
        loss = self.criterion(out, labels)
        if self._sm_reserve:
            from . import ops
            ops.set_sm_reserve(self._sm_reserve)
        try:
            with SINKS.active(self.buckets.sink_views, self.buckets.sink_notify):
                loss.backward()
        finally:
            if self._sm_reserve:
                ops.set_sm_reserve(0)
        return loss

    def _update(self):
        self.optimizer.step()
        if self._cuda:
            from . import ops
            ops.rng_epoch_advance()

    def _eager_step(self, inputs, labels):
        loss = self._fwd_bwd(inputs, labels)
        self.buckets.finish()
        self._update()
        return loss

    def _capture(self, inputs, labels):
        cur = torch.cuda.current_stream()
        self._sx, self._sy = torch.empty_like(inputs), torch.empty_like(labels)
        self._sx.copy_(inputs)
        self._sy.copy_(labels)
        opt = self.optimizer
        # warm-up steps (lazy initialisation, autotuned attributes, allocator) must not train: snapshot and restore
        snap = (self.buckets.flat_params.clone(), opt.m.clone(), opt.v.clone(), opt.t_dev.clone(), opt.shadow.clone())
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(3):
                self._eager_step(self._sx, self._sy)
        cur.wait_stream(side)
        for dst, src in zip((self.buckets.flat_params, opt.m, opt.v, opt.t_dev, opt.shadow), snap):
            dst.copy_(src)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.LAUNCHES.count
        self._graph = torch.cuda.CUDAGraph()
        if not self._two_graphs:
            mode = "thread_local" if self.buckets.world > 1 else "global"
            with torch.cuda.graph(self._graph, capture_error_mode=mode):
                self._sloss = self._eager_step(self._sx, self._sy)
            self._graph2 = None
            # The same executable graph cannot overlap its own previous launch: the relaunch of a ~200-node graph
            # leaves the GPU idle between replays. A second capture of the identical step (same static inputs and
            # persistent buffers, its own activations from the same pool) lets consecutive steps alternate between two
            # executables, so step i+1's launch is prepared while step i runs.
            self._graph_alt = None
            if os.environ.get("NEUROVIT_GRAPH_PINGPONG", "1") == "1":
                self._graph_alt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph_alt, pool=self._graph.pool(), capture_error_mode=mode):
                    self._sloss_alt = self._eager_step(self._sx, self._sy)
                n_per = (_lib.LAUNCHES.count - n0) // 2
                _lib.LAUNCHES.count = n0 + n_per
        else:
            # several ranks: NCCL stays OUT of the capture. Graph 1 = forward + backward into the flat gradient
            # buffer, then ONE eager all-reduce of that buffer, then graph 2 = optimizer + epoch advance.
            # (thread_local error mode: the process group's watchdog thread may touch the CUDA API meanwhile)
            with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
                self._sloss = self._fwd_bwd(self._sx, self._sy)
            self._graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph2, pool=self._graph.pool(), capture_error_mode="thread_local"):
                self._update()
        self._graph_launches = _lib.LAUNCHES.count - n0  # C-ABI kernels recorded in the graphs = launched per step
        _lib.LAUNCHES.count = n0

    def step(self, inputs, labels):
        if not self.use_graph:
            return self._eager_step(inputs, labels)
        if self._graph is None:
            try:
                self._capture(inputs, labels)
            except Exception as e:  # capture is an optimisation: report loudly and launch kernel by kernel instead
                import warnings
                warnings.warn(f"neurovit_b200: CUDA-graph capture of the training step failed ({type(e).__name__}: {e}); "
                              "continuing with eager launches")
                self.use_graph, self._graph = False, None
                torch.cuda.synchronize()
                return self._eager_step(inputs, labels)
        if inputs.shape != self._sx.shape or labels.shape != self._sy.shape:
            raise ValueError(f"graph=True needs static shapes: captured {tuple(self._sx.shape)}, got {tuple(inputs.shape)}")
        self._sx.copy_(inputs, non_blocking=True)
        self._sy.copy_(labels, non_blocking=True)
        loss = self._sloss
        if getattr(self, "_graph_alt", None) is not None and self._flip:
            self._graph_alt.replay()
            loss = self._sloss_alt
        else:
            self._graph.replay()
        self._flip = not self._flip
        if self._graph2 is not None:
            self.buckets.finish()   # deferred mode: one all-reduce (AVG) of the whole flat gradient buffer
            self._graph2.replay()
        from . import _lib
        _lib.LAUNCHES.count += self._graph_launches
        return loss

    def close(self):
        """Release the captured graphs BEFORE the communicator: NCCL keeps a reference per captured collective and
        ncclCommDestroy waits for them."""
        self.reset_graph()
        self._sloss = self._sloss_alt = None
        if self._cuda:
            torch.cuda.synchronize()
        self.buckets.close()

    def reset_graph(self):
        """Drop the captured step (new batch shape, changed hyper-parameters): the next step() captures again."""
        self._graph = self._graph2 = self._graph_alt = None
        self._flip = False

    def load_state_dict(self, model_sd=None, optimizer_sd=None):
        """Load a checkpoint into a live trainer: parameters go into the flat fp32 buffer (p.data are views of it),
        the bf16 weight copies are re-cast at once, and a captured graph — which baked the old hyper-parameters in
        and would replay the first step on stale bf16 copies — is dropped and re-captured on the next step."""
        if model_sd is not None:
            self.model.load_state_dict(model_sd)
        if optimizer_sd is not None:
            self.optimizer.load_state_dict(optimizer_sd)
        if isinstance(self.optimizer, FlatAdamW):
            self.optimizer.refresh_shadow()
        self.reset_graph()

    @torch.no_grad()
    def predict(self, inputs):
        """Inference: each rank evaluates its own shard, no communication."""
        return self.model(inputs)
