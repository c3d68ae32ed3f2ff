"""Host side of the data-parallel exchange step: one NCCL communicator per process, owned by the C-ABI library
(csrc/nv_dp.cu, nv_dp_*), bootstrapped through whatever torch.distributed process group the caller already has
(only to hand rank 0's 128-byte unique id to the other ranks — gloo or nccl, it does not matter).

Why not torch.distributed's own all_reduce: ProcessGroupNCCL's collectives could not be captured into the training
step's CUDA graph here (the capture hung at world_size 2, round 1), which forced the step into two graphs around one
exposed all-reduce of the whole 155 MB gradient buffer. ncclAllReduce issued directly on the capturing stream is a
plain kernel launch: it is captured like any other node, per gradient bucket, on a side stream, under backward."""
from __future__ import annotations

import ctypes
import glob
import os

import torch
import torch.distributed as dist

from . import _lib


def _find_nccl() -> str:
    """The libnccl.so.2 PyTorch itself uses (nvidia-nccl wheel, or bundled under torch/lib)."""
    env = os.environ.get("NEUROVIT_NCCL_LIB")
    if env:
        return env
    sp = os.path.dirname(os.path.dirname(torch.__file__))
    for pat in (os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so.2"),
                os.path.join(os.path.dirname(torch.__file__), "lib", "libnccl.so.2")):
        hits = glob.glob(pat)
        if hits:
            return hits[0]
    return ""  # let the dynamic loader search


class NcclComm:
    """`NcclComm(group)` creates the library's communicator for this rank on the current CUDA device."""

    # CTAs this communicator's collectives may use (env NEUROVIT_NCCL_MAX_CTAS; ncclConfig_t.maxCTAs — NOT the
    # NCCL_MAX_CTAS environment variable, which NCCL reads once per process and PyTorch's own communicator has usually
    # consumed already): the all-reduce runs UNDER backward, where every SM it takes is taken from the persistent GEMMs
    MAX_CTAS = int(os.environ.get("NEUROVIT_NCCL_MAX_CTAS", "4"))

    def __init__(self, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NcclComm needs an initialised torch.distributed process group to exchange the NCCL id")
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        path = _find_nccl()
        _lib.call("nv_dp_load", ctypes.c_char_p(path.encode()) if path else None)
        uid = ctypes.create_string_buffer(128)
        if self.rank == 0:
            _lib.call("nv_dp_unique_id", uid)
        box = [bytes(uid.raw)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._uid = ctypes.create_string_buffer(box[0], 128)
        _lib.call("nv_dp_init", self._uid, self.rank, self.world, self.MAX_CTAS)
        self.version = _lib.load().nv_dp_nccl_version()
        self._alive = True

    def register(self, t: torch.Tensor):
        _lib.call("nv_dp_register", ctypes.c_void_p(t.data_ptr()), t.numel() * t.element_size())

    def all_reduce_(self, t: torch.Tensor, average: bool = True):
        """In place, on the current stream; t fp32 or bf16, contiguous."""
        assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.bfloat16)
        _lib.call("nv_dp_allreduce_bucket", ctypes.c_void_p(t.data_ptr()), t.numel(),
                  0 if t.dtype == torch.float32 else 1, 1 if average else 0,
                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))

    def close(self):
        if self._alive:
            self._alive = False
            _lib.call("nv_dp_destroy")
