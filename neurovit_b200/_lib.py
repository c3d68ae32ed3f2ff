"""ctypes binding of the C-ABI library ``libneurovit_b200.so`` (see include/neurovit_b200.h).

There is deliberately no CPU or PyTorch fallback: if the shared library is missing, or the device is
not a B200 (sm_100), every entry point raises. ``oracle/`` is test infrastructure and is never imported
from here.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# NEUROVIT_LIB: load an instrumented build instead (the profiling tools; see _build.build)
LIB_PATH = os.environ.get("NEUROVIT_LIB") or os.path.join(_HERE, "libneurovit_b200.so")

_i = ctypes.c_int
_l = ctypes.c_int64
_f = ctypes.c_float
_p = ctypes.c_void_p

# name -> argtypes; must mirror include/neurovit_b200.h exactly (tests/test_abi.py checks the symbol list)
SIGNATURES = {
    "nv_version": [],
    "nv_device_check": [],
    "nv_gemm_bf16": [_i, _i, _i, _i, _i, _p, _l, _p, _l, _p, _p, _l, _p, _l, _p, _l, _p, _l, _p, _l, _p,
                     _i, _i, _f, _i, _i, _i, _f, _l, _i, _p, _i, _p],
    "nv_head_fwd": [_p, _l, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p],
    "nv_head_bwd": [_p, _p, _l, _p, _p, _p, _p, _p, _p, _l, _p, _l, _p, _p, _p, _p, _i, _i, _i, _p],
    "nv_dropout_bits": [_p, _l, _f, _l, _i, _p],
    "nv_adamw_flat": [_p, _p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _i, _p, _p],
    "nv_counter_add": [_p, _f, _p],
    "nv_rng_epoch_advance": [_p],
    "nv_rng_epoch_get": [_p, _p],
    "nv_dropout": [_p, _l, _p, _l, _p, _l, _p, _l, _p, _i, _i, _f, _l, _i, _i, _p],
    "nv_gemm_f32": [_i, _i, _i, _i, _i, _p, _l, _l, _l, _l, _p, _l, _l, _l, _l, _p, _l, _l, _l,
                    _p, _p, _l, _p, _l, _p, _l, _i, _i, _f, _p],
    "nv_layernorm_fwd": [_p, _l, _i, _i, _i, _p, _p, _p, _l, _i, _i, _p, _i, _l, _i, _i, _i, _p, _p,
                         _i, _i, _f, _p],
    "nv_layernorm_bwd": [_p, _i, _l, _i, _i, _i, _p, _l, _i, _i, _i, _p, _p, _p, _p, _l, _p, _l, _i, _i, _i,
                         _p, _l, _p, _p, _p, _i, _i, _f, _l, _i, _p, _p],
    "nv_cls_row": [_p, _p, _p, _l, _i, _i, _p],
    "nv_patch_gather_ln": [_p, _p, _p, _p, _p, _p, _p, _i, _l, _p, _p, _p, _f, _p],
    "nv_ln_fold": [_p, _p, _p, _p, _p, _i, _l, _p, _i, _i, _p],
    "nv_ln_fold_grads": [_p, _l, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "nv_patch_ln_param_grad": [_p, _p, _p, _p, _p, _l, _p, _p, _p, _p, _p],
    "nv_attention_fwd": [_p, _p, _p, _l, _l, _p, _l, _l, _p, _i, _i, _i, _i, _f, _f, _l, _p, _i, _p],
    "nv_attention_bwd": [_p, _p, _p, _l, _l, _p, _p, _l, _l, _p, _p, _p, _p, _p, _l, _l,
                         _i, _i, _i, _i, _f, _f, _p, _p],
    "nv_attention_cls_fwd": [_p, _p, _p, _l, _l, _p, _l, _p, _i, _i, _i, _i, _f, _f, _l, _p, _i, _p],
    "nv_attention_cls_bwd": [_p, _p, _p, _l, _l, _p, _l, _p, _l, _p, _p, _p, _p, _l, _l, _i, _i, _i, _i, _f, _f, _p, _p],
    "nv_softmax_fwd": [_p, _l, _i, _p],
    "nv_softmax_bwd": [_p, _p, _l, _i, _p],
    "nv_cast_f32_bf16": [_p, _p, _l, _p],
    "nv_cast_transpose_f32_bf16": [_p, _p, _p, _i, _i, _p],
    "nv_colsum": [_p, _i, _l, _p, _i, _i, _p],
    "nv_batch_sum": [_p, _l, _p, _i, _l, _p],
    "nv_mean_pool_fwd": [_p, _p, _i, _i, _i, _p],
    "nv_mean_pool_bwd": [_p, _p, _p, _i, _i, _i, _p],
    "nv_temporal_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _f, _f, _f, _l, _p],
    "nv_temporal_bwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _f, _f, _f, _l, _p],
    "nv_fmri_deinterleave": [_p, _p, _i, _l, _i, _p, ctypes.c_double, _p],
    "nv_set_sm_reserve": [_i],
    "nv_dp_load": [_p],
    "nv_dp_nccl_version": [],
    "nv_dp_unique_id": [_p],
    "nv_dp_init": [_p, _i, _i, _i],
    "nv_dp_register": [_p, _l],
    "nv_dp_allreduce_bucket": [_p, _l, _i, _i, _p],
    "nv_dp_world": [_p, _p],
    "nv_dp_destroy": [],
}

_lock = threading.Lock()
_lib = None
_device_ok = set()


class NeuroViTLibraryError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (once). Raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NeuroViTLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). neurovit_b200 has no CPU / PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
        lib.nv_last_error.argtypes = []
        lib.nv_last_error.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().nv_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str) -> None:
    if status != 0:
        msg = last_error()
        if status == 1:
            raise ValueError(f"{what}: {msg}")
        raise NeuroViTLibraryError(f"{what} failed (status {status}): {msg}")


def require_device(device_index: int) -> None:
    """Fail loudly unless the given CUDA device is sm_100 (B200)."""
    if device_index in _device_ok:
        return
    import torch

    with torch.cuda.device(device_index):
        check(load().nv_device_check(), "nv_device_check")
    _device_ok.add(device_index)


class _LaunchCounter:
    """Counts the CUDA kernels launched through the C ABI (bench.py reports it as `gpu_launches`)."""
    KERNELS_PER_CALL = {"nv_attention_bwd": 2, "nv_version": 0, "nv_rng_epoch_get": 0, "nv_device_check": 0, "nv_dp_load": 0, "nv_set_sm_reserve": 0,
                        "nv_dp_nccl_version": 0, "nv_dp_unique_id": 0, "nv_dp_init": 0, "nv_dp_register": 0,
                        "nv_dp_world": 0, "nv_dp_destroy": 0}

    def __init__(self):
        self.count = 0

    def reset(self):
        self.count = 0


LAUNCHES = _LaunchCounter()


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
    LAUNCHES.count += _LaunchCounter.KERNELS_PER_CALL.get(name, 1)
