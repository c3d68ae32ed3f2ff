"""In-tree build of libneurovit_b200.so (nvcc, sm_100a only).

The shared library is a plain C-ABI object (no libtorch, no pybind): it is loaded with ctypes by
``neurovit_b200._lib``. Objects are rebuilt only when a source or header is newer than the object.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
REPO = os.path.dirname(HERE)
BUILD_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(HERE, "libneurovit_b200.so")

SOURCES = [
    "nv_host.cu",
    "gemm_tc.cu",
    "simt_gemm.cu",
    "layernorm.cu",
    "patch_embed.cu",
    "attention_tc.cu",
    "misc.cu",
    "temporal.cu",
    "fmri4d.cu",
    "nv_dp.cu",
    "api.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; neurovit_b200 needs the CUDA 12.9 toolchain to build")
    return cand


def _deps_mtime() -> float:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(REPO, "include", "neurovit_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs if os.path.exists(h))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path.
    NV_PROFILE=1 / NV_DEBUG_PROGRESS=1 build an instrumented variant next to the product library
    (libneurovit_b200_prof.so, objects under build/prof/); NV_VARIANT=<name> NV_DEFINES="-D..." builds an A/B variant
    (libneurovit_b200_<name>.so). Load either with NEUROVIT_LIB=<path>."""
    nvcc = _nvcc()
    hdr_mtime = _deps_mtime()
    extra = ["-DNV_PROFILE"] if os.environ.get("NV_PROFILE") == "1" else []  # in-kernel phase clocks (tools/attn_phases.py)
    if os.environ.get("NV_DEBUG_PROGRESS") == "1":  # progress markers into a host-mapped buffer
        extra.append("-DNV_DEBUG_PROGRESS")
    build_dir, lib_path = (os.path.join(BUILD_DIR, "prof"), LIB_PATH.replace(".so", "_prof.so")) if extra else (BUILD_DIR, LIB_PATH)
    variant = os.environ.get("NV_VARIANT")   # A/B builds: NV_VARIANT=g1 NV_DEFINES="-DNV_SIMT_GROUPS=1" -> libneurovit_b200_g1.so
    if variant:
        extra += os.environ.get("NV_DEFINES", "").split()
        build_dir, lib_path = os.path.join(BUILD_DIR, variant), LIB_PATH.replace(".so", f"_{variant}.so")
    os.makedirs(build_dir, exist_ok=True)
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(build_dir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_mtime):
            jobs.append([nvcc, *NVCC_FLAGS, *extra, "-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    need_link = force or bool(jobs) or not os.path.exists(lib_path) or any(
        os.path.getmtime(o) > os.path.getmtime(lib_path) for o in objs)
    if need_link:
        run([nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
             "-Xcompiler", "-fPIC", "-cudart", "static", "-ldl"])
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
