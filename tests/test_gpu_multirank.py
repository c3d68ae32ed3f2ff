"""On-hardware multi-rank correctness (SURVEY §4 'distributed'): 2 ranks on 2 GPUs over real NCCL, through the graphed
step bench.py runs. Skipped with fewer than 2 GPUs (the driver's 1-GPU round-end run; run it with gpurun --gpus 2).
The gloo world_size-2 tests in tests/test_trainer_cpu.py cover the host logic on CPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs 2 CUDA devices", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env,expect", [
    ({}, "graph=1 two_graphs=0 own_nccl=1"),                                   # the default at 2 ranks: NCCL inside ONE graph
    ({"DP_TEST_GRAPH": "0"}, "graph=0 two_graphs=0 own_nccl=1"),                # eager, bucketed, own communicator
    ({"NEUROVIT_DP_NCCL": "torch"}, "graph=1 two_graphs=1 own_nccl=0"),         # torch.distributed: two graphs + one all-reduce
    ({"NEUROVIT_DP_NCCL": "torch", "DP_TEST_GRAPH": "0"}, "graph=0 two_graphs=0 own_nccl=0"),
])
def test_two_rank_gradients_match_full_batch(env, expect):
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_dp_worker.py")]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env={**os.environ, **env})
    except subprocess.TimeoutExpired as e:
        out = (e.stdout or b"").decode("utf-8", "replace")[-3000:] + (e.stderr or b"").decode("utf-8", "replace")[-3000:]
        raise AssertionError("2-rank worker timed out; output so far:\n" + out)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DP_OK " + expect in r.stdout, r.stdout[-2000:]
