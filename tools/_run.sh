cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "env0 or env1" > gpurun_out/r2_multirank.log 2>&1; tail -n 3 gpurun_out/r2_multirank.log | cut -c1-300
run() { echo "--- $1"; env $1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary --no-kernel-events > gpurun_out/r2_bench_n2_$2.log 2>&1; tail -n 1 gpurun_out/r2_bench_n2_$2.log | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["e2e"]["value"])
except Exception as e: print("ERR", e)'; }
echo "--- N=1"; timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --skip-cpu-baseline --no-kernel-events 2>/dev/null | tail -n 1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])'
run "NEUROVIT_NCCL_MAX_CTAS=4 NEUROVIT_SM_RESERVE=4" c4r4
run "NEUROVIT_NCCL_MAX_CTAS=8 NEUROVIT_SM_RESERVE=8" c8r8
run "NEUROVIT_NCCL_MAX_CTAS=2 NEUROVIT_SM_RESERVE=2" c2r2
run "NEUROVIT_NCCL_MAX_CTAS=0 NEUROVIT_SM_RESERVE=4" c0r4
