cd $GRAFT_REPO_ROOT
run() { echo "--- $1"; env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 --no-secondary --no-kernel-events > gpurun_out/r2_bench_n8_$2.log 2>&1; tail -n 1 gpurun_out/r2_bench_n8_$2.log | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["e2e"]["value"])
except Exception as e: print("ERR", e)'; }
run "NEUROVIT_NCCL_MAX_CTAS=4 NEUROVIT_SM_RESERVE=4" cfg4r4
run "NEUROVIT_NCCL_MAX_CTAS=8 NEUROVIT_SM_RESERVE=8" cfg8r8
run "NEUROVIT_DP_NCCL=torch" torch2
