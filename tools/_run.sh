cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4.log 2>&1; tail -n 1 gpurun_out/r02_bench_n4.log | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"].get("config3_global512"), d["clocks"])
except Exception as e: print("ERR", e)'
tail -n 3 gpurun_out/r02_bench_n4.log | cut -c1-300
