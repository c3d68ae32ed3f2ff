# scratch: the command list of the last gpurun call (tools/gpurun_retry.sh -- 'bash tools/_run.sh')
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q | tail -n 3
timeout 300 python bench.py --steps 20 --warmup 5 | tail -n 1 | cut -c1-400
