cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q | tail -n 2
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --no-secondary --no-kernel-events | tail -n 1 | cut -c1-200
timeout 200 python tools/step_profile.py 2>/dev/null | grep -E "ln_bwd_pipe|wall" | cut -c1-200
