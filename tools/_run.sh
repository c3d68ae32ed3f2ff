cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_a.log 2>&1; tail -n 3 gpurun_out/r2_gputests_a.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_a.log 2>&1; tail -n 1 gpurun_out/r2_bench_a.log | cut -c1-1500
timeout 200 python tools/step_profile.py --out gpurun_out/r2_step_profile_a.md > gpurun_out/r2_step_profile_a.log 2>&1; tail -n 3 gpurun_out/r2_step_profile_a.log | cut -c1-200
timeout 200 python bench.py --impl torch_gpu --steps 10 --warmup 3 > gpurun_out/r2_torchgpu_a.log 2>&1; tail -n 1 gpurun_out/r2_torchgpu_a.log | cut -c1-900
