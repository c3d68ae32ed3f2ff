# Per-kernel evidence for the default bench: (1) torch.profiler step profile at the bench's dropout, (2) the ncu launch
# list of the same bench command (run only after that command exited 0 without ncu), (3) a full capture of the
# out-projection forward GEMM (store mode 3).   Usage (GPU box): bash tools/profile_step.sh -> gpurun_out/
set -u
timeout 200 python tools/step_profile.py --out gpurun_out/step_profile_p01.md > gpurun_out/step_profile.log 2>&1; tail -n 3 gpurun_out/step_profile.log | cut -c1-200
timeout 200 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1200 --csv --log-file gpurun_out/bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -n 1 gpurun_out/bench_plain.log | cut -c1-200
wc -l gpurun_out/bench_launches.csv
bash tools/ncu_gemm_res.sh 2>&1 | tail -n 3 | cut -c1-200
