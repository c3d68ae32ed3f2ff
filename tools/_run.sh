cd $GRAFT_REPO_ROOT
timeout 200 python tools/attn_probe.py --dropout 0.1 > gpurun_out/r2_attn_probe9.log 2>&1; echo "probe exit $?" >> gpurun_out/r2_attn_probe9.log
grep -v "OK$" gpurun_out/r2_attn_probe9.log | tail -6
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention or fmri" > gpurun_out/r2_attn_tests9.log 2>&1; tail -n 3 gpurun_out/r2_attn_tests9.log
timeout 100 python tools/fmri_probe.py 2>&1 | tail -3
NEUROVIT_LIB=$PWD/neurovit_b200/libneurovit_b200_prof.so timeout 100 python tools/attn_phases.py 2>&1 | tail -12
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_d.log 2>&1; tail -n 1 gpurun_out/r2_bench_d.log | cut -c1-3000
