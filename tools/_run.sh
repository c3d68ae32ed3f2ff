cd $GRAFT_REPO_ROOT
timeout 100 python tools/dbg_sinks.py 2>&1 | grep -v Warn | grep -v "^order" | tail -4
timeout 200 python tools/attn_probe.py --dropout 0.1 > gpurun_out/r2_attn_probe8.log 2>&1; echo "probe exit $?" >> gpurun_out/r2_attn_probe8.log
grep -v "OK$" gpurun_out/r2_attn_probe8.log | tail -8
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k attention > gpurun_out/r2_attn_tests8.log 2>&1; tail -n 5 gpurun_out/r2_attn_tests8.log
export NEUROVIT_LIB=$PWD/neurovit_b200/libneurovit_b200_prof.so
timeout 100 python tools/attn_phases.py 2>&1 | tail -12
