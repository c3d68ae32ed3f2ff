cd $GRAFT_REPO_ROOT
timeout 200 python tools/attn_probe.py --dropout 0.1 > gpurun_out/r2_attn_probe12.log 2>&1; echo "probe exit $?" >> gpurun_out/r2_attn_probe12.log
grep -v "OK$" gpurun_out/r2_attn_probe12.log | tail -6
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" | tail -n 2
NEUROVIT_LIB=$PWD/neurovit_b200/libneurovit_b200_prof.so timeout 100 python tools/attn_phases.py 2>&1 | tail -11 | head -8
bash tools/ncu_step_kernels.sh
