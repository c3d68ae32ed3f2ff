cd $GRAFT_REPO_ROOT
timeout 150 python tools/attn_probe.py --dropout 0.1 > gpurun_out/r2_attn_probe7.log 2>&1; echo "probe exit $?" >> gpurun_out/r2_attn_probe7.log
grep -v "OK$" gpurun_out/r2_attn_probe7.log | tail -8
run() { echo "--- $1"; env $1 timeout 60 python tools/attn_probe.py --time-only 64 385 8 2>&1 | grep "fwd"; }
run "NV_ATTN_ONLY=dq"
run "NV_ATTN_ONLY=dkv"
export NEUROVIT_LIB=$PWD/neurovit_b200/libneurovit_b200_prof.so
timeout 100 python tools/attn_phases.py 2>&1 | tail -4
