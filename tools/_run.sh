cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gputests_b.log 2>&1; tail -n 12 gpurun_out/r2_gputests_b.log
export NEUROVIT_LIB=$PWD/neurovit_b200/libneurovit_b200_prof.so
timeout 100 python tools/attn_phases.py > gpurun_out/r2_phases_p0.log 2>&1; cat gpurun_out/r2_phases_p0.log | tail -20
ATTN_DROPOUT=0.1 timeout 100 python tools/attn_phases.py > gpurun_out/r2_phases_p01.log 2>&1; cat gpurun_out/r2_phases_p01.log | tail -20
