cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_d.log 2>&1; tail -n 3 gpurun_out/r2_gputests_d.log
bash tools/fusion_ab.sh > gpurun_out/r02_fusion_ab.txt 2>&1; cat gpurun_out/r02_fusion_ab.txt
