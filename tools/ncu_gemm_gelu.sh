# ncu full capture (with source) of the bias+GELU forward GEMM (mlp-up), after the same command exited 0 without ncu.
C2='{"cg": 2, "a_mn": 0, "b_mn": 0, "block_n": 256, "M": 24640, "N": 2048, "K": 1024, "epi": "gelu", "name": "mlp-up fwd", "perf": 1}'
python tools/gemm_probe.py --case "$C2" > gpurun_out/p_gelu.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 1 -o gpurun_out/prof_gelu -f python tools/gemm_probe.py --case "$C2" > gpurun_out/n_gelu.log 2>&1
tail -n 1 gpurun_out/p_gelu.log | cut -c1-300
