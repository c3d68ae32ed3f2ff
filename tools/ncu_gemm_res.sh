C1='{"cg": 2, "a_mn": 0, "b_mn": 0, "block_n": 256, "M": 24640, "N": 1024, "K": 512, "epi": "bias_res", "name": "out fwd", "perf": 1}'
python tools/gemm_probe.py --case "$C1" > gpurun_out/p_res.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 1 -o gpurun_out/prof3_outfwd -f python tools/gemm_probe.py --case "$C1" > gpurun_out/n_res.log 2>&1
tail -n 2 gpurun_out/p_res.log gpurun_out/n_res.log
