C1='{"cg": 2, "a_mn": 0, "b_mn": 0, "block_n": 256, "M": 24640, "N": 1536, "K": 1024, "epi": "plain_bf16", "name": "qkv fwd", "perf": 1}'
C2='{"cg": 2, "a_mn": 0, "b_mn": 0, "block_n": 256, "M": 24640, "N": 2048, "K": 1024, "epi": "gelu", "name": "mlp-up fwd", "perf": 1}'
python tools/gemm_probe.py --case "$C1" > gpurun_out/p1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 1 -o gpurun_out/prof2_qkv_cg2 -f python tools/gemm_probe.py --case "$C1" > gpurun_out/n1.log 2>&1
python tools/gemm_probe.py --case "$C2" > gpurun_out/p2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 1 -o gpurun_out/prof2_mlpup_cg2 -f python tools/gemm_probe.py --case "$C2" > gpurun_out/n2.log 2>&1
tail -n 2 gpurun_out/n1.log gpurun_out/n2.log
