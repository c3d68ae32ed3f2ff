# ncu --set full captures of the dominant kernels, each after the same command exited 0 without ncu.
# Usage (on the GPU box): bash tools/ncu_round1.sh   -> gpurun_out/r01_*.ncu-rep
set -u
C1='{"cg": 2, "a_mn": 0, "b_mn": 0, "block_n": 256, "M": 24640, "N": 1536, "K": 1024, "epi": "plain_bf16", "name": "qkv fwd", "perf": 1}'
C2='{"cg": 2, "a_mn": 0, "b_mn": 0, "block_n": 256, "M": 24640, "N": 2048, "K": 1024, "epi": "gelu", "name": "mlp-up fwd", "perf": 1}'
C3='{"cg": 2, "a_mn": 0, "b_mn": 1, "block_n": 256, "M": 24640, "N": 2048, "K": 1024, "epi": "gelu_grad", "name": "mlp-down dgrad", "perf": 1}'
C4='{"cg": 2, "a_mn": 1, "b_mn": 1, "block_n": 256, "M": 2048, "N": 1024, "K": 24640, "epi": "splitk", "name": "mlp-up wgrad", "perf": 1}'
i=1
for C in "$C1" "$C2" "$C3" "$C4"; do
  python tools/gemm_probe.py --case "$C" > gpurun_out/r01_gemm$i.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 1 -o gpurun_out/r01_gemm$i -f python tools/gemm_probe.py --case "$C" > gpurun_out/r01_gemm$i.ncu.log 2>&1
  tail -n 1 gpurun_out/r01_gemm$i.log | cut -c1-220
  i=$((i+1))
done
python tools/attn_probe.py --impl 0 --time-only > gpurun_out/r01_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 16 -c 4 -o gpurun_out/r01_attn -f python tools/attn_probe.py --impl 0 --time-only > gpurun_out/r01_attn.ncu.log 2>&1
tail -n 2 gpurun_out/r01_attn.log
python tools/ln_probe.py > gpurun_out/r01_ln.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ln_ -s 6 -c 3 -o gpurun_out/r01_ln -f python tools/ln_probe.py > gpurun_out/r01_ln.ncu.log 2>&1
cat gpurun_out/r01_ln.log
