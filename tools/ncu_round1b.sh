# ncu --set full captures of the LayerNorm-backward pipeline kernel and the attention forward (v2), each after the
# same command exited 0 without ncu. Usage (GPU box): bash tools/ncu_round1b.sh -> gpurun_out/r01_lnbwd / r01_attnfwd
python tools/ln_probe.py > gpurun_out/r01_ln2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ln_bwd_pipe -s 4 -c 2 -o gpurun_out/r01_lnbwd -f python tools/ln_probe.py > gpurun_out/r01_lnbwd.ncu.log 2>&1
tail -n 4 gpurun_out/r01_ln2.log
python tools/attn_probe.py --impl 0 --time-only > gpurun_out/r01_attn2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 4 -c 1 -o gpurun_out/r01_attnfwd -f python tools/attn_probe.py --impl 0 --time-only > gpurun_out/r01_attnfwd.ncu.log 2>&1
tail -n 2 gpurun_out/r01_attn2.log
