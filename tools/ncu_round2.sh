# Round-2 evidence: (1) ncu launch list of the default bench command (after it exited 0 without ncu), (2) --set full
# captures of the dominant GEMM instance (roofline.traffic), the three attention kernels at the bench's dropout, the
# LayerNorm kernels and the 4D de-interleave kernel. Usage (GPU box): bash tools/ncu_round2.sh -> gpurun_out/r02_*
set -u
cd ${GRAFT_REPO_ROOT:-.}
B="python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --no-secondary"
timeout 300 $B > gpurun_out/r02_bench_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1200 --csv --log-file gpurun_out/r02_bench_launches.csv $B > gpurun_out/r02_bench_ncu.log 2>&1
tail -n 1 gpurun_out/r02_bench_plain.log | cut -c1-160; wc -l gpurun_out/r02_bench_launches.csv
NCU="ncu --set full --clock-control none --import-source on -f"
C1='{"cg": 2, "a_mn": 0, "b_mn": 0, "block_n": 256, "M": 24640, "N": 1536, "K": 1024, "epi": "plain_bf16", "name": "qkv fwd", "perf": 1}'
python tools/gemm_probe.py --case "$C1" > gpurun_out/r02_gemm_qkv.log 2>&1 && \
timeout 300 $NCU -k regex:gemm_tc -s 4 -c 1 -o gpurun_out/r02_gemm_qkv python tools/gemm_probe.py --case "$C1" > gpurun_out/r02_gemm_qkv.ncu.log 2>&1
tail -n 1 gpurun_out/r02_gemm_qkv.log | cut -c1-200
python tools/attn_probe.py --time-only --dropout 0.1 64 385 8 > gpurun_out/r02_attn.log 2>&1 && \
timeout 400 $NCU -k regex:attn_tc -s 50 -c 4 -o gpurun_out/r02_attn python tools/attn_probe.py --time-only --dropout 0.1 64 385 8 > gpurun_out/r02_attn.ncu.log 2>&1
tail -n 3 gpurun_out/r02_attn.log
python tools/ln_probe.py > gpurun_out/r02_ln.log 2>&1 && \
timeout 300 $NCU -k regex:ln_ -s 6 -c 2 -o gpurun_out/r02_ln python tools/ln_probe.py > gpurun_out/r02_ln.ncu.log 2>&1
tail -n 3 gpurun_out/r02_ln.log
python tools/fmri_probe.py > gpurun_out/r02_fmri.log 2>&1 && \
timeout 300 $NCU -k regex:fmri_ -s 4 -c 2 -o gpurun_out/r02_fmri python tools/fmri_probe.py > gpurun_out/r02_fmri.ncu.log 2>&1
cat gpurun_out/r02_fmri.log
ls -la gpurun_out/r02_*.ncu-rep
