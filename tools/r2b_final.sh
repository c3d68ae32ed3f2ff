#!/bin/bash
# Round-2 (second half) evidence: full GPU suite, smoke, the default bench line, the ncu launch list of the bench command,
# the step profile / graph timeline, and --set full captures of the attention kernels (new dQ schedule) and the mask
# generator. Usage (GPU box): bash tools/r2b_final.sh -> gpurun_out/r02b_*
set -u
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
echo "== pytest -m gpu" > $O/r02b_final.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -n 3 >> $O/r02b_final.log
echo "== smoke" >> $O/r02b_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3 >> $O/r02b_final.log
echo "== default bench" >> $O/r02b_final.log
( nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/r02b_clocks.csv & echo $! > /tmp/smi.pid )
timeout 900 python bench.py > $O/r02b_bench_default.log 2>$O/r02b_bench_default.err
kill $(cat /tmp/smi.pid) 2>/dev/null
tail -n 1 $O/r02b_bench_default.log | cut -c1-400 >> $O/r02b_final.log
echo "== ncu launch list" >> $O/r02b_final.log
B="python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --no-secondary"
timeout 300 $B > $O/r02b_bench_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1200 --csv --log-file $O/r02b_bench_launches.csv $B > $O/r02b_bench_ncu.log 2>&1
wc -l $O/r02b_bench_launches.csv >> $O/r02b_final.log
echo "== step profile" >> $O/r02b_final.log
timeout 200 python tools/step_profile.py --out $O/r02b_step_profile_p01.md > $O/r02b_step_profile.log 2>&1; head -n 1 $O/r02b_step_profile_p01.md | cut -c1-250 >> $O/r02b_final.log
NCU="ncu --set full --clock-control none --import-source on -f"
echo "== ncu attention" >> $O/r02b_final.log
python tools/attn_probe.py --time-only --dropout 0.1 64 385 8 > $O/r02b_attn.log 2>&1 && \
timeout 400 $NCU -k regex:attn_tc -s 50 -c 3 -o $O/r02b_attn python tools/attn_probe.py --time-only --dropout 0.1 64 385 8 > $O/r02b_attn.ncu.log 2>&1
tail -n 2 $O/r02b_attn.log >> $O/r02b_final.log
echo "== ncu mask generator" >> $O/r02b_final.log
timeout 300 $NCU -k regex:dropout_bits -s 3 -c 1 -o $O/r02b_bits python tools/bits_probe.py > $O/r02b_bits.ncu.log 2>&1
ls -la $O/r02b_*.ncu-rep >> $O/r02b_final.log 2>&1
cat $O/r02b_final.log
