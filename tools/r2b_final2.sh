#!/bin/bash
# evidence of the FINAL code of round 2: default bench line, ncu launch list of the bench command, step profile, timeline
set -u
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
timeout 900 python bench.py > $O/r02d_bench_default.log 2>$O/r02d_bench_default.err
tail -n 1 $O/r02d_bench_default.log | cut -c1-200
B="python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --no-secondary"
timeout 300 $B > $O/r02d_bench_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1200 --csv --log-file $O/r02d_bench_launches.csv $B > $O/r02d_bench_ncu.log 2>&1
wc -l $O/r02d_bench_launches.csv
timeout 200 python tools/step_profile.py --out $O/r02d_step_profile_p01.md > $O/r02d_step_profile.log 2>&1; head -n 1 $O/r02d_step_profile_p01.md | cut -c1-200
timeout 200 python tools/graph_timeline.py --steps 3 --out $O/r02d_graph_timeline_p01.md --dump $O/r02d_step_kernels.csv > $O/r02d_timeline.log 2>&1; head -n 3 $O/r02d_graph_timeline_p01.md | cut -c1-400
