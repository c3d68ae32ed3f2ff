#!/bin/bash
# last check of the round: full GPU suite, smoke, default bench (all secondaries)
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
echo "== pytest -m gpu" > $O/r02c_final.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -n 2 >> $O/r02c_final.log
echo "== smoke" >> $O/r02c_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 >> $O/r02c_final.log
echo "== default bench" >> $O/r02c_final.log
timeout 900 python bench.py > $O/r02c_bench_default.log 2>$O/r02c_bench_default.err
tail -n 1 $O/r02c_bench_default.log | cut -c1-300 >> $O/r02c_final.log
echo "== reference arm (short)" >> $O/r02c_final.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -n 1 | cut -c1-300 >> $O/r02c_final.log
cat $O/r02c_final.log
