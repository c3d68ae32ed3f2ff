#!/bin/bash
# full GPU suite, smoke and two short benches: the check run after a change
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
echo "== pytest -m gpu" > $O/r2b_check.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^E|Error" | head -n 12 >> $O/r2b_check.log
echo "== smoke" >> $O/r2b_check.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 >> $O/r2b_check.log
for i in 1 2; do
echo "== bench ($i)" >> $O/r2b_check.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-secondary --skip-cpu-baseline 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' ', d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['model_frac_of_peak_sustained'], d['roofline']['frac'], d['gpu_launches'])" >> $O/r2b_check.log 2>&1
done
cat $O/r2b_check.log
