#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
echo "== pytest model subset" > $O/r2b_early.log
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "dropout or predrawn or graph" 2>&1 | tail -n 2 >> $O/r2b_early.log
for i in 1 2; do for v in 1 0; do
  echo "== bench NEUROVIT_BITS_EARLY=$v ($i)" >> $O/r2b_early.log
  NEUROVIT_BITS_EARLY=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-secondary --skip-cpu-baseline --no-kernel-events 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' ', d['ms_per_step'], d['value'], d['e2e']['value'])" >> $O/r2b_early.log 2>&1
done; done
echo "== graph gap probe" >> $O/r2b_early.log
timeout 200 python tools/graph_gap_probe.py 2>&1 | grep -E "^[A-E]" >> $O/r2b_early.log
cat $O/r2b_early.log
