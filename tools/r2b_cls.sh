#!/bin/bash
# cls-only last layer: full GPU suite, then the step with and without it
cd $GRAFT_REPO_ROOT
O=gpurun_out
echo "== pytest -m gpu" > $O/r2b_cls.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 15 >> $O/r2b_cls.log
for v in 1 0; do
  echo "== bench NEUROVIT_CLS_LAST=$v" >> $O/r2b_cls.log
  NEUROVIT_CLS_LAST=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --skip-cpu-baseline 2>$O/r2b_cls_err$v.log | tail -n 1 > $O/r2b_cls_bench$v.json
  python -c "import sys,json; d=json.loads(open('$O/r2b_cls_bench$v.json').read()); print(' ', d['ms_per_step'], d['value'], d['e2e']['value'], d['config'].get('model_frac_of_peak_sustained'), d['roofline']['achieved'], d['roofline']['share_of_step'], d['gpu_launches'])" >> $O/r2b_cls.log 2>&1
done
echo "== attn probe (groups 2 default)" >> $O/r2b_cls.log
timeout 200 python tools/attn_probe.py --time-only --dropout 0.1 2>&1 | tail -n 4 >> $O/r2b_cls.log
cat $O/r2b_cls.log
