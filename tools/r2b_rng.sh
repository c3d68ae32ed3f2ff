#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
echo "== pytest -m gpu" > $O/r2b_rng.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 3 >> $O/r2b_rng.log
echo "== bits probe" >> $O/r2b_rng.log
timeout 100 python tools/bits_probe.py 2>&1 | tail -n 5 >> $O/r2b_rng.log
echo "== gemm gelu shapes" >> $O/r2b_rng.log
timeout 100 python tools/gelu_gemm_probe.py 2>&1 | tail -n 3 >> $O/r2b_rng.log
for i in 1 2; do
echo "== bench default ($i)" >> $O/r2b_rng.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --skip-cpu-baseline 2>$O/r2b_rng_err.log | tail -n 1 > $O/r2b_rng_bench.json
python -c "import sys,json; d=json.loads(open('$O/r2b_rng_bench.json').read()); print(' ', d['ms_per_step'], d['value'], d['e2e']['value'], d['config'].get('model_frac_of_peak_sustained'), d['roofline']['achieved'], d['roofline']['share_of_step'], d['gpu_launches'], d.get('clocks'))" >> $O/r2b_rng.log 2>&1
done
cat $O/r2b_rng.log
