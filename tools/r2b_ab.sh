#!/bin/bash
# Round-2b A/B: attention schedules (dQ S-double-buffer, dK/dV scores-first, SIMT pair groups), mask generator, draw order.
cd $GRAFT_REPO_ROOT
O=gpurun_out
L=$PWD/neurovit_b200
echo "== pytest attention / dropout" > $O/r2b_ab.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention or keep_bits or dropout" 2>&1 | tail -n 4 >> $O/r2b_ab.log
echo "== probe default (dq2, dkv scores-first, groups 3): check + time" >> $O/r2b_ab.log
timeout 200 python tools/attn_probe.py --dropout 0.1 2>&1 | grep -v "OK$" | tail -n 8 >> $O/r2b_ab.log
for cfg in "NV_ATTN_DQ=1 NV_ATTN_DKV=1" "NV_ATTN_DQ=1" "NV_ATTN_DKV=1"; do
  echo "== time-only $cfg" >> $O/r2b_ab.log
  env $cfg timeout 120 python tools/attn_probe.py --time-only --dropout 0.1 2>&1 | tail -n 4 >> $O/r2b_ab.log
done
for v in g1 g2; do
  echo "== lib $v: check + time" >> $O/r2b_ab.log
  NEUROVIT_LIB=$L/libneurovit_b200_$v.so timeout 200 python tools/attn_probe.py --dropout 0.1 2>&1 | grep -v "OK$" | tail -n 5 >> $O/r2b_ab.log
done
echo "== per-kernel (NV_ATTN_ONLY) default lib, p=0.1" >> $O/r2b_ab.log
for only in dq dkv; do
  NV_ATTN_ONLY=$only timeout 120 python tools/attn_probe.py --time-only --dropout 0.1 2>&1 | grep "p=0.1" | sed "s/^/  only=$only /" >> $O/r2b_ab.log
  NV_ATTN_ONLY=$only NV_ATTN_DQ=1 NV_ATTN_DKV=1 timeout 120 python tools/attn_probe.py --time-only --dropout 0.1 2>&1 | grep "p=0.1" | sed "s/^/  only=$only old /" >> $O/r2b_ab.log
done
echo "== bits probe" >> $O/r2b_ab.log
timeout 100 python tools/bits_probe.py 2>&1 | tail -n 5 >> $O/r2b_ab.log
echo "== bench default" >> $O/r2b_ab.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --skip-cpu-baseline 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' ', d['ms_per_step'], d['value'], d.get('roofline'))" >> $O/r2b_ab.log 2>&1
echo "== bench NEUROVIT_BITS_ORDER=after" >> $O/r2b_ab.log
NEUROVIT_BITS_ORDER=after timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --skip-cpu-baseline 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' ', d['ms_per_step'], d['value'])" >> $O/r2b_ab.log 2>&1
echo "== bench old attention schedules" >> $O/r2b_ab.log
NV_ATTN_DQ=1 NV_ATTN_DKV=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --skip-cpu-baseline 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' ', d['ms_per_step'], d['value'])" >> $O/r2b_ab.log 2>&1
cat $O/r2b_ab.log
