for C in \
 '{"cg": 2, "a_mn": 0, "b_mn": 0, "block_n": 128, "M": 24640, "N": 1024, "K": 512, "epi": "bias_res", "name": "out fwd bn128", "perf": 1}' \
 '{"cg": 1, "a_mn": 0, "b_mn": 0, "block_n": 128, "M": 24640, "N": 1024, "K": 512, "epi": "bias_res", "name": "out fwd bn128 cg1", "perf": 1}' \
 '{"cg": 2, "a_mn": 0, "b_mn": 0, "block_n": 128, "M": 24640, "N": 1024, "K": 2048, "epi": "bias_res", "name": "down fwd bn128", "perf": 1}' \
 '{"cg": 2, "a_mn": 0, "b_mn": 1, "block_n": 128, "M": 24640, "N": 512, "K": 1024, "epi": "plain_bf16", "name": "out dgrad bn128", "perf": 1}' \
 '{"cg": 2, "a_mn": 0, "b_mn": 1, "block_n": 256, "M": 24640, "N": 512, "K": 1024, "epi": "plain_bf16", "name": "out dgrad bn256", "perf": 1}' \
 '{"cg": 2, "a_mn": 1, "b_mn": 1, "block_n": 128, "M": 1024, "N": 512, "K": 24640, "epi": "splitk", "name": "out wgrad bn128", "perf": 1}' ; do
  python tools/gemm_probe.py --case "$C" 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
r = json.loads(l[7:]) if l.startswith('RESULT ') else None
print(r['case']['name'], 'splits', r.get('k_splits'), '%.1f us  %.0f TFLOP/s' % (r['ms_med'] * 1e3, r['tflops'])) if r else print(l[-300:])"
done
