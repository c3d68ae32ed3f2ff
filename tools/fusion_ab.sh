# Upper bound of the two north_star fusions that are NOT built: remove the kernel a fusion would absorb
# and time the step. The step then computes wrong values — these runs only bound the possible gain.
#   skip_ln      no ln_fwd_kernel launches for the 12 block LayerNorms (LN folded into the QKV / MLP-up prologue at
#                zero cost would look like this; the bf16 LN output the wgrad GEMMs need would still have to be written)
#   skip_gather  no patch_gather_ln kernel (gather + LN fused into the patch GEMM's operand staging at zero cost)
# Usage (GPU box): bash tools/fusion_ab.sh > gpurun_out/fusion_ab.txt
cd ${GRAFT_REPO_ROOT:-.}
cat > /tmp/_ab_line.py <<'PY'
import json, sys
d = json.loads(sys.stdin.read())
print(f"{d['ms_per_step']:.3f} ms/step  {d['value']:.0f} vol/s  sm {d['clocks']['sm_mhz']} MHz {d['clocks']['reasons']}")
PY
for rep in 1 2; do
for ab in "" skip_ln skip_gather skip_ln,skip_gather; do
  line=$(NEUROVIT_AB=$ab timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --no-kernel-events --no-secondary 2>/dev/null | tail -n 1)
  echo "AB='${ab}' rep $rep: $(echo "$line" | python /tmp/_ab_line.py)"
done
done
