# ncu --set full captures of individual kernels INSIDE the benchmark step (real operands, real neighbours), one
# invocation per kernel variant. Usage (GPU box): bash tools/ncu_step_kernels.sh -> gpurun_out/r02_step_<tag>.ncu-rep
cd ${GRAFT_REPO_ROOT:-.}
B="python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --no-secondary --no-kernel-events"
cap() {  # tag, demangled-name regex, launches to skip
  timeout 400 ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 \
    -o gpurun_out/r02_step_$1 $B > gpurun_out/r02_step_$1.log 2>&1
  ls -la gpurun_out/r02_step_$1.ncu-rep 2>&1 | cut -c1-120
}
cap gelu_dgrad 'gemm_tc_kernel<.int.256, .int.5, .bool.0, .bool.1, .int.2, .int.2, .bool.1>' 8
cap gelu_fwd 'gemm_tc_kernel<.int.256, .int.5, .bool.0, .bool.0, .int.2, .int.1, .bool.1>' 8
cap down_fwd 'gemm_tc_kernel<.int.256, .int.6, .bool.0, .bool.0, .int.2, .int.0, .bool.1>' 8
cap out_fwd 'gemm_tc_kernel<.int.256, .int.4, .bool.0, .bool.0, .int.2, .int.3, .bool.1>' 8
cap wgrad 'gemm_tc_kernel<.int.256, .int.6, .bool.1, .bool.1, .int.2, .int.0, .bool.0>' 30
