#!/bin/bash
# compute-sanitizer passes (memcheck + racecheck):
#   * one small training iteration (module path + fused path incl. the fused optimiser) -- __graft_entry__.smoke()
#   * memcheck only (API-error reports off: cudart's lazy kernel lookup trips a benign cuKernelGetFunction report): the round-2 kernels at a size that selects them -- CTA-pair / register-direct conv variants, optimiser pass
#     writing the GEMM matrices, BatchNorm-backward apply-from-totals, forward BatchNorm from totals, probe stitching
# Run on the GPU box:  bash tools/sanitize.sh  -> gpurun_out/r02_sanitize_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitize_$tool.log 2>&1
  echo "$tool: exit $? -- $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_sanitize_$tool.log | tail -1)"
done
RNR_BN_BWD_FUSED=1 RNR_BN_FWD_TOTALS=1 timeout 1500 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 --print-limit 20 \
    python -m pytest -x -q -m gpu "tests/test_fused_gpu.py::test_optimiser_pass_writes_the_next_steps_gemm_matrices_bit_exactly" \
    tests/test_unet_gpu.py::test_conv_kernel_variants_agree tests/test_stitch.py::test_device_stitcher_matches_the_restatement \
    > gpurun_out/r02_sanitize_memcheck_round2_kernels.log 2>&1
echo "memcheck (round-2 kernels): exit $? -- $(grep -E 'ERROR SUMMARY' gpurun_out/r02_sanitize_memcheck_round2_kernels.log | tail -1) -- $(grep -E 'passed|failed' gpurun_out/r02_sanitize_memcheck_round2_kernels.log | tail -1)"
