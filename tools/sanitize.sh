#!/bin/bash
# compute-sanitizer pass over one small training iteration (module path + fused path, incl. the fused optimiser): memcheck and racecheck.
# Run on the GPU box:  bash tools/sanitize.sh  -> gpurun_out/r02_sanitize_{memcheck,racecheck}.log
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitize_$tool.log 2>&1
  echo "$tool: exit $? -- $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_sanitize_$tool.log | tail -1)"
done
