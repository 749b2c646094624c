#!/bin/bash
# per-kernel durations of the fused step's head / tail kernels for a shared-memory carve-out setting (ncu, one eager step)
for c in ${CARVEOUTS:-100 50 25}; do
  RNR_TAIL_CARVEOUT=$c ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"tail_" --csv \
      --log-file gpurun_out/tail_c$c.csv python bench.py --profile-steps 1 --no-graph > /dev/null 2>&1
  echo "carveout $c: $(grep tail_ gpurun_out/tail_c$c.csv | awk -F'","' '{print $5, $NF}' | tr -d '"' | tr '\n' ' ')"
done
