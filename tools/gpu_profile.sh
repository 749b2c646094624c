#!/bin/bash
# ncu launch list of exactly 2 eager training steps (+ optionally one full capture of a kernel).  Outputs under gpurun_out/.
#   TAG=r01_v1 KERNEL=conv_tc_kernel bash tools/gpu_profile.sh
mkdir -p gpurun_out
TAG=${TAG:-cur}
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --profile-steps 2 --no-graph > gpurun_out/bench_ncu_${TAG}.log 2>&1
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv 2 > gpurun_out/launches_${TAG}_summary.txt
head -45 gpurun_out/launches_${TAG}_summary.txt
if [ -n "$KERNEL" ]; then
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${KERNEL} -s ${KSKIP:-0} -c ${KCOUNT:-3} -f -o gpurun_out/prof_${TAG} \
    python bench.py --profile-steps 1 --no-graph > gpurun_out/bench_ncu2_${TAG}.log 2>&1
fi
ls -la gpurun_out | tail -6
