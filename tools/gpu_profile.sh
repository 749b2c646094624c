#!/bin/bash
# ncu launch list of a short bench run + one full capture of the conv kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1500} -c ${COUNT:-900} --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-conv_tc_kernel} -s ${KSKIP:-100} -c 3 -f -o gpurun_out/prof_conv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu2.log 2>&1
ls -la gpurun_out | tail -8
