#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -x -q -m gpu -s -k "batchnorm_backward_sums or forward_backward" > gpurun_out/c15_unet.log 2>&1; echo "unet rc=$?" >> gpurun_out/c15_unet.log
grep -E "rel-L2|passed|failed|rc=" gpurun_out/c15_unet.log | tail -30
timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c15_bench_on.json 2> gpurun_out/c15_bench_on.err
RNR_BN_BWD_FUSED=0 timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c15_bench_off.json 2> gpurun_out/c15_bench_off.err
for f in on off; do grep '^{' gpurun_out/c15_bench_$f.json | cut -c1-200; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c15_launches.csv python bench.py --profile-steps 2 --no-graph > gpurun_out/c15_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/c15_launches.csv 2 > gpurun_out/c15_launches_summary.txt; head -16 gpurun_out/c15_launches_summary.txt
