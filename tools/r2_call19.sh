#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_fused_gpu.py -x -q -m gpu > gpurun_out/c19_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c19_tests.log
tail -4 gpurun_out/c19_tests.log
timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c19_bench_on.json 2> gpurun_out/c19_bench_on.err
RNR_PDL=0 timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c19_bench_off.json 2> gpurun_out/c19_bench_off.err
timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c19_bench_on2.json 2>> gpurun_out/c19_bench_on.err
for f in on off on2; do grep '^{' gpurun_out/c19_bench_$f.json | cut -c1-200; done; tail -3 gpurun_out/c19_bench_on.err
