#!/bin/bash
# round 2, call 13: BatchNorm-backward sums inside the data-gradient epilogue -- parity, then A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -x -q -m gpu -s -k "batchnorm_backward_sums or forward_backward" > gpurun_out/c13_unet.log 2>&1; echo "unet rc=$?" >> gpurun_out/c13_unet.log
tail -30 gpurun_out/c13_unet.log
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu > gpurun_out/c13_fused.log 2>&1; echo "fused rc=$?" >> gpurun_out/c13_fused.log
tail -5 gpurun_out/c13_fused.log
timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c13_bench_on.json 2> gpurun_out/c13_bench_on.err
RNR_BN_BWD_FUSED=0 timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c13_bench_off.json 2> gpurun_out/c13_bench_off.err
timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c13_bench_on2.json 2>> gpurun_out/c13_bench_on.err
for f in on off on2; do grep '^{' gpurun_out/c13_bench_$f.json | cut -c1-200; done
timeout 300 python tools/perf_unet.py tc 64 512 1 108 78 tc > gpurun_out/c13_perf_unet.txt 2>&1
tail -30 gpurun_out/c13_perf_unet.txt
