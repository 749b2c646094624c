#!/bin/bash
mkdir -p gpurun_out
for m in 2 1; do
RNR_CONV_PAIR=$m timeout 300 python -m pytest tests/test_unet_gpu.py -x -q -m gpu -k "forward_backward or batchnorm_backward_sums" > gpurun_out/c21_pair${m}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c21_pair${m}_tests.log
echo "== tests RNR_CONV_PAIR=$m"; tail -4 gpurun_out/c21_pair${m}_tests.log
done
for m in 0 1 2; do echo "== RNR_CONV_PAIR=$m"; RNR_CONV_PAIR=$m timeout 200 python tools/perf_unet.py tc 64 512 1 108 78 tc > gpurun_out/c21_perf_pair$m.txt 2>&1; tail -2 gpurun_out/c21_perf_pair$m.txt | cut -c1-150; done
for m in 0 1 2; do RNR_CONV_PAIR=$m timeout 300 python bench.py --steps 200 --warmup 20 --cpu-budget 0 2>/dev/null | grep "^{" | cut -c1-120; done
