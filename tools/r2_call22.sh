#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/c22_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c22_pytest_gpu.log
tail -4 gpurun_out/c22_pytest_gpu.log
timeout 200 python tools/perf_unet.py tc 64 512 1 108 78 tc > gpurun_out/c22_perf_unet.txt 2>&1; tail -2 gpurun_out/c22_perf_unet.txt | cut -c1-150
RNR_CONV_PAIR_MINBN=64 timeout 200 python tools/perf_unet.py tc 64 512 1 108 78 tc > gpurun_out/c22_perf_unet_minbn64.txt 2>&1; tail -2 gpurun_out/c22_perf_unet_minbn64.txt | cut -c1-150
for i in 1 2; do for m in 1 0; do RNR_CONV_PAIR=$m timeout 300 python bench.py --steps 300 --warmup 20 --cpu-budget 0 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"PAIR=$m\", round(d[\"value\"],1), round(d[\"e2e\"][\"value\"],1), d[\"roofline\"][\"frac\"])"; done; done
timeout 300 python bench.py --config dnr_train --steps 100 --warmup 10 --cpu-budget 0 2>/dev/null | grep "^{" | cut -c1-130
