#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_unet_gpu.py tests/test_fused_gpu.py -x -q -m gpu > gpurun_out/c24_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c24_tests.log; tail -3 gpurun_out/c24_tests.log
timeout 200 python tools/perf_unet.py tc 64 512 1 108 78 tc > gpurun_out/c24_perf_unet.txt 2>&1; cut -c1-62 gpurun_out/c24_perf_unet.txt | tail -26
for i in 1 2; do for m in 1 0; do RNR_BN_BWD_FUSED=$m timeout 300 python bench.py --steps 300 --warmup 20 --cpu-budget 0 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"BN_BWD_FUSED=$m\", round(d[\"value\"],1), round(d[\"e2e\"][\"value\"],1), d[\"roofline\"][\"frac\"])"; done; done
