#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_fused_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu > gpurun_out/c28_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c28_tests.log; tail -3 gpurun_out/c28_tests.log
for m in 0 1 2; do RNR_CONV_DIRECT=$m timeout 200 python tools/perf_unet.py tc 64 512 1 108 78 tc > gpurun_out/c28_perf_direct$m.txt 2>&1; done
paste <(cut -c1-47 gpurun_out/c28_perf_direct0.txt) <(cut -c33-47 gpurun_out/c28_perf_direct1.txt) <(cut -c33-47 gpurun_out/c28_perf_direct2.txt) | tail -25
for i in 1 2; do for m in 1 0; do RNR_CONV_DIRECT=$m timeout 300 python bench.py --steps 300 --warmup 20 --cpu-budget 0 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"DIRECT=$m\", round(d[\"value\"],1), round(d[\"e2e\"][\"value\"],1), d[\"roofline\"][\"frac\"])"; done; done
