#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_fused_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu > gpurun_out/c30_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c30_tests.log; tail -3 gpurun_out/c30_tests.log
for i in 1 2; do for m in 1 0; do RNR_BN_FWD_TOTALS=$m timeout 300 python bench.py --steps 300 --warmup 20 --cpu-budget 0 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"BN_FWD_TOTALS=$m\", round(d[\"value\"],1), round(d[\"e2e\"][\"value\"],1), d[\"gpu_launches\"]//d[\"steps\"])"; done; done
