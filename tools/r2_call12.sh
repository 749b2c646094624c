#!/bin/bash
# round 2, call 12: optimiser pass writing the GEMM matrices -- parity, then A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py -x -q -m gpu -s > gpurun_out/c12_fused.log 2>&1; echo "fused rc=$?" >> gpurun_out/c12_fused.log
timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c12_bench_on.json 2> gpurun_out/c12_bench_on.err
RNR_ADAM_WRITES_GEMM=0 timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c12_bench_off.json 2> gpurun_out/c12_bench_off.err
timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c12_bench_on2.json 2>> gpurun_out/c12_bench_on.err
timeout 600 python bench.py --config dnr_train --steps 100 --warmup 10 --cpu-budget 0 > gpurun_out/c12_bench_dnr.json 2> gpurun_out/c12_bench_dnr.err
tail -5 gpurun_out/c12_fused.log
cat gpurun_out/c12_bench_on.json gpurun_out/c12_bench_off.json gpurun_out/c12_bench_on2.json gpurun_out/c12_bench_dnr.json | cut -c1-400
