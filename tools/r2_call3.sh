#!/bin/bash
# round 2, call 3: fused BN finalize + fused optimiser + small-loss kernels: parity, bench, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_fused_gpu.py tests/test_pipeline_gpu.py -m gpu -q -s 2>&1 | tail -120 > gpurun_out/r02_pytest_c3.log
grep -E "passed|failed|error|Error" gpurun_out/r02_pytest_c3.log | tail -15
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
cut -c1-300 gpurun_out/r02_bench_c3.json; tail -5 gpurun_out/r02_bench_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_c3.csv \
    python bench.py --profile-steps 2 --no-graph > gpurun_out/r02_ncu_c3.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_c3.csv 2 > gpurun_out/r02_launches_c3_summary.txt
head -40 gpurun_out/r02_launches_c3_summary.txt
