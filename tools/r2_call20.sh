#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/c20_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c20_pytest_gpu.log
tail -6 gpurun_out/c20_pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c20_bench_on.json 2> gpurun_out/c20_bench_on.err
RNR_PDL=0 timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 > gpurun_out/c20_bench_off.json 2> gpurun_out/c20_bench_off.err
for f in on off; do grep '^{' gpurun_out/c20_bench_$f.json | cut -c1-200; done
for c in dnr_train rnr_infer rnr_relight; do timeout 600 python bench.py --config $c --steps 100 --warmup 10 --cpu-budget 0 > gpurun_out/c20_bench_$c.json 2> gpurun_out/c20_bench_$c.err; grep '^{' gpurun_out/c20_bench_$c.json | cut -c1-160; done
