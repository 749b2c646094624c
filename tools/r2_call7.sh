#!/bin/bash
# round 2, call 7: out-of-bounds K chunks (DNR on the halo kernel), B2 fixes, full GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r02_pytest_c7.log; tail -25 gpurun_out/r02_pytest_c7.log
timeout 300 python tools/perf_unet.py tc 80 512 1 16 3 tc > gpurun_out/r02_perf_dnr_c7.txt 2>&1; tail -30 gpurun_out/r02_perf_dnr_c7.txt
timeout 300 python bench.py --config dnr_train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c7_dnr.json 2> gpurun_out/r02_bench_c7_dnr.err
echo "dnr: $(cut -c1-160 gpurun_out/r02_bench_c7_dnr.json)"; tail -3 gpurun_out/r02_bench_c7_dnr.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain 0 > gpurun_out/r02_bench_c7.json 2> gpurun_out/r02_bench_c7.err
echo "rnr: $(cut -c1-160 gpurun_out/r02_bench_c7.json)"; tail -3 gpurun_out/r02_bench_c7.err
