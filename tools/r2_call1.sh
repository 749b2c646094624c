#!/bin/bash
# round 2, call 1: full GPU parity suite (new tests included), smoke, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r02_pytest_gpu_c1.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_c1.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c1.json 2> gpurun_out/r02_bench_c1.err
grep -E "passed|failed|error" gpurun_out/r02_pytest_gpu_c1.log | tail -5; tail -2 gpurun_out/r02_smoke_c1.log; cut -c1-600 gpurun_out/r02_bench_c1.json; tail -3 gpurun_out/r02_bench_c1.err
