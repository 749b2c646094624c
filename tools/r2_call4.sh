#!/bin/bash
# round 2, call 4: re-measure after fused-finalize / sh_project / BN-reduce changes; BN RU sweep; GCN timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_fused_gpu.py tests/test_pipeline_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_pytest_c4.log
grep -E "passed|failed|error|Error" gpurun_out/r02_pytest_c4.log | tail -15
for ru in 2 3 4; do
  RNR_BN_RU=$ru timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c4_ru$ru.json 2> gpurun_out/r02_bench_c4_ru$ru.err
  echo "RU=$ru: $(cut -c1-120 gpurun_out/r02_bench_c4_ru$ru.json)"; tail -2 gpurun_out/r02_bench_c4_ru$ru.err
done
RNR_BN_FUSED_FINALIZE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c4_nofin.json 2>/dev/null
echo "no fused finalize: $(cut -c1-120 gpurun_out/r02_bench_c4_nofin.json)"
RNR_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c4_pdl.json 2>/dev/null
echo "PDL: $(cut -c1-120 gpurun_out/r02_bench_c4_pdl.json)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --profile-steps 2 --no-graph > gpurun_out/r02_ncu_c4.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_c4.csv 2 > gpurun_out/r02_launches_c4_summary.txt
head -24 gpurun_out/r02_launches_c4_summary.txt
timeout 300 python tools/time_gcn.py > gpurun_out/r02_gcn_c4.txt 2>&1; tail -12 gpurun_out/r02_gcn_c4.txt
