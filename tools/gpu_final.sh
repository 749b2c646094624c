#!/bin/bash
# Round-end GPU pass: parity tests, smoke, bench line, ncu launch list, per-launch DRAM traffic and one full capture of the
# dominant kernel.  Outputs under gpurun_out/ (copied into profiles/ by hand).
TAG=${TAG:-r01_v7}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_${TAG}.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --profile-steps 2 --no-graph > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv 2 > gpurun_out/launches_${TAG}_summary.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    -k regex:conv_halo_kernel --csv --log-file gpurun_out/conv_halo_traffic_${TAG}.csv python bench.py --profile-steps 1 --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_halo_kernel -s 5 -c 2 -f \
    -o gpurun_out/prof_${TAG} python bench.py --profile-steps 1 --no-graph > /dev/null 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.log; tail -3 gpurun_out/smoke_${TAG}.log; cut -c1-400 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
head -16 gpurun_out/launches_${TAG}_summary.txt
