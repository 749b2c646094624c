#!/bin/bash
# Round-2 final GPU pass: parity suite, smoke, bench lines (all configs + reference arm), ncu launch list, full-set capture of the
# conv kernel (exported as CSV on the box: the reports are too large to ship).  Outputs under gpurun_out/.
TAG=${TAG:-r02_final}
mkdir -p gpurun_out /tmp/prof
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -3 gpurun_out/smoke_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; grep '^{' gpurun_out/bench_${TAG}.json | cut -c1-300
timeout 900 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; grep '^{' gpurun_out/bench_${TAG}_reference.json | cut -c1-300
timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 0 --extras --sustain 5 > gpurun_out/bench_${TAG}_extras.json 2> gpurun_out/bench_${TAG}_extras.err; grep '^{' gpurun_out/bench_${TAG}_extras.json | cut -c1-200
for c in dnr_train rnr_infer rnr_relight; do timeout 600 python bench.py --config $c --steps 100 --warmup 10 --cpu-budget 0 > gpurun_out/bench_${TAG}_$c.json 2> gpurun_out/bench_${TAG}_$c.err; grep '^{' gpurun_out/bench_${TAG}_$c.json | cut -c1-160; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --profile-steps 2 --no-graph > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv 2 > gpurun_out/launches_${TAG}_summary.txt; head -24 gpurun_out/launches_${TAG}_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -k "regex:conv_halo_kernel" -c 44 -o /tmp/prof/conv_halo python bench.py --profile-steps 1 --no-graph > /dev/null 2>&1
ncu -i /tmp/prof/conv_halo.ncu-rep --page raw --csv > gpurun_out/ncu_conv_halo_${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/ncu_conv_halo_${TAG}_raw.csv > gpurun_out/ncu_conv_halo_${TAG}_summary.txt 2>&1; head -12 gpurun_out/ncu_conv_halo_${TAG}_summary.txt
du -sh gpurun_out
