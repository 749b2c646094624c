#!/bin/bash
# round 2, call 5: new bench.py (all configs), defaults reverted (RU=2, separate finalize), sh_project unrolled
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py tests/test_pixel_ops_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_pytest_c5.log; tail -3 gpurun_out/r02_pytest_c5.log
timeout 600 python bench.py --steps 20 --warmup 5 --extras > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err
echo "train: $(cut -c1-160 gpurun_out/r02_bench_c5.json)"; tail -3 gpurun_out/r02_bench_c5.err
for c in rnr_infer rnr_relight dnr_train; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c5_$c.json 2> gpurun_out/r02_bench_c5_$c.err
  echo "$c: $(cut -c1-200 gpurun_out/r02_bench_c5_$c.json)"; tail -3 gpurun_out/r02_bench_c5_$c.err
done
python - <<'PY'
import json
for f in ('r02_bench_c5', 'r02_bench_c5_rnr_infer', 'r02_bench_c5_rnr_relight', 'r02_bench_c5_dnr_train'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), {k: d.get(k) for k in ('sustained', 'extras', 'rasterizer', 'images_per_s') if d.get(k)})
        if d.get('roofline'): print('   roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'kernel_ms_per_step', 'whole_step_tflops')}, 'cpu', d.get('cpu_baseline'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
