#!/bin/bash
# round 2, call 9: DNR 256 parity detail, new tests (metrics, precompute pass), ncu captures exported as CSV on the box (reports are too large to ship)
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_pipeline_gpu.py::test_dnr_real_widths_match_oracle" tests/test_metrics_gpu.py -m gpu -q -s 2>&1 | grep -v Warning | tail -40 > gpurun_out/r02_pytest_c9.log; tail -30 gpurun_out/r02_pytest_c9.log
timeout 900 python -m pytest tests/test_scripts_gpu.py -m gpu -q -s -k "one_pass" 2>&1 | tail -15 > gpurun_out/r02_pytest_c9b.log; tail -8 gpurun_out/r02_pytest_c9b.log
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
mkdir -p /tmp/prof
cap() {  # name, kernel regex, count
  timeout 900 $NCU -k "regex:$2" -c $3 -o /tmp/prof/$1 python bench.py --profile-steps 1 --no-graph > /dev/null 2>&1
  ncu -i /tmp/prof/$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1_raw.csv 2>/dev/null
  ls -la /tmp/prof/$1.ncu-rep | awk '{print $5, $9}'
}
cap conv_halo conv_halo_kernel 44
cap wgrad "wgrad_halo_kernel|wgrad_tc_kernel" 22
cap pixel "head_fwd_kernel|tail_fwd_kernel|tail_bwd_kernel|texmap_bwd_kernel|adam_wunpack_kernel|wprep_batch_kernel|adam_multi_kernel|sh_project_kernel|sh_reconstruct_kernel|flatten_mipmap_kernel" 14
cap bn "bn_bwd_reduce_fin_kernel|bn_act_fwd_kernel|bn_bwd_apply_kernel" 12
ls -la gpurun_out/r02_ncu_*; du -sh gpurun_out
