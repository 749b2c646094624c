#!/bin/bash
# round 2, call 8: rest of the GPU suite + ncu full-set captures of the top kernels (conv_halo 512^2 / 32^2, wgrad_halo, wgrad_tc, BN reduce, head / tail)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_pytest_c8.log; tail -12 gpurun_out/r02_pytest_c8.log
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
timeout 900 $NCU -k regex:conv_halo_kernel -c 44 -o gpurun_out/r02_prof_conv_halo python bench.py --profile-steps 1 --no-graph > /dev/null 2>&1
timeout 600 $NCU -k regex:wgrad_halo_kernel -c 6 -o gpurun_out/r02_prof_wgrad_halo python bench.py --profile-steps 1 --no-graph > /dev/null 2>&1
timeout 600 $NCU -k regex:wgrad_tc_kernel -c 16 -o gpurun_out/r02_prof_wgrad_tc python bench.py --profile-steps 1 --no-graph > /dev/null 2>&1
timeout 600 $NCU -k "regex:head_fwd_kernel|tail_fwd_kernel|tail_bwd_kernel|texmap_bwd_kernel|adam_wunpack_kernel|wprep_batch_kernel|adam_multi_kernel|sh_project_kernel" -c 12 -o gpurun_out/r02_prof_pixel python bench.py --profile-steps 1 --no-graph > /dev/null 2>&1
timeout 600 $NCU -k "regex:bn_bwd_reduce_fin_kernel|bn_act_fwd_kernel|bn_bwd_apply_kernel" -c 12 -o gpurun_out/r02_prof_bn python bench.py --profile-steps 1 --no-graph > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
