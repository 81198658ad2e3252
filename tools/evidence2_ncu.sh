#!/bin/bash
# `ncu --set full` of the hot kernel families inside one eager batch-16 step (one launch window per family, bounded):
#   gpurun --timeout 900 -- 'bash tools/evidence2_ncu.sh'     (results: gpurun_out/ev2n/*.ncu-rep, read here with ncu -i)
set -x
O=gpurun_out/ev2n; mkdir -p $O
B=${1:-16}
# kernels of the families per step: ~290; three warm-up steps are skipped, then 40 launches of the backward-heavy tail and
# 40 of the forward head are profiled
CSTP_GRAPH=0 timeout 400 ncu --set full --clock-control none -k regex:'conv_halo|conv_gemm|wgrad_halo|wgrad_gemm|bn_reduce|bn_bwd_apply|bn_apply' --launch-skip 880 -c 60 -o $O/ncu_full_fwd_b$B python bench.py --batch $B --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_fwd.log 2>&1
CSTP_GRAPH=0 timeout 400 ncu --set full --clock-control none -k regex:'conv_halo|conv_gemm|wgrad_halo|wgrad_gemm|bn_reduce|bn_bwd_apply|bn_apply' --launch-skip 1090 -c 70 -o $O/ncu_full_bwd_b$B python bench.py --batch $B --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_bwd.log 2>&1
ls -la $O
