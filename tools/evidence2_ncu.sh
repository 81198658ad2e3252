#!/bin/bash
# `ncu --set full` of one launch of every hot kernel family at the batch-60 layer sizes, through the small per-layer probes
# (a capture inside bench.py replays against 60 GB of live allocations: minutes per kernel).
#   gpurun --timeout 900 -- 'bash tools/evidence2_ncu.sh'     (results: gpurun_out/ev2n/*.ncu-rep, each a few MB)
set -x
O=gpurun_out/ev2n; mkdir -p $O
cap() {  # name, kernel regex, launches to skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 150 ncu --set full --clock-control none -k regex:$rx --launch-skip $skip -c 1 -o $O/$name "$@" > $O/$name.log 2>&1 || echo "capture $name failed"
}
# conv2 64 -> 144 1x3x3 (plain): forward / weight gradient / dgrad; the probe launches each 7 times, the 4th is captured
cap c2s_fwd conv_halo 3 python tools/prologue_probe.py c2s 0 120
cap c2s_wgrad wgrad_halo 3 python tools/prologue_probe.py c2s 0 120
cap c2s_dgrad conv_halo 10 python tools/prologue_probe.py c2s 0 120
# conv2 144 -> 64 3x1x1 with the operand prologue: forward / weight gradient (flipped form)
cap c2t_fwd_fused conv_halo 3 python tools/prologue_probe.py c2t 1 120
cap c2t_wgrad_fused wgrad_halo 3 python tools/prologue_probe.py c2t 1 120
# conv4 256 -> 576 1x3x3 on 14 x 14 planes: conv_gemm forward / dgrad, wgrad_gemm
cap c4s_fwd conv_gemm 3 python tools/prologue_probe.py c4s 0 120
cap c4s_wgrad wgrad_gemm 3 python tools/prologue_probe.py c4s 0 120
cap c4s_dgrad conv_gemm 10 python tools/prologue_probe.py c4s 0 120
# BatchNorm streaming kernels on the conv2 144-channel tensor
cap bn_reduce_fwd 'bn_reduce_kernel<\(bool\)0' 1 python tools/bn_probe.py 120
cap bn_apply bn_apply 1 python tools/bn_probe.py 120
cap bn_reduce_bwd 'bn_reduce_kernel<\(bool\)1' 1 python tools/bn_probe.py 120
cap bn_bwd_apply bn_bwd_apply 1 python tools/bn_probe.py 120
du -sh $O; ls -la $O | head -40
