# parity after the slab / wgrad changes, then ncu --set full (with source) of the 144 -> 64 3x1x1 forward: prologue + statistics
# and statistics only
mkdir -p gpurun_out/ev3
timeout 500 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -4
cap() {
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c 1 -o gpurun_out/ev3/$name "$@" > gpurun_out/ev3/$name.log 2>&1 || echo "capture $name failed"
}
cap c2t_fwd_fused conv_halo 3 python tools/prologue_probe.py c2t 1 120
cap c2t_fwd_plain conv_halo 3 python tools/prologue_probe.py c2t 0 120
ls -la gpurun_out/ev3
