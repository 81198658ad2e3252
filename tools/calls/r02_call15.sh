# slab mode for the forward 64-column temporal layers (prologue + fused statistics): parity, then same-box A/B
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -5
timeout 200 python - <<'P'
from tools.gpu_kernel_check import run_case
for n in ("pro_conv2_temporal_slab", "pro_stem_temporal_slab"):
    r = run_case(n)
    print(n, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})
P
for sl in 0 1; do
  CSTP_SLABS_FWD=$sl timeout 300 python tools/layer_profile.py 60 2>/dev/null | grep -E "^conv_fwd|TOTAL" > gpurun_out/r02b_layer_fwd_slabs$sl.txt
  grep -E "conv1|conv2|TOTAL" gpurun_out/r02b_layer_fwd_slabs$sl.txt
  CSTP_SLABS_FWD=$sl timeout 300 python tools/step_time.py 60 --steps 20 | cut -c1-100
done
