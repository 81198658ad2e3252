# wgrad_gemm with several M tiles per CTA: parity, then same-box A/B of the per-layer weight-gradient times and the step
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -3
for mt in 1 4; do
  CSTP_WGRAD_MT=$mt timeout 300 python tools/layer_profile.py 60 2>/dev/null | grep -E "^wgrad|TOTAL" > gpurun_out/r02b_layer_wgrad_mt$mt.txt
  tail -3 gpurun_out/r02b_layer_wgrad_mt$mt.txt
  CSTP_WGRAD_MT=$mt timeout 300 python tools/step_time.py 60 --steps 20 | cut -c1-100
done
