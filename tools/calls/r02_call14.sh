# slab-mode dgrad (64-column layers): parity, then same-box A/B per layer and step
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -5
for sl in 0 1; do
  CSTP_SLABS=$sl timeout 300 python tools/layer_profile.py 60 2>/dev/null | grep -E "^conv_dgrad|TOTAL" > gpurun_out/r02b_layer_dgrad_slabs$sl.txt
  grep -E "conv2|TOTAL" gpurun_out/r02b_layer_dgrad_slabs$sl.txt
  CSTP_SLABS=$sl timeout 300 python tools/step_time.py 60 --steps 20 | cut -c1-100
done
