# (1) which gradients differ between the fuse policies (bit-identity of the prologue path), (2) the whole GPU suite without -x,
# (3) BatchNorm streaming kernels, rows per trip (2 vs 4): the conv2-sized probe, then the same-box step time
mkdir -p gpurun_out
timeout 200 python tools/fuse_bits_diag.py 2>&1 | tail -40 | tee gpurun_out/r02d_fuse_bits.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest_gpu.log 2>&1; tail -8 gpurun_out/r02d_pytest_gpu.log
for b in 2 4; do for a in 2 4; do
  echo "== CSTP_BN_BWD_ROWS=$b CSTP_BN_APPLY_ROWS=$a"
  CSTP_BN_BWD_ROWS=$b CSTP_BN_APPLY_ROWS=$a timeout 100 python tools/bn_probe.py 120
done; done 2>&1 | tee gpurun_out/r02d_bn_rows_probe.txt
for cfg in "2 2" "4 2" "2 4" "4 4"; do set -- $cfg
  echo "== step: CSTP_BN_BWD_ROWS=$1 CSTP_BN_APPLY_ROWS=$2"
  CSTP_BN_BWD_ROWS=$1 CSTP_BN_APPLY_ROWS=$2 timeout 150 python tools/step_time.py 60 --steps 20 | cut -c1-100
done 2>&1 | tee gpurun_out/r02d_bn_rows_step.txt
