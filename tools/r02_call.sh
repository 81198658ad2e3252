set -x
mkdir -p gpurun_out
CSTP_FUSE_POLICY=all python tools/layer_profile.py 60 > gpurun_out/r02_layer_all.txt 2>&1
CSTP_FUSE_POLICY=none python tools/layer_profile.py 60 > gpurun_out/r02_layer_none.txt 2>&1
CSTP_FUSE_POLICY=all python tools/kernel_profile.py 60 --no-overlap > gpurun_out/r02_kprof_all.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_config3.py -x -q 2>&1 | tail -3
