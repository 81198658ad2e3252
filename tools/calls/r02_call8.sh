set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_config3.py tests/test_gpu_finetune.py -x -q 2>&1 | tail -3
timeout 300 python tools/layer_profile.py 60 2>/dev/null > gpurun_out/r02_layers_all.txt; grep -E "dgrad|TOTAL" gpurun_out/r02_layers_all.txt
timeout 300 python tools/step_time.py 60 --steps 20 | cut -c1-100
CSTP_DGRAD_ONE_LAUNCH=0 timeout 300 python tools/step_time.py 60 --steps 20 | cut -c1-100
