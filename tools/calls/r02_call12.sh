mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_config3.py tests/test_gpu_finetune.py -x -q 2>&1 | tail -2
timeout 300 python tools/kernel_profile.py 60 --no-overlap 2>/dev/null | grep -E "sum of|stem_pack"
for i in 1 2; do timeout 300 python tools/step_time.py 60 --steps 20 | cut -c1-90; done
