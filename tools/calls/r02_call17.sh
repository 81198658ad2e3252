# bn_apply with two rows per trip; slab-layout fix: parity (kernels + step), BatchNorm probe, step time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -x -q 2>&1 | tail -4
timeout 120 python tools/bn_probe.py 120
timeout 300 python tools/step_time.py 60 --steps 20 | cut -c1-100
