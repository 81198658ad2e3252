# bn_finalize / bn_bwd_finalize / bn_partials_reduce with 32 warps per CTA: GPU suite, per-kernel times (CUPTI), step time
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02e_pytest_gpu.log
timeout 120 python tools/kernel_profile.py 60 --no-overlap 2>&1 | grep -E "sum of kernel|finalize|bn_" | tee gpurun_out/r02e_kernel_profile.txt
timeout 100 python tools/bn_probe.py 120 | tee gpurun_out/r02e_bn_probe.txt
timeout 150 python tools/step_time.py 60 --steps 20 | cut -c1-100 | tee gpurun_out/r02e_step.txt
