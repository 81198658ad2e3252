python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -m gpu -x -q 2>&1 | tail -4
python tools/step_time.py 60 2>&1 | tail -1
python tools/layer_profile.py 60 > gpurun_out/layer_profile_2f.log 2>&1
grep -E "TOTAL" gpurun_out/layer_profile_2f.log
python tools/step_time.py 16 2>&1 | tail -1
