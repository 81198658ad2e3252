python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -m gpu -q 2>&1 | tail -3
python tools/step_time.py 60 | tail -1
python tools/layer_profile.py 60 > gpurun_out/layer_profile_2g.log 2>&1
grep -E "^wgrad|TOTAL" gpurun_out/layer_profile_2g.log | grep -E "conv2|conv3.block1.conv2|conv1|TOTAL"
