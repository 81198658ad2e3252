set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_config3.py -x -q 2>&1 | tail -15 > gpurun_out/r02_wg_tests.log; cat gpurun_out/r02_wg_tests.log
timeout 300 python tools/step_time.py 60 > gpurun_out/r02_wg_step.json 2> gpurun_out/r02_wg_step.err; cat gpurun_out/r02_wg_step.json
timeout 300 python tools/layer_profile.py 60 2>/dev/null | grep -E "wgrad|TOTAL" > gpurun_out/r02_wg_layers.txt; cat gpurun_out/r02_wg_layers.txt
CSTP_FUSE_POLICY=none timeout 300 python tools/layer_profile.py 60 2>/dev/null | grep -E "wgrad|TOTAL" > gpurun_out/r02_wg_layers_none.txt; cat gpurun_out/r02_wg_layers_none.txt
