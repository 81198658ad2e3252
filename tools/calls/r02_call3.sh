set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_finetune.py -x -q 2>&1 | tail -15 > gpurun_out/r02_stem_tests.log; cat gpurun_out/r02_stem_tests.log
timeout 300 python tools/step_time.py 60 > gpurun_out/r02_stem_step.json 2> gpurun_out/r02_stem_step.err; cat gpurun_out/r02_stem_step.json
timeout 300 python tools/layer_profile.py 60 2>/dev/null | grep -E "conv1|TOTAL" > gpurun_out/r02_stem_layers.txt; cat gpurun_out/r02_stem_layers.txt
timeout 300 python tools/kernel_profile.py 60 2>/dev/null | head -20 > gpurun_out/r02_stem_kprof.txt; cat gpurun_out/r02_stem_kprof.txt
