set -x
mkdir -p gpurun_out
for d in 0 1; do CSTP_WG_DBG=$d python tools/prologue_probe.py c2t 1 60 2>&1 | grep wgrad; done
python tools/prologue_probe.py c2t 0 60 2>&1 | grep wgrad
for d in 0 1; do CSTP_WG_DBG=$d python tools/prologue_probe.py c2s 1 60 2>&1 | grep wgrad; done
python tools/prologue_probe.py c2s 0 60 2>&1 | grep wgrad
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_config3.py -x -q 2>&1 | tail -3
timeout 300 python tools/step_time.py 60
