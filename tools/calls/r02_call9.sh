mkdir -p gpurun_out
for c in c2t; do
for d in 0 4; do echo "dbg=$d"; CSTP_WG_DBG=$d CSTP_HC_DBG=$d python tools/prologue_probe.py $c 1 60 2>&1 | grep -v dgrad; done
python tools/prologue_probe.py $c 0 60 2>&1 | grep -v dgrad
done
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_config3.py -x -q 2>&1 | tail -3
for i in 1 2; do timeout 300 python tools/step_time.py 60 --steps 20 2>/dev/null | tail -1 | cut -c1-90; done
