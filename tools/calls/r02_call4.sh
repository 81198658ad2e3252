set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02_gpu_suite.log; tail -40 gpurun_out/r02_gpu_suite.log
