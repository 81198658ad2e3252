# batched weight pack with eight columns per thread: GPU suite, per-kernel times (CUPTI), then the final bench line of the round
mkdir -p gpurun_out/ev3i
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/ev3i/pytest_gpu.log 2>&1; tail -3 gpurun_out/ev3i/pytest_gpu.log
timeout 120 python tools/kernel_profile.py 60 --no-overlap > gpurun_out/ev3i/kernel_profile_b60.txt 2>&1; grep -E "sum of kernel|pack|finalize" gpurun_out/ev3i/kernel_profile_b60.txt
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/ev3i/bench_n1_b60.json 2> gpurun_out/ev3i/bench_n1_b60.err; cut -c1-330 gpurun_out/ev3i/bench_n1_b60.json
