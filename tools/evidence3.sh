#!/bin/bash
# One-GPU evidence run behind profiles/r02f_* (final kernels of round 2): GPU tests, bench lines, kernel / layer profiles and
# the ncu launch list (with DRAM bytes) of one eager step at batch 60.
#   gpurun --timeout 1000 -- 'bash tools/evidence3.sh'   (results: gpurun_out/ev3f/)
set -x
O=gpurun_out/ev3f; mkdir -p $O
timeout 420 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 200 python bench.py --steps 20 --warmup 5 > $O/bench_n1_b60.json 2> $O/bench_n1_b60.err; tail -c 400 $O/bench_n1_b60.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 120 python bench.py --batch 16 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1_b16.json 2>/dev/null
timeout 120 python tools/layer_profile.py 60 > $O/layer_profile_b60.txt 2>&1; tail -4 $O/layer_profile_b60.txt
timeout 120 python tools/kernel_profile.py 60 --no-overlap > $O/kernel_profile_b60.txt 2>&1
CSTP_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/launches_b60.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
python tools/summarize_launches.py $O/launches_b60.csv > $O/launches_summary_b60.txt; head -14 $O/launches_summary_b60.txt
ls -la $O
