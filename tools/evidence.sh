#!/bin/bash
# One-GPU evidence run behind profiles/r01_*: GPU tests, the bench lines, kernel / layer profiles and the ncu launch list.
#   gpurun --timeout 1800 -- 'bash tools/evidence.sh'      (results land in gpurun_out/ev1/)
set -x
O=gpurun_out/ev1; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/bench_n1_b60.json 2> $O/bench_n1_b60.err; tail -c 300 $O/bench_n1_b60.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --batch 16 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n1_b16.json 2>/dev/null
python tools/kernel_profile.py 60 --no-overlap > $O/kernel_profile_b60.txt 2>&1
python tools/layer_profile.py 60 > $O/layer_profile_b60.txt 2>&1
python tools/clip_bench.py 60 > $O/clip_bench.json 2>/dev/null
python tools/finetune_bench.py > $O/finetune_bench.json 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/launches_b60.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
ls -la $O
