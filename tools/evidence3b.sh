#!/bin/bash
# Final one-GPU evidence of round 2 (after the bit-identity fix and bn_bwd_apply at four rows per trip): GPU tests, smoke(),
# bench lines, the ncu launch list of one eager step and `ncu --set full` of the kernels that changed in the second half of
# the round (slab-mode dgrad, CTA-pair conv_gemm, multi-M-tile wgrad_gemm, bn_bwd_apply).
#   gpurun --timeout 900 -- 'bash tools/evidence3b.sh'   (results: gpurun_out/ev3g/)
set -x
O=gpurun_out/ev3g; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
timeout 200 python bench.py --steps 20 --warmup 5 > $O/bench_n1_b60.json 2> $O/bench_n1_b60.err; cut -c1-330 $O/bench_n1_b60.json
timeout 120 python bench.py --batch 16 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1_b16.json 2>/dev/null
timeout 120 python tools/kernel_profile.py 60 --no-overlap > $O/kernel_profile_b60.txt 2>&1
CSTP_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/launches_b60.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
python tools/summarize_launches.py $O/launches_b60.csv > $O/launches_summary_b60.txt; head -12 $O/launches_summary_b60.txt
cap() {  # name, kernel regex, launches to skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 120 ncu --set full --clock-control none -k regex:$rx --launch-skip $skip -c 1 -o $O/$name "$@" > $O/$name.log 2>&1 || echo "capture $name failed"
}
cap c2s_dgrad conv_halo 10 python tools/prologue_probe.py c2s 0 120
cap c4s_fwd conv_gemm 3 python tools/prologue_probe.py c4s 0 120
cap c4s_wgrad wgrad_gemm 3 python tools/prologue_probe.py c4s 0 120
cap c4s_dgrad conv_gemm 10 python tools/prologue_probe.py c4s 0 120
cap bn_bwd_apply bn_bwd_apply 1 python tools/bn_probe.py 120
for f in $O/*.ncu-rep; do ncu -i $f --page raw --csv 2>/dev/null | python tools/summarize_ncu_full.py > ${f%.ncu-rep}.txt 2>&1; done
rm -f $O/launches_b60.csv.tmp; du -sh $O; ls -la $O | head -40
