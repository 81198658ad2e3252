# last capture of the round: ncu launch list (with DRAM bytes) of one eager step at the final kernels
O=gpurun_out/ev3j; mkdir -p $O
CSTP_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/launches_b60.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
python tools/summarize_launches.py $O/launches_b60.csv > $O/launches_summary_b60.txt; head -30 $O/launches_summary_b60.txt
