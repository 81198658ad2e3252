mkdir -p gpurun_out
python tools/prologue_probe.py c2t 1 60 2>&1 | grep -v dgrad
ncu --set full --clock-control none --import-source on -k regex:conv_halo -c 1 -o gpurun_out/r02_hcx_c2t python tools/prologue_probe.py c2t 1 16 > gpurun_out/ncu11.log 2>&1
tail -3 gpurun_out/ncu11.log
