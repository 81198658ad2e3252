python tools/step_time.py 60 2>&1 | tail -1
CSTP_HALO_2D=0 python tools/step_time.py 60 2>&1 | tail -1
python tools/layer_profile.py 60 > gpurun_out/layer_profile_2d.log 2>&1
grep -E "conv2|conv3.block1.conv2|TOTAL" gpurun_out/layer_profile_2d.log | grep -v wgrad
