timeout 600 python tools/gpu_kernel_check.py conv2_spatial_small conv2_temporal_small stem_temporal conv2_spatial_big conv2_temporal_big conv3_spatial_big 2>&1 | grep -o "^== [a-z0-9_]* \|fwd_rel\": [0-9.e-]*\|dgrad_rel\": [0-9.e-]*\|FAILED.*" | paste -sd' ' 
python tools/step_time.py 60 2>&1 | tail -1
python tools/layer_profile.py 60 > gpurun_out/layer_profile_2e.log 2>&1
grep -E "conv_halo|TOTAL" gpurun_out/layer_profile_2e.log | grep -v "target"
