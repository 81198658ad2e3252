# conv_gemm CTA pairs (clusters of two, multicast weight tiles): parity of the pair cases (each in its own subprocess with a
# timeout), then same-box A/B of the per-layer times and the step time with the pairs off / on
mkdir -p gpurun_out
timeout 500 python tools/gpu_kernel_check.py conv4_spatial_pairs conv5_temporal_pairs conv4_spatial_s2_pairs conv3_temporal_s2_pairs conv5_spatial conv4_spatial_s2 2>&1 | cut -c1-900 | tee gpurun_out/r02c_pairs_cases.txt
CSTP_CONV_CLUSTER=0 timeout 200 python tools/layer_profile.py 60 > gpurun_out/r02c_layer_pairs0.txt 2>&1
CSTP_CONV_CLUSTER=1 timeout 200 python tools/layer_profile.py 60 > gpurun_out/r02c_layer_pairs1.txt 2>&1
tail -4 gpurun_out/r02c_layer_pairs0.txt; tail -4 gpurun_out/r02c_layer_pairs1.txt
CSTP_CONV_CLUSTER=0 timeout 200 python tools/step_time.py 60 --steps 20 | cut -c1-100
CSTP_CONV_CLUSTER=1 timeout 200 python tools/step_time.py 60 --steps 20 | cut -c1-100
