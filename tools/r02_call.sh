set -x
mkdir -p gpurun_out
nvidia-smi -L
python tools/gpu_kernel_check.py pro_conv2_spatial pro_conv2_temporal pro_stem_temporal pro_conv3_temporal_s2 pro_conv3_spatial pro_conv3_temporal pro_conv4_spatial pro_conv5_spatial pro_conv5_temporal pro_ds_temporal pro_ragged > gpurun_out/r02_pro_check.log 2>&1
cat gpurun_out/r02_pro_check.log | cut -c1-600
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest1.log 2>&1; tail -15 gpurun_out/r02_gputest1.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; tail -3 gpurun_out/r02_bench1.err; cat gpurun_out/r02_bench1.json | cut -c1-1500
CSTP_FUSE_BN_APPLY=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench1_nofuse.json 2> gpurun_out/r02_bench1_nofuse.err; cat gpurun_out/r02_bench1_nofuse.json | cut -c1-600
