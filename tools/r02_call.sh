set -x
mkdir -p gpurun_out
python tools/gpu_kernel_check.py conv2_spatial_small conv2_temporal_small stem_temporal conv2_spatial_big conv2_temporal_big conv3_spatial_big pro_conv2_spatial pro_conv2_temporal pro_stem_temporal pro_conv3_spatial > gpurun_out/r02_check_issue.log 2>&1
grep -o '^== [a-z0-9_]* \|"fwd_rel": [0-9.e-]*\|"dgrad_rel": [0-9.e-]*\|"fwd_equal": [a-z]*\|"fwd_tflops": [0-9.]*\|"dgrad_ms": [0-9.]*\|FAILED.*' gpurun_out/r02_check_issue.log | paste -s -d' ' | sed 's/== /\n/g'
for c in c2s c2t; do python tools/prologue_probe.py $c 0; python tools/prologue_probe.py $c 1; done 2>&1 | grep -v "^+"
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -2
for pol in all none; do CSTP_FUSE_POLICY=$pol timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$pol', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), round(d['roofline']['all_tensor_kernels_tflops'],1), {k:round(v['ms_per_step'],2) for k,v in d['roofline']['kernels'].items()})
"; done
