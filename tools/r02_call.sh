set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_config3.py -x -q -s -k config4 2>&1 | grep -E "finetune|eval logits|passed|failed|Error" | cut -c1-300
for rep in 1 2; do for pol in none all auto; do CSTP_FUSE_POLICY=$pol timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$pol', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), round(d['roofline']['all_tensor_kernels_tflops'],1))
"; done; done | tee gpurun_out/r02_policy_ab.txt
