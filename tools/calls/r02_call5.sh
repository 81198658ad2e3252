set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config3.py tests/test_gpu_kernels.py -x -q -k "graph or bit_identical or bn_sync or elementwise" 2>&1 | tail -15 > gpurun_out/r02_graph_tests.log; cat gpurun_out/r02_graph_tests.log
for B in 60 16; do
timeout 300 python tools/step_time.py $B 2>gpurun_out/err.txt | tail -1; tail -2 gpurun_out/err.txt
timeout 300 python tools/step_time.py $B --no-graph 2>/dev/null | tail -1
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err; tail -3 gpurun_out/r02_bench5.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench5.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['roofline']['frac'])"
timeout 600 python bench.py --steps 10 --warmup 3 --batch 16 > gpurun_out/r02_bench5_b16.json 2> gpurun_out/r02_bench5_b16.err; tail -3 gpurun_out/r02_bench5_b16.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench5_b16.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])"
