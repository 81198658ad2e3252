set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ntxent.py -x -q 2>&1 | tail -3
timeout 300 python tools/ntxent_sweep.py > gpurun_out/r02_ntxent_sweep.jsonl 2> gpurun_out/r02_ntxent_sweep.err; python - <<'PY'
import json
for l in open('gpurun_out/r02_ntxent_sweep.jsonl'):
    d=json.loads(l); print(d['rows'], round(d['kernel_ms_fwd_bwd'],4), round(d['module_ms_fwd_bwd'],4), round(d['tflops'],1), d['loss'], d['anchor'])
PY
python tools/ntxent_profile.py 8192 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ntxent_launches.csv python tools/ntxent_profile.py 8192 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r02_ntxent_launches.csv') if not l.startswith('=='))]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
for r in rows[-8:]: print(r[ki][:60], r[vi], r[ui])
PY
