#!/bin/bash
mkdir -p gpurun_out
./tools/gpu_first.sh "$@"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'step_tflops',d['step_tflops'],'e2e',d['e2e']['value'], 'roofline',d['roofline']['achieved'],d['roofline']['frac'])
for k,v in d['kernels'].items(): print('  ',k,v)
PY
tail -3 gpurun_out/bench.err
