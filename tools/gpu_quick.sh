#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/quick; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager > $O/cfg2.json 2> $O/cfg2.err; tail -c 300 $O/cfg2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/quick/cfg2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","step_tflops")}, d["e2e"]["value"], d["clocks"])
for k,v in d["kernels"].items(): print("   ",k,v)
PY
