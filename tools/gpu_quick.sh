#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/quick; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -12 $O/pytest.log
for wl in cfg1 cfg2; do
timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu > $O/$wl.json 2> $O/$wl.err; tail -c 300 $O/$wl.err
python - $wl <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/quick/{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ("value","ms_per_step","step_tflops","launches_per_step")}, d["e2e"]["value"], d.get("gpu_eager_baseline",{}).get("ours_over_eager"))
PY
done
