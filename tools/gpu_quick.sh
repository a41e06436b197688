#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/quick; mkdir -p $O
for st in 20 60; do
timeout 300 python bench.py --steps $st --warmup 5 --no-cpu --no-eager > $O/s$st.json 2> $O/s$st.err
python - $st <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/quick/s{sys.argv[1]}.json').read().strip().splitlines()[-1])
print("steps", sys.argv[1], {k:d.get(k) for k in ("value","ms_per_step","kernel_time_share_of_step")}, "e2e", d["e2e"]["ms_per_step"], d["clocks"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["peak_regime"])
PY
done
