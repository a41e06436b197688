#!/usr/bin/env bash
# GPU tests + the bench line of a few configurations (cfg2 by default)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/quick; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for wl in ${@:-cfg2}; do
  f=$O/$(echo $wl | tr ':' '_')
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu > $f.json 2> $f.err
  python - "$f.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=d.get("kernels",{})
print(d["config"]["workload"][:40], "value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),"eager",(d.get("gpu_eager_baseline") or {}).get("value"),"attn",k.get("attention",{}).get("ms_per_step"),k.get("attention",{}).get("tflops"),"clk",d["clocks"]["sm_mhz"])
PY
done
