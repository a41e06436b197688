#!/usr/bin/env bash
# round-2 confirmation: GPU tests, smoke, the bench line of every BASELINE.json configuration (with the eager-bf16 reference and CPU reference legs), the reference arm alone
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2f; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -rs > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/cfg2.json 2> $O/cfg2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/ref_cpu.json 2> $O/ref_cpu.err
timeout 600 python bench.py --impl reference --device cuda --steps 10 --warmup 3 > $O/ref_cuda.json 2> $O/ref_cuda.err
for wl in cfg1 cfg3 cfg4 cfg5:224 cfg5:672 cfg5:896 cfg5:1344; do
  f=$O/$(echo $wl | tr ':' '_')
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > $f.json 2> $f.err
done
tail -c 300 $O/*.err
