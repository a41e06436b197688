#!/usr/bin/env bash
# round-2 first measurement: tests, bench lines per workload with the eager-bf16 reference next to them
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2a; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
python bench.py --steps 10 --warmup 3 > $O/cfg2.json 2> $O/cfg2.err
python bench.py --steps 10 --warmup 3 --no-graph --no-cpu --no-eager > $O/cfg2_nograph.json 2> $O/cfg2_nograph.err
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu > $O/cfg4.json 2> $O/cfg4.err
python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu > $O/cfg3.json 2> $O/cfg3.err
python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu > $O/cfg1.json 2> $O/cfg1.err
python bench.py --workload cfg5:1344 --steps 5 --warmup 3 --no-cpu > $O/cfg5_1344.json 2> $O/cfg5_1344.err
python tools/attn_vs_sdpa.py > $O/attn_vs_sdpa.txt 2>&1
tail -c 600 $O/*.err
