#!/usr/bin/env bash
# round-2: tests after the single-protocol folded LayerNorm / new surface kernels, smoke, one bench line
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -30 $O/pytest.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -3 $O/smoke.log

