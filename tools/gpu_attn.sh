#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/attn; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 300 python tools/attn_bench.py 64 784 12 16 6272 12 8 197 12 64 196 12 8 7056 12 32 3136 12 > $O/v1d.txt 2>&1; cat $O/v1d.txt
