#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/attn; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 300 python tools/attn_bench.py 64 784 12 16 6272 12 8 197 12 256 196 12 28 1764 12 32 3136 12 8 7056 12 > $O/bench.txt 2>&1; cat $O/bench.txt
VF_ATTN_FLAGS=2 timeout 120 python tools/attn_trace.py 64 784 12 0 30 0 1 8 > $O/trace.txt 2>&1; grep -E "item end|S\(0\)|mean" $O/trace.txt | cut -c1-200
