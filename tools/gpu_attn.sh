#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/attn; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for f in 0 1; do
echo "== fuse $f"
VF_ATT_FUSE=$f timeout 300 python tools/attn_bench.py 64 784 12 16 6272 12 256 196 12 28 1764 12 2>&1 | cut -c15-90
done
VF_ATTN_FLAGS=2 timeout 120 python tools/attn_trace.py 64 784 12 0 30 0 > $O/trace.txt 2>&1; grep -E "^w00 step +(1[1-5]|2[4-8]):" $O/trace.txt | sed 's/start [0-9]* //' | cut -c1-200
