#!/usr/bin/env bash
# two GPUs: the whole GPU suite (incl. the fused all-gather == NCCL test), bench at N=2 for cfg2 / cfg5:1344 / cfg4
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2n2; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests -m gpu -q -rs > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -6 $O/pytest.log
for wl in cfg2 cfg5:1344 cfg4; do
  f=$O/n2_$(echo $wl | tr ':' '_')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --workload $wl > $f.json 2> $f.err
  tail -c 1500 $f.json; tail -c 400 $f.err
done
