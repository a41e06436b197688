#!/usr/bin/env bash
# eight GPUs: bench at N=8 for cfg2 / cfg5:1344 / cfg4 / cfg3 (weak scaling, fused all-gather checked against NCCL inside bench.py)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-8}
O=gpurun_out/r2n$N; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
for wl in ${WLS:-cfg2 cfg5:1344 cfg4 cfg3}; do
  f=$O/n${N}_$(echo $wl | tr ':' '_')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --workload $wl > $f.json 2> $f.err
  tail -c 1200 $f.json; tail -c 300 $f.err
done
