#!/bin/bash
# every kernel group in its own process (a trap in one must not poison the rest)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest -q -m gpu -p no:cacheprovider -s "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
for g in "$@"; do
  case $g in
    gemm) run gemm tests/test_gpu_kernels.py -k "gemm";;
    patch) run patch tests/test_gpu_kernels.py -k "patch_embed";;
    attn) run attn tests/test_gpu_kernels.py -k "attention";;
    misc) run misc tests/test_gpu_kernels.py -k "not gemm and not patch_embed and not attention";;
    models) run models tests/test_gpu_models.py;;
  esac
done
cat gpurun_out/summary.txt
