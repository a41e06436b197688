#!/usr/bin/env bash
# round-2 profiles: ncu launch list of the bench command, full captures of the top kernels, sanitizer passes
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2p; mkdir -p $O
# (1) every launch of one eager step with its device time (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 100 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-eager --no-graph > $O/launches_bench.log 2>&1
tail -2 $O/launches_bench.log | cut -c1-300
# (2) full captures: the GEMM family of one layer + merger, attention, row stats
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 60 -c 8 -o $O/gemm -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-eager --no-graph > $O/ncu_gemm.log 2>&1; tail -2 $O/ncu_gemm.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 12 -c 1 -o $O/attn -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-eager --no-graph > $O/ncu_attn.log 2>&1; tail -2 $O/ncu_attn.log | cut -c1-200
timeout 900 ncu --set full --clock-control none -k regex:ln_row_stats -s 12 -c 1 -o $O/rowstats -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-eager --no-graph > $O/ncu_rowstats.log 2>&1
# (3) sanitizers over the kernel tests (every kernel incl. the round-2 ones, ragged shapes)
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $O/san_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/san_memcheck.log; tail -5 $O/san_memcheck.log
timeout 1200 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm or attention or layernorm or gelu or embed or tiny" > $O/san_synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/san_synccheck.log; tail -5 $O/san_synccheck.log
ls -la $O
