#!/bin/bash
# Profiling recipe (B200_PROFILING.md), run on the GPU box under gpurun, 1 GPU.
# Outputs land in gpurun_out/; summaries are copied to profiles/ (profiles/summarize.py) and committed.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# (1) launch lists: every kernel with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_h1.csv \
    python bench.py --workload h1 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_h1_stdout.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_v1.csv \
    python bench.py --workload v1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v1_stdout.log 2>&1
# (2) full captures of the top kernels (same command as the bench)
# the sweep is the 2nd launch of emb_gemm_cvt_kernel in a step (the 1st is the one-tile threshold pass): skip 3 steps + 1
ncu --set full --clock-control none --import-source on -k regex:emb_gemm_cvt_kernel -s 7 -c 1 -f -o gpurun_out/prof_gemm_h1 \
    python bench.py --workload h1 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
OC_SIDE_STREAM=0 ncu --set full --clock-control none --import-source on -k regex:bm25_warp_kernel -s 4 -c 1 -f -o gpurun_out/prof_bm25_h1 \
    python bench.py --workload h1 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:emb_gemm_merge_kernel -s 4 -c 1 -f -o gpurun_out/prof_merge_h1 \
    python bench.py --workload h1 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:emb_scan_kernel -s 3 -c 1 -f -o gpurun_out/prof_scan_v1 \
    python bench.py --workload v1 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/
