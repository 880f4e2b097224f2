set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cvt.csv \
    python bench.py --workload h1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cvt_stdout.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:emb_gemm_cvt_kernel -s 5 -c 1 -f -o gpurun_out/prof_cvt \
    python bench.py --workload h1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cvt_stdout.log 2>&1
ls -la gpurun_out/prof_cvt.ncu-rep
