# end-of-round evidence run on one B200: GPU tests, smoke, the four bench workloads, ncu launch lists + captures
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_h1.json 2> gpurun_out/bench_h1.err; tail -c 300 gpurun_out/bench_h1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_h1_ref.json 2> gpurun_out/bench_h1_ref.err; tail -c 600 gpurun_out/bench_h1_ref.json
timeout 300 python bench.py --workload v1 > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; tail -c 300 gpurun_out/bench_v1.json
timeout 600 python bench.py --workload t1 --steps 20 > gpurun_out/bench_t1.json 2> gpurun_out/bench_t1.err; tail -c 300 gpurun_out/bench_t1.json
bash profiles/run_ncu.sh > gpurun_out/run_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
