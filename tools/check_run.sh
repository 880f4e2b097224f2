set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py --cpu-seconds 6 > gpurun_out/bench_h1.json 2> gpurun_out/bench_h1.err; tail -c 200 gpurun_out/bench_h1.json
