set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_h1.json 2> gpurun_out/bench_h1.err; tail -c 200 gpurun_out/bench_h1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_h1_ref.json 2> gpurun_out/bench_h1_ref.err
