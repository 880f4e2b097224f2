set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/sharded_worker.py > gpurun_out/sharded_worker.log 2>&1
grep -v "^W0\|^\[W" gpurun_out/sharded_worker.log | tail -40
