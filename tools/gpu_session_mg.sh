#!/bin/bash
# inside `gpurun --gpus N`: sharded parity (one process per GPU and one process for all GPUs), then the bench at N
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_mg${N}_bench.json 2> $O/r02_mg${N}_bench.err; echo "bench rc=$?"; tail -c 7000 $O/r02_mg${N}_bench.json; tail -5 $O/r02_mg${N}_bench.err
