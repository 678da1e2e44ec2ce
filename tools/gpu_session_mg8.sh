#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-8}
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29620 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_mg${N}_bench.json 2> $O/r02_mg${N}_bench.err; echo "bench rc=$?"; tail -c 6500 $O/r02_mg${N}_bench.json; tail -3 $O/r02_mg${N}_bench.err
timeout 400 $RUN --master-port 29621 tools/mg_tune.py $2 5 > $O/r02_mg${N}_tune.jsonl 2> $O/r02_mg${N}_tune.err; echo "tune rc=$?"; cat $O/r02_mg${N}_tune.jsonl; tail -3 $O/r02_mg${N}_tune.err
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q -k "single_process or c_host" 2>&1 | tail -4
