#!/bin/bash
# Reduced multi-GPU run: parity worker + weak-scaling bench point + (N >= 4) the n = 35 point.
N=$1
OUT=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29511 tests/dist_worker.py > $OUT/mgq_${N}_dist.log 2>&1
grep "DIST_\|FAIL" $OUT/mgq_${N}_dist.log | tail -4
timeout 300 $RUN --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > $OUT/mgq_${N}_bench_weak.json 2> $OUT/mgq_${N}_bench_weak.err
cut -c1-200 $OUT/mgq_${N}_bench_weak.json
if [ "$N" -ge 4 ]; then
  timeout 300 $RUN --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --qubits 35 --no-e2e > $OUT/mgq_${N}_bench_n35.json 2> $OUT/mgq_${N}_bench_n35.err
  cut -c1-200 $OUT/mgq_${N}_bench_n35.json
fi
