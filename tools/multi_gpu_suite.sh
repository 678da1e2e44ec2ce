#!/bin/bash
# Multi-GPU evidence run (inside gpurun --gpus N): sharded parity worker, the weak-scaling bench
# point (n = 30 + log2 N), the strong-scaling n = 33 point and, on 8 GPUs, the n = 35 point of
# BASELINE configs[4].  Usage: tools/multi_gpu_suite.sh N
N=$1
OUT=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
free -g | head -2 > $OUT/mg_${N}_host.txt
nvidia-smi topo -m > $OUT/mg_${N}_topo.txt 2>&1
timeout 900 $RUN --master-port 29511 tests/dist_worker.py > $OUT/mg_${N}_dist.log 2>&1
grep -c "ok" $OUT/mg_${N}_dist.log; grep "DIST_\|FAIL\|peer memory" $OUT/mg_${N}_dist.log | tail -8
avail=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
E2E="--e2e-steps 2"
if [ "$avail" -lt $((N * 40)) ]; then E2E="--no-e2e"; fi
timeout 600 $RUN --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 $E2E > $OUT/mg_${N}_bench_weak.json 2> $OUT/mg_${N}_bench_weak.err
cut -c1-260 $OUT/mg_${N}_bench_weak.json
timeout 600 $RUN --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --qubits 33 --no-e2e > $OUT/mg_${N}_bench_n33.json 2> $OUT/mg_${N}_bench_n33.err
cut -c1-260 $OUT/mg_${N}_bench_n33.json
if [ "$N" -ge 4 ]; then
  timeout 600 $RUN --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --qubits 35 --no-e2e > $OUT/mg_${N}_bench_n35.json 2> $OUT/mg_${N}_bench_n35.err
  cut -c1-260 $OUT/mg_${N}_bench_n35.json
fi
if [ "$N" -eq 2 ]; then
  timeout 600 $RUN --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --qubits 34 --no-e2e > $OUT/mg_${N}_bench_n34.json 2> $OUT/mg_${N}_bench_n34.err
  cut -c1-260 $OUT/mg_${N}_bench_n34.json
fi
