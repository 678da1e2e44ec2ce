#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for rep in 1 2; do
  (cd build/r01tree && timeout 120 python tools/run_qft.py 30 0 20) 2>&1 | head -1 | sed 's/^/r01 n=30: /'
  timeout 120 python tools/run_qft.py 30 0 20 2>&1 | head -1 | sed 's/^/now n=30: /'
done
(cd build/r01tree && timeout 120 python tools/run_qft.py 33 0 5) 2>&1 | head -1 | sed 's/^/r01 n=33: /'
timeout 120 python tools/run_qft.py 33 0 5 2>&1 | head -1 | sed 's/^/now n=33 (paired): /'
timeout 200 python tools/run_qft_variants.py 30 20 444 2>&1 | cut -c1-110
QCS_LIB_PATH=build/timing/libqcs.so QCS_PIPE_TIMING=1 timeout 120 python tools/run_qft.py 30 0 1 2>&1 | head -3
