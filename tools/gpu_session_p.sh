#!/bin/bash
# round 2, re-entry: the measurement-scan parity tests on the final measure.cu (DUAL chunks), 1 GPU
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 70 python -m pytest tests/test_gates_gpu.py -x -q -k "parallel_measurement" > $O/r02_pytest_measure_dual.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_measure_dual.log
