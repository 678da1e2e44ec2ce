#!/bin/bash
# round 2, re-entry: sampling without collapse (the RECORD walk) on the final measure.cu, 1 GPU
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 16 python -m pytest tests/test_extensions_gpu.py -x -q -k "sampling_without_collapse" > $O/r02_pytest_sampling_dual.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_sampling_dual.log
