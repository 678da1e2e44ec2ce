#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 200 python tools/run_qft_variants.py 30 20 444 > $O/r02_variants_e_n30.log 2>&1; echo "rc=$?"; cut -c1-200 $O/r02_variants_e_n30.log
timeout 200 python tools/run_qft_variants.py 33 5 444 > $O/r02_variants_e_n33.log 2>&1; echo "rc=$?"; cut -c1-200 $O/r02_variants_e_n33.log
QCS_LIB_PATH=build/timing/libqcs.so QCS_PIPE_TIMING=1 timeout 120 python tools/run_qft.py 30 0 1 > $O/r02_pipe_role_timing_e_n30.log 2>&1; cat $O/r02_pipe_role_timing_e_n30.log | head -4
tools/gpu_session_ncu.sh
