#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 200 python tools/run_qft_variants.py 30 10 444,64 > $O/r02_pair_variants_b_n30.log 2>&1; echo "variants rc=$?"; cat $O/r02_pair_variants_b_n30.log
timeout 200 python tools/run_qft_variants.py 33 5 444 > $O/r02_pair_variants_b_n33.log 2>&1; echo "n=33 rc=$?"; cat $O/r02_pair_variants_b_n33.log
QCS_LIB_PATH=build/timing/libqcs.so QCS_PIPE_TIMING=1 timeout 120 python tools/run_qft.py 30 0 1 > $O/r02_pipe_role_timing_b_n30.log 2>&1; echo "timing rc=$?"; cat $O/r02_pipe_role_timing_b_n30.log
timeout 120 python tools/run_qft.py 30 0 2 > $O/plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_qft_sweep_tma --csv --log-file $O/r02_ncu_pair_dram_b_n30.csv python tools/run_qft.py 30 0 2 > $O/ncu.log 2>&1; echo "ncu rc=$?"; grep -E "dram__bytes|gpu__time" $O/r02_ncu_pair_dram_b_n30.csv | cut -d, -f1,13- | tail -12
