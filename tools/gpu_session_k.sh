#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python tools/sweep_matrix.py 30 6 0,1,2,3,4,5 > $O/r02_tuning_matrix_shapes_n30.jsonl 2>&1; cut -c1-200 $O/r02_tuning_matrix_shapes_n30.jsonl
timeout 600 python tools/sweep_matrix.py 33 3 0,3,4 > $O/r02_tuning_matrix_shapes_n33.jsonl 2>&1; cut -c1-200 $O/r02_tuning_matrix_shapes_n33.jsonl
timeout 120 python tools/run_qft.py 30 0 1 > $O/plain_qft.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qft_sweep_tma -s 3 -c 3 -o $O/r02_qft_sweeps_final_n30 -f python tools/run_qft.py 30 0 1 > $O/ncu_qft.log 2>&1; echo "ncu qft rc=$?"
timeout 120 python tools/run_qft.py 33 0 1 > $O/plain_qft33.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qft_sweep_tma -s 2 -c 2 -o $O/r02_qft_pairs_n33 -f python tools/run_qft.py 33 0 1 > $O/ncu_qft33.log 2>&1; echo "ncu qft33 rc=$?"
