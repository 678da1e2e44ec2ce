#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 200 python tools/run_qft_variants.py 30 20 444 > $O/r02_variants_f_n30.log 2>&1; echo "rc=$?"; cut -c1-120 $O/r02_variants_f_n30.log
timeout 200 python tools/run_qft_variants.py 33 5 444 > $O/r02_variants_f_n33.log 2>&1; echo "rc=$?"; cut -c1-120 $O/r02_variants_f_n33.log
QCS_LIB_PATH=build/timing/libqcs.so QCS_PIPE_TIMING=1 timeout 120 python tools/run_qft.py 30 0 1 > $O/r02_pipe_role_timing_f_n30.log 2>&1; head -3 $O/r02_pipe_role_timing_f_n30.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1_f.json 2> $O/r02_bench_n1_f.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -c 2500 $O/r02_bench_n1_f.json; tail -5 $O/r02_bench_n1_f.err
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_ref_f.json 2> $O/r02_bench_ref_f.err ) 2>&1 | grep real; cut -c1-400 $O/r02_bench_ref_f.json
