#!/bin/bash
# one GPU session: paired-sweep variants, role timing, DRAM traffic, GPU tests, access-pattern probes
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 200 python tools/run_qft_variants.py 30 10 444,148,0,1024 > $O/r02_pair_variants_n30.log 2>&1; echo "variants rc=$?"; cat $O/r02_pair_variants_n30.log
QCS_NO_SPLIT3=1 timeout 120 python tools/run_qft_variants.py 30 10 444 > $O/r02_pair_variants_n30_nosplit.log 2>&1; echo "nosplit rc=$?"; cat $O/r02_pair_variants_n30_nosplit.log
for n in 31 33; do timeout 200 python tools/run_qft_variants.py $n 5 444 > $O/r02_pair_variants_n$n.log 2>&1; echo "n=$n rc=$?"; cat $O/r02_pair_variants_n$n.log; done
QCS_LIB_PATH=build/timing/libqcs.so QCS_PIPE_TIMING=1 timeout 120 python tools/run_qft.py 30 0 1 > $O/r02_pipe_role_timing_pair_n30.log 2>&1; echo "timing rc=$?"; cat $O/r02_pipe_role_timing_pair_n30.log
timeout 120 python tools/run_qft.py 30 0 2 > $O/plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum --clock-control none -k regex:k_qft_sweep_tma --csv --log-file $O/r02_ncu_pair_dram_n30.csv python tools/run_qft.py 30 0 2 > $O/ncu.log 2>&1; echo "ncu rc=$?"; grep -E "dram__bytes|gpu__time" $O/r02_ncu_pair_dram_n30.csv | cut -d, -f1,5,13- | tail -24
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 100 build/ubench_l2 5 > $O/ubench_order.log 2>&1; cat $O/ubench_order.log
timeout 100 build/ubench_l2 6 > $O/ubench_rows.log 2>&1; cat $O/ubench_rows.log
