#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python tools/run_qft_variants.py 30 10 444 > $O/r02_variants_c_n30.log 2>&1; echo "variants rc=$?"; cat $O/r02_variants_c_n30.log
timeout 300 python tools/run_qft_variants.py 33 5 444 > $O/r02_variants_c_n33.log 2>&1; echo "n=33 rc=$?"; cat $O/r02_variants_c_n33.log
timeout 200 python tools/run_qft_variants.py 26 20 444 > $O/r02_variants_c_n26.log 2>&1; echo "n=26 rc=$?"; cat $O/r02_variants_c_n26.log
