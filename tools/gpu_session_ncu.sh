#!/bin/bash
# ncu evidence (one GPU): streaming kernels, the three sweeps of the n = 30 inverse QFT, the bench launch list
cd "$(dirname "$0")/.."
O=gpurun_out
K='regex:k_hadamard_exact|k_phase_masked|k_amodc|k_modexp_sweep|k_norm2_partial|k_chunk_sums|k_chunk_maps|k_exact_walk|k_basis_state|k_scale'
timeout 200 python tools/run_streaming_kernels.py 16 12 > $O/plain_stream.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -c 40 -o $O/r02_streaming -f python tools/run_streaming_kernels.py 16 12 > $O/ncu_stream.log 2>&1; echo "ncu streaming rc=$?"; tail -3 $O/ncu_stream.log
timeout 120 python tools/run_qft.py 30 0 1 > $O/plain_qft.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qft_sweep_tma -s 3 -c 3 -o $O/r02_qft_sweeps_n30 -f python tools/run_qft.py 30 0 1 > $O/ncu_qft.log 2>&1; echo "ncu qft rc=$?"; tail -3 $O/ncu_qft.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-north-star --no-configs > $O/plain_bench.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launches_bench_n30.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-north-star --no-configs > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
ls -la $O/*.ncu-rep
