#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
( time timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -c 6000 $O/r02_bench_n1.json; tail -5 $O/r02_bench_n1.err
( time timeout 300 python bench.py --workload shor > $O/r02_bench_shor.json 2> $O/r02_bench_shor.err ) 2>&1 | grep real; tail -c 4000 $O/r02_bench_shor.json; tail -5 $O/r02_bench_shor.err
timeout 200 python tools/run_streaming_kernels.py > $O/r02_streaming_kernels.json 2>&1; cat $O/r02_streaming_kernels.json
