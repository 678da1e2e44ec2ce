#!/bin/bash
# Host-side sanitizer pass (runs without a GPU):
#  1. the C host driver under ASan + UBSan, its libqcs.so calls answered by tests/mock (the CPU oracle);
#  2. libqcs.so's host code built with -fsanitize=undefined (build/san/libqcs.so, ~4 min), the host-only
#     CPU tests and tools/fuzz_host_entry_points.py (planner, gate-stream scheduler, pair bookkeeping) on it.
# Device code is not covered: compute-sanitizer is closed on this pool.
set -e
cd "$(dirname "$0")/.."
S=build/san
mkdir -p $S
gcc -O1 -g -Wall -Wextra -fsanitize=undefined,address -fno-omit-frame-pointer -Iinclude -o $S/qc_shor_b200_san \
    quantumcomputer_b200/host/qc_shor_b200.c quantumcomputer_b200/host/mt19937.c quantumcomputer_b200/host/shor_classical.c \
    -Lquantumcomputer_b200/lib -lqcs -Wl,-rpath,$PWD/quantumcomputer_b200/lib -lm
gcc -O2 -fPIC -shared -Wall -Iinclude -Ioracle -o $S/libqcs_mock.so tests/mock/mock_qcs.c -Loracle/_build -lqcsoracle -Wl,-rpath,$PWD/oracle/_build
for args in "-C 15 -L 3 -M 4 -a 7 -s 1 -V" "-C 21 -L 4 -M 5 -s 7 -v" "-C 77 -L 13 -M 7 -a 2 -s 1 -r" "-C 33 -L 11 -M 6 -s 2 -r -v" \
            "-C 15 -L 3 -M 4 -a 14 -s 3" "-C 21 -L 5 -M 5 -a 4 -s 10 -V" "-C 33 -L 5 -M 6 -s 4"; do
    set +e
    ASAN_OPTIONS=verify_asan_link_order=0:detect_leaks=0 LD_PRELOAD="$(gcc -print-file-name=libasan.so):$PWD/$S/libqcs_mock.so" \
        $S/qc_shor_b200_san $args > $S/out.txt 2> $S/err.txt
    rc=$?
    set -e
    echo "host driver $args: rc=$rc, sanitizer reports: $(grep -c 'runtime error\|ERROR: AddressSanitizer' $S/err.txt || true)"
done
if [ ! -f $S/libqcs.so ] || [ -n "$(find quantumcomputer_b200/csrc -newer $S/libqcs.so -print -quit)" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 -Xcompiler -fPIC --fmad=true -Xcompiler -fsanitize=undefined \
         -shared -o $S/libqcs.so quantumcomputer_b200/csrc/*.cu -ldl -lubsan
fi
QCS_LIB_PATH=$PWD/$S/libqcs.so python -m pytest tests/test_scheduler_cpu.py tests/test_abi.py tests/test_bench_cpu.py -q -s 2>&1 | tee $S/pytest.log | tail -2
QCS_LIB_PATH=$PWD/$S/libqcs.so python tools/fuzz_host_entry_points.py 2>&1 | tee $S/fuzz.log | tail -3
echo "UBSan reports in libqcs host code: $(cat $S/pytest.log $S/fuzz.log | grep -c 'runtime error' || true)"
