#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import math, sys, json
sys.path.insert(0, ".")
import quantumcomputer_b200 as q
from quantumcomputer_b200.workloads import apply_gates, layered_circuit
# n = 30: L2 pairs with 32 MiB blocks again, now that the contiguous tile's steps are shorter
with q.Register(30, 0) as reg:
    reg.fill_synthetic(1234); reg.scale(1.0 / math.sqrt(reg.norm2()))
    for name, opts in (("pairs <= 16 MiB (default)", {q.OPT_L2_PAIR_MAX_BLOCK: 16 << 20}), ("pairs <= 32 MiB", {q.OPT_L2_PAIR_MAX_BLOCK: 32 << 20}),
                       ("pairs <= 32 MiB, lag 148", {q.OPT_L2_PAIR_MAX_BLOCK: 32 << 20, q.OPT_L2_PAIR_LAG: 148})):
        for k, v in opts.items(): reg.set_option(k, v)
        for _ in range(3): reg.inverse_QFT()
        reg.synchronize(); reg.timer_start()
        for _ in range(20): reg.inverse_QFT()
        ms = reg.timer_stop() / 20
        reg.timer_start()
        for _ in range(20): reg.QFT()
        print(f"n=30 {name}: iqft {ms:.3f} ms qft {reg.timer_stop() / 20:.3f} ms", flush=True)
# layered circuit n = 33 with and without paired launches
n = 33
circuit = layered_circuit(n, 8)
with q.Register(n, 0) as reg:
    reg.fill_synthetic(1234); reg.scale(1.0 / math.sqrt(reg.norm2()))
    for name, pair in (("pair on", 1), ("pair off", 0)):
        reg.set_option(q.OPT_L2_PAIR, pair)
        with reg.fused(): apply_gates(reg, circuit)
        reg.synchronize(); before = reg.launch_count; reg.timer_start()
        for _ in range(2):
            with reg.fused(): apply_gates(reg, circuit)
        ms = reg.timer_stop() / 2
        print(f"layered n=33 {name}: {ms:.1f} ms per step, {(reg.launch_count - before) // 2} launches, norm {reg.norm2():.15f}", flush=True)
PY
