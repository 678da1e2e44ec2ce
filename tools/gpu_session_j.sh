#!/bin/bash
cd "$(dirname "$0")/.."
python - <<'PY'
import math, sys
sys.path.insert(0, ".")
import quantumcomputer_b200 as q
for n, reps in ((30, 20), (31, 10), (32, 8), (33, 5), (28, 40)):
    with q.Register(n, 0) as reg:
        reg.fill_synthetic(1234); reg.scale(1.0 / math.sqrt(reg.norm2()))
        for name, opts in (("pair off", {q.OPT_L2_PAIR: 0}),
                           ("pairs <= 16 MiB lag 444", {q.OPT_L2_PAIR: 1, q.OPT_L2_PAIR_MAX_BLOCK: 16 << 20, q.OPT_L2_PAIR_LAG: 444}),
                           ("pairs <= 32 MiB lag 444", {q.OPT_L2_PAIR_MAX_BLOCK: 32 << 20, q.OPT_L2_PAIR_LAG: 444}),
                           ("pairs <= 32 MiB lag 148", {q.OPT_L2_PAIR_MAX_BLOCK: 32 << 20, q.OPT_L2_PAIR_LAG: 148}),
                           ("pairs <= 32 MiB lag 296", {q.OPT_L2_PAIR_MAX_BLOCK: 32 << 20, q.OPT_L2_PAIR_LAG: 296}),
                           ("pairs <= 32 MiB lag 64", {q.OPT_L2_PAIR_MAX_BLOCK: 32 << 20, q.OPT_L2_PAIR_LAG: 64})):
            for k, v in opts.items(): reg.set_option(k, v)
            for _ in range(3): reg.inverse_QFT()
            reg.synchronize(); reg.timer_start()
            for _ in range(reps): reg.inverse_QFT()
            ms = reg.timer_stop() / reps
            reg.timer_start()
            for _ in range(reps): reg.QFT()
            print(f"n={n} {name}: iqft {ms:.3f} ms qft {reg.timer_stop() / reps:.3f} ms", flush=True)
PY
