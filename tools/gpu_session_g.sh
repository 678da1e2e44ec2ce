#!/bin/bash
# same-box A/B: the round-1 library (build/r01tree) against the current one, with and without the split-3 layout
cd "$(dirname "$0")/.."
O=gpurun_out
for rep in 1 2; do
  (cd build/r01tree && timeout 120 python tools/run_qft.py 30 0 20) 2>&1 | head -1 | sed 's/^/r01 n=30: /'
  timeout 120 python tools/run_qft.py 30 0 20 2>&1 | head -1 | sed 's/^/now n=30: /'
done
(cd build/r01tree && timeout 120 python tools/run_qft.py 33 0 5) 2>&1 | head -1 | sed 's/^/r01 n=33: /'
timeout 120 python tools/run_qft.py 33 0 5 2>&1 | head -1 | sed 's/^/now n=33 (paired): /'
python - <<'PY'
import math, sys
sys.path.insert(0, ".")
import quantumcomputer_b200 as q
for n, reps in ((30, 20), (33, 5)):
    with q.Register(n, 0) as reg:
        reg.fill_synthetic(1234); reg.scale(1.0 / math.sqrt(reg.norm2()))
        for name, opts in (("pair off split on", {q.OPT_L2_PAIR: 0, q.OPT_SPLIT3: 1}), ("pair off split off", {q.OPT_L2_PAIR: 0, q.OPT_SPLIT3: 0}),
                           ("pair on split on", {q.OPT_L2_PAIR: 1, q.OPT_SPLIT3: 1}), ("pair on split off", {q.OPT_L2_PAIR: 1, q.OPT_SPLIT3: 0})):
            for k, v in opts.items(): reg.set_option(k, v)
            for _ in range(3): reg.inverse_QFT()
            reg.synchronize(); reg.timer_start()
            for _ in range(reps): reg.inverse_QFT()
            ms = reg.timer_stop() / reps
            reg.timer_start()
            for _ in range(reps): reg.QFT()
            msf = reg.timer_stop() / reps
            print(f"now n={n} {name}: iqft {ms:.3f} ms  qft {msf:.3f} ms  norm {reg.norm2():.16f}", flush=True)
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
