"""measure_state at n qubits on a synthetic state and on a Shor state (the command profiled for the scan kernels)."""
import math
import sys
import time

sys.path.insert(0, ".")
import quantumcomputer_b200 as q

L = int(sys.argv[1]) if len(sys.argv) > 1 else 18
M = int(sys.argv[2]) if len(sys.argv) > 2 else 12
with q.Register(L, M) as reg:
    for name in ("synthetic", "shor"):
        for r in (0.25, 0.9):
            if name == "synthetic":
                reg.fill_synthetic(1234)
                reg.scale(1.0 / math.sqrt(reg.norm2()))
            else:
                reg.reset_register()
                reg.quantum_computation(4087 if M >= 12 else 21, 7 if M >= 12 else 2, q.POW_MODULAR)
            reg.norm2()                      # flushes anything deferred
            reg.timer_start()
            t0 = time.perf_counter()
            idx = reg.measure_state(r)
            dt = time.perf_counter() - t0
            ev = reg.timer_stop()
            print(f"{name} r={r}: index {idx}, measure_state {1e3 * dt:.3f} ms (host clock), {ev:.3f} ms (CUDA events)", flush=True)
