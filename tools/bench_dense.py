"""Micro-benchmark behind DESIGN.md section 4.1: the inverse QFT on the 4 (or 3) low qubits
at n qubits, three ways: (a) one dense 2^k x 2^k block on FP64 tensor cores (DMMA),
(b) one fused tile sweep with register butterflies, (c) gate by gate.
Prints one JSON line; run on the GPU box."""
import json
import math
import sys

import numpy as np

sys.path.insert(0, ".")
import quantumcomputer_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
out = {"n": n, "reps": reps}
with q.Register(n, 0) as reg:
    reg.fill_synthetic(1234)
    reg.scale(1.0 / math.sqrt(reg.norm2()))
    for k in (4, 3):
        R = 1 << k
        # the k-qubit inverse-QFT matrix, built with the engine itself on a small register
        U = np.zeros((R, R), dtype=np.complex128)
        with q.Register(max(k, 6), 0) as small:
            small.set_option(q.OPT_FUSION, 0)
            for j in range(R):
                e = np.zeros(1 << max(k, 6), dtype=np.complex128)
                e[j] = 1.0
                small.set_state(e)
                small.inverse_QFT(0, k)
                U[:, j] = small.get_state()[:R]
        gates = k + k * (k - 1) // 2
        for name, fn in (("dense_dmma", lambda: reg.apply_dense_block(k, U)),
                         ("fused_sweep", lambda: reg.inverse_QFT(0, k)),
                         ("gate_by_gate", None)):
            if name == "gate_by_gate":
                reg.set_option(q.OPT_FUSION, 0)
                fn = lambda: reg.inverse_QFT(0, k)
            fn()
            reg.synchronize()
            reg.timer_start()
            for _ in range(reps):
                fn()
            ms = reg.timer_stop() / reps
            reg.set_option(q.OPT_FUSION, 1)
            gb = 32.0 * (1 << n) / 1e9
            out[f"k{k}_{name}"] = {"ms": round(ms, 4), "gates_per_s": round(gates / (ms * 1e-3), 1),
                                   "sweep_GBps": round(gb / (ms * 1e-3), 1) if name != "gate_by_gate" else None}
print(json.dumps(out))
