"""One launch (or a few) of every streaming kernel class north_star names, for ncu captures:
pair-stride Hadamard, controlled phase, modular-exponentiation permutation, norm^2 reduction and the
exact measurement scan.  Also prints CUDA-event GB/s per class (not under ncu).

    python tools/run_streaming_kernels.py [L] [M]      (default L = 18, M = 12: n = 30)
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantumcomputer_b200 as q  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    n = L + M
    with q.Register(L, M) as reg:
        reg.fill_synthetic(7)
        reg.scale(1.0 / math.sqrt(reg.norm2()))
        reg.set_option(q.OPT_FUSION, 0)
        reg.set_option(q.OPT_PROFILE, 1)
        reg.profile_reset()
        for qb in (0, 2, 5, 13, n - 1):
            reg.hadamard_gate(qb)                              # k_hadamard_exact
        for c, t in ((n - 1, n - 2), (n - 1, 3), (7, 1)):
            reg.c_phase_shift_gate(c, t, math.pi / 8)          # k_phase_masked
        reg.c_amodc_gate(4087, 7, n - 1)                       # k_amodc (one controlled gate)
        reg.set_option(q.OPT_FUSION, 1)
        reg.fill_synthetic(7)
        reg.scale(1.0 / math.sqrt(reg.norm2()))                # k_norm2_partial
        reg.quantum_computation(4087, 7, q.POW_MODULAR)        # Walsh sweeps + k_modexp_sweep + QFT sweeps
        nrm = reg.norm2()
        idx = reg.measure_state(0.7310585786)                  # k_chunk_sums / k_classify / k_chunk_maps / k_exact_walk
        prof = reg.profile()
        out = {k: {"launches": v[0], "ms": round(v[1], 4), "GBps": round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None}
               for k, v in prof.items() if v[0]}
        print(json.dumps({"n": n, "norm": nrm, "measured": int(idx), "classes": out}))


if __name__ == "__main__":
    main()
