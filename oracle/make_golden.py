#!/usr/bin/env python
"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libqcref.so).

Run from the repo root in a container where /root/reference exists:

    make -C oracle && python oracle/make_golden.py

Arrays are stored as hex of little-endian float64 so that comparisons can be
bit-exact.  Test infrastructure only.
"""
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import Reference, Restatement  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def hexs(a):
    return np.ascontiguousarray(a).view(np.uint8).tobytes().hex()


def synthetic_state(n, seed):
    """The benchmark's counter-based state, normalised with numpy."""
    k = np.arange(2 * (1 << n), dtype=np.uint64)
    u = np.array([Restatement.synthetic_u(seed, int(x)) for x in k])
    amps = u[0::2] + 1j * u[1::2]
    return amps / math.sqrt(float(np.sum(np.abs(amps) ** 2)))


def dump(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    with open(path, "w") as f:
        json.dump(obj, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


def shor_states():
    cases = []
    for (Cn, a, L, M, seeds) in [(15, 7, 3, 4, [12345] + list(range(1, 21))),
                                 (21, 2, 5, 5, [2021, 1, 2, 3, 4, 5]),
                                 (15, 2, 4, 4, [7, 8]),
                                 (33, 5, 4, 6, [1, 2]),       # 5^8 = 390625 fits: verbatim == modular
                                 (15, 6, 3, 4, [3]),          # gcd(6,15)=3: non-bijective a^x mod C
                                 (21, 4, 6, 5, [11])]:        # 4^32 overflows INT_POW -> A = 0 gate
        ref = Reference(L, M)
        ref.reset_register()
        ref.quantum_computation(Cn, a)
        state = ref.get_state()
        norm2 = ref.norm2()
        measured = []
        for s in seeds:
            ref.set_state(state)
            ref.seed(s)
            r = ref.rng_uniform()
            ref.seed(s)
            idx = ref.measure_state()
            measured.append({"seed": s, "r": r.hex(), "index": idx, "omega": ref.read_omega(idx)})
        atox = [Reference.int_pow(a, 1 << k) if (1 << k) < 2 ** 32 else None for k in range(L)]
        cases.append({"C": Cn, "a": a, "L": L, "M": M, "state": hexs(state), "norm2": norm2.hex(),
                      "measured": measured, "atox_verbatim": atox})
        ref.close()
    dump("shor_states.json", {"doc": "state after quantum_computation (qc_shor.c:712-737) and "
                                     "measure_state results (qc_shor.c:272-306) from the unmodified reference",
                              "cases": cases})


def single_gates():
    L, M = 3, 3
    n = L + M
    base = synthetic_state(n, 99)
    out = {"L": L, "M": M, "input": hexs(base), "hadamard": [], "cphase": [], "amodc": []}
    for q in range(n):
        ref = Reference(L, M)
        ref.set_state(base)
        ref.hadamard_gate(q)
        out["hadamard"].append({"q": q, "state": hexs(ref.get_state())})
        ref.close()
    for (c, q, th) in [(5, 4, math.pi / 2), (5, 0, math.pi / 32), (1, 3, 0.3), (0, 5, -1.1), (2, 1, math.pi),
                       (4, 2, math.pi / 4)]:
        ref = Reference(L, M)
        ref.set_state(base)
        ref.c_phase_shift_gate(c, q, th)
        out["cphase"].append({"c": c, "q": q, "theta": th.hex(), "state": hexs(ref.get_state())})
        ref.close()
    # (C, atox, control): bijective, non-bijective (gcd != 1), A = 0, C > 2^M, f >= C rows
    for (Cn, atox, c) in [(7, 3, 3), (7, 3, 5), (6, 4, 4), (6, 3, 3), (5, 10, 4), (8, 5, 5), (8, 6, 3),
                          (11, 7, 3), (5, 2 ** 40 + 3, 4), (3, 2, 5)]:
        ref = Reference(L, M)
        ref.set_state(base)
        ref.c_amodc_gate(Cn, atox, c)
        out["amodc"].append({"C": Cn, "atox": atox, "c": c, "state": hexs(ref.get_state())})
        ref.close()
    dump("single_gates.json", out)


def iqft():
    cases = []
    for (L, M, seed) in [(6, 2, 5), (8, 0, 6), (5, 1, 7), (9, 0, 1234), (10, 0, 1234)]:
        n = L + M
        base = synthetic_state(n, seed)
        ref = Reference(L, M)
        ref.set_state(base)
        ref.inverse_QFT()
        cases.append({"L": L, "M": M, "seed": seed, "input": hexs(base), "output": hexs(ref.get_state())})
        ref.close()
    dump("inverse_qft.json", {"doc": "inverse_QFT (qc_shor.c:678-690) on the synthetic state", "cases": cases})


def scalars():
    ref = Reference(1, 1)
    ref.seed(5489)
    mt = {"seed5489_first_uniform": [ref.rng_uniform().hex() for _ in range(4)]}
    ref.seed(0)
    mt["seed0_first_uniform"] = [ref.rng_uniform().hex() for _ in range(2)]
    ref.seed(4357)
    mt["seed4357_first_uniform"] = [ref.rng_uniform().hex() for _ in range(2)]
    ref.close()
    int_pow = [{"base": b, "power": p, "value": Reference.int_pow(b, p)}
               for b in (2, 3, 4, 5, 7, 10, 13) for p in (0, 1, 2, 3, 4, 8, 12, 16, 20, 31, 32, 33, 40, 64, 128, 512)]
    cf = []
    for L in (3, 5, 8):
        for x in range(1, 1 << L):      # omega = 0 divides by zero in the reference (Appendix B #4)
            omega = x / (1 << L)
            cf.append({"omega": omega.hex(), "den": Reference.cf_denominators(omega, 15)})
    gcds = [{"a": a, "b": b, "g": Reference.gcd(a, b)} for a in (0, 1, 6, 15, 16, 21, 344, 2401)
            for b in (0, 1, 4, 15, 21, 35)]
    omegas = []
    for (L, M) in [(3, 4), (5, 5), (4, 1)]:
        ref = Reference(L, M)
        for s in range(0, 1 << (L + M), max(1, (1 << (L + M)) // 37)):
            omegas.append({"L": L, "M": M, "state": s, "omega": ref.read_omega(s).hex()})
        ref.close()
    dump("scalars.json", {"mt19937": mt, "int_pow": int_pow, "continued_fractions": cf, "gcd": gcds,
                          "read_omega": omegas})


def shor_runs():
    runs = []
    for (Cn, L, M, a, seed) in [(15, 3, 4, 7, 12345), (15, 3, 4, 7, 1), (15, 3, 4, 2, 5), (15, 3, 4, 11, 9),
                                (21, 5, 5, 2, 2021), (21, 5, 5, 2, 3), (15, 3, 4, 0, 42), (21, 4, 5, 0, 7)]:
        ref = Reference(L, M)
        ref.seed(seed)
        # shors_algorithm prints progress; silence stdout of the C library
        fd = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        sys.stdout.flush()
        os.dup2(devnull, 1)
        try:
            err, factors = ref.shors_algorithm(Cn, a)
        finally:
            os.dup2(fd, 1)
            os.close(devnull)
            os.close(fd)
        runs.append({"C": Cn, "L": L, "M": M, "a": a, "seed": seed, "error": err, "factors": list(factors)})
        ref.close()
    dump("shor_runs.json", {"doc": "shors_algorithm (qc_shor.c:1003-1134) end to end, gsl_rng_set(seed); "
                                   "only runs whose period search succeeds legitimately are kept "
                                   "(Appendix B #1: period_found is uninitialised in the reference)",
                            "runs": runs})


def shor_stdout():
    """Everything shors_algorithm prints, per verbosity level; the wall-time figure is replaced by <t>.
    Only runs in which a candidate period passes legitimately (Appendix B #1)."""
    import re
    runs = [(15, 3, 4, 7, 12345), (15, 3, 4, 7, 1), (15, 3, 4, 2, 5), (15, 3, 4, 11, 9), (21, 5, 5, 2, 2021),
            (21, 5, 5, 2, 3), (15, 3, 4, 0, 42), (21, 4, 5, 0, 7),
            (15, 3, 4, 14, 3),      # period 2, 14 = C - 1: rejected
            (15, 3, 4, 4, 8),       # period 2
            (21, 5, 5, 8, 6),       # period 2
            (21, 5, 5, 20, 2),      # period 2, 20 = C - 1: rejected
            (21, 5, 5, 4, 10)]      # period 3, odd: rejected (and INT_POW(4, 16) wraps: an A = 0 gate on the way)
    cases = []
    for (Cn, L, M, a, seed) in runs:
        for (v, vv) in ((0, 0), (1, 0), (1, 1)):
            err, factors, out = Reference.shor_stdout(Cn, L, M, a, seed, v, vv)
            out = re.sub(r"Algorithm: [0-9.]+s\.", "Algorithm: <t>s.", out)
            cases.append({"C": Cn, "L": L, "M": M, "a": a, "seed": seed, "verbose": v, "very_verbose": vv,
                          "error": err, "factors": factors, "stdout": out})
    dump("shor_stdout.json", {"doc": "stdout of shors_algorithm (qc_shor.c:1003-1134) of the unmodified reference, "
                                     "gsl_rng_set(seed), flags -v / -V; <t> stands for the wall-time figure",
                              "cases": cases})


def debug_stdout():
    cases = [{"C": Cn, "a": a, "L": L, "M": M, "stdout": Reference.debug_stdout(Cn, a, L, M)}
             for (Cn, a, L, M) in [(15, 7, 3, 4), (21, 2, 5, 5), (15, 6, 3, 4), (21, 4, 6, 5)]]
    dump("debug_stdout.json", {"doc": "stdout of display_state, then check_normalisation (testing_and_debug.c:7-37, unmodified) "
                                      "on the state after reset_register + quantum_computation(C, a)", "cases": cases})


def warnings():
    cases = [{"C": Cn, "L": L, "M": M, "stdout": Reference.warnings_text(Cn, L, M)}
             for (Cn, L, M) in [(15, 3, 4), (15, 8, 4), (15, 8, 3), (21, 3, 3), (21, 10, 5), (21, 9, 5), (33, 5, 5),
                                (35, 11, 6), (16, 8, 4), (17, 8, 4), (255, 16, 8), (257, 16, 8), (4087, 18, 12)]]
    dump("warnings.json", {"doc": "stdout of issue_warnings (qc_shor.c:340-351) of the unmodified reference", "cases": cases})


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (shor_states, single_gates, iqft, scalars, shor_runs, warnings, shor_stdout, debug_stdout):
        if not only or fn.__name__ in only:
            fn()
