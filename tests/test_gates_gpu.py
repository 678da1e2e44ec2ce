"""GPU parity tests, per-gate (reference-order) kernels, through the C ABI.

Checker: the oracle restatement (bit-identical to the reference, see
test_oracle.py) and the committed golden vectors of the unmodified reference.
Bar: amplitudes value-identical (every double compares ==; the sign of an exact
zero is the only freedom), measured indices and factors identical."""
import math

import numpy as np
import pytest

from conftest import load_golden, unhex_c128, values_equal

pytestmark = pytest.mark.gpu


def exact_register(qcs, L, M):
    reg = qcs.Register(L, M)
    reg.set_option(qcs.OPT_FUSION, 0)
    return reg


def test_reset_register(qcs):
    with exact_register(qcs, 3, 4) as reg:
        reg.reset_register()
        s = reg.get_state()
        assert s[1] == 1.0 and np.count_nonzero(s) == 1
        assert reg.norm2() == 1.0


def test_single_gates_match_reference_golden(qcs):
    g = load_golden("single_gates.json")
    base = unhex_c128(g["input"])
    with exact_register(qcs, g["L"], g["M"]) as reg:
        for h in g["hadamard"]:
            reg.set_state(base)
            reg.hadamard_gate(h["q"])
            assert values_equal(reg.get_state(), unhex_c128(h["state"])), ("H", h["q"])
        for c in g["cphase"]:
            reg.set_state(base)
            reg.c_phase_shift_gate(c["c"], c["q"], float.fromhex(c["theta"]))
            assert values_equal(reg.get_state(), unhex_c128(c["state"])), ("CP", c["c"], c["q"])
        for a in g["amodc"]:
            reg.set_state(base)
            reg.c_amodc_gate(a["C"], a["atox"], a["c"])
            assert values_equal(reg.get_state(), unhex_c128(a["state"])), ("amodc", a["C"], a["atox"], a["c"])


def test_inverse_qft_gate_by_gate_matches_reference_golden(qcs):
    for case in load_golden("inverse_qft.json")["cases"]:
        with exact_register(qcs, case["L"], case["M"]) as reg:
            reg.set_state(unhex_c128(case["input"]))
            reg.inverse_QFT()
            assert values_equal(reg.get_state(), unhex_c128(case["output"])), (case["L"], case["M"])


def test_shor_states_and_measurement_match_reference_golden(qcs):
    for case in load_golden("shor_states.json")["cases"]:
        with exact_register(qcs, case["L"], case["M"]) as reg:
            reg.reset_register()
            reg.quantum_computation(case["C"], case["a"], qcs.POW_VERBATIM)
            want = unhex_c128(case["state"])
            assert values_equal(reg.get_state(), want), case["C"]
            assert abs(reg.norm2() - float.fromhex(case["norm2"])) < 1e-14
            for m in case["measured"]:
                reg.set_state(want)
                idx = reg.measure_state(float.fromhex(m["r"]))
                assert idx == m["index"], (case["C"], m)
                collapsed = reg.get_state()
                assert collapsed[idx] == 1.0 and np.count_nonzero(collapsed) == 1


@pytest.mark.parametrize("L,M", [(1, 1), (2, 3), (7, 4), (9, 5), (10, 7), (13, 3)])
def test_random_circuits_match_oracle(qcs, oracle_built, L, M):
    """Random H / C-phase / a^x mod C sequences at sizes up to n = 17."""
    n = L + M
    rng = np.random.default_rng(1000 * L + M)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    v /= np.linalg.norm(v)
    o = oracle_built.Restatement(L, M)
    o.set_state(v)
    with exact_register(qcs, L, M) as reg:
        reg.set_state(v)
        for step in range(14):
            kind = int(rng.integers(0, 3))
            if kind == 0:
                q = int(rng.integers(0, n))
                o.hadamard_gate(q)
                reg.hadamard_gate(q)
            elif kind == 1:
                c, q = (int(x) for x in rng.choice(n, size=2, replace=False))
                th = float(rng.uniform(-math.pi, math.pi))
                o.c_phase_shift_gate(c, q, th)
                reg.c_phase_shift_gate(c, q, th)
            else:
                Cn = int(rng.integers(2, (1 << M) + 1))
                atox = int(rng.integers(0, 5000))
                c = int(rng.integers(M, n))
                o.c_amodc_gate(Cn, atox, c)
                reg.c_amodc_gate(Cn, atox, c)
        assert values_equal(reg.get_state(), o.get_state())
        # ragged measurement cases: first index, last index, fall-through, interior
        state = o.get_state().copy()
        for r in (0.0, 1e-300, 0.25, 0.5, 0.999999, 1.0 - 2 ** -32, 2.0):
            o.set_state(state)
            reg.set_state(state)
            assert reg.measure_state(r) == o.measure_state(r), r


def test_amodc_edge_cases_match_oracle(qcs, oracle_built):
    """control inside the M register, C > 2^M, C = 1, A = 0, atox >= 2^32."""
    L, M = 4, 4
    n = L + M
    rng = np.random.default_rng(77)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    with exact_register(qcs, L, M) as reg:
        for (Cn, atox, c) in [(13, 5, 2), (13, 5, 0), (23, 7, 5), (40, 3, 6), (1, 9, 4), (12, 24, 7),
                              (16, 3, 4), (9, 2 ** 50 + 1, 5), (15, 10, 6)]:
            o = oracle_built.Restatement(L, M)
            o.set_state(v)
            o.c_amodc_gate(Cn, atox, c)
            reg.set_state(v)
            reg.c_amodc_gate(Cn, atox, c)
            assert values_equal(reg.get_state(), o.get_state()), (Cn, atox, c)


def test_bad_arguments(qcs):
    with exact_register(qcs, 3, 4) as reg:
        for call in (lambda: reg.hadamard_gate(7), lambda: reg.c_phase_shift_gate(7, 0, 0.1),
                     lambda: reg.c_amodc_gate(0, 3, 4), lambda: reg.c_amodc_gate(15, 3, 9)):
            with pytest.raises(qcs.QcsError) as e:
                call()
            assert e.value.code == qcs.BAD_ARGUMENTS


def test_nonzero_states_is_display_state(qcs):
    case = load_golden("shor_states.json")["cases"][0]
    with exact_register(qcs, case["L"], case["M"]) as reg:
        reg.set_state(unhex_c128(case["state"]))
        idx, mag, count = reg.nonzero_states()
        assert count == 16 and idx == [1, 4, 7, 13, 17, 20, 23, 29, 33, 36, 39, 45, 49, 52, 55, 61]
        assert all(abs(m - 0.25) < 1e-15 for m in mag)


def test_hh_identity_and_norm_at_scale(qcs):
    """Size-independent properties at n = 26 (1 GiB): H.H = I to rounding, norm kept."""
    n = 26
    with exact_register(qcs, n, 0) as reg:
        reg.fill_synthetic(1234)
        s0 = reg.norm2()
        reg.scale(1.0 / math.sqrt(s0))
        assert abs(reg.norm2() - 1.0) < 1e-12
        probe = [0, 1, 12345, (1 << n) - 1]
        before = [reg.get_state(i, 1)[0] for i in probe]
        for q in (0, 3, 5, 17, n - 1):
            reg.hadamard_gate(q)
            assert abs(reg.norm2() - 1.0) < 1e-12
            reg.hadamard_gate(q)
        after = [reg.get_state(i, 1)[0] for i in probe]
        assert max(abs(a - b) for a, b in zip(before, after)) < 1e-15


@pytest.mark.parametrize("n,kind", [(18, "random"), (20, "random"), (22, "random"), (21, "sparse"),
                                    (20, "ties"), (19, "spiky"), (21, "plateau")])
def test_parallel_measurement_reproduces_sequential_rounding(qcs, oracle_built, n, kind):
    """The parallel scan (csrc/measure.cu) must return the index of the reference's
    sequential loop (qc_shor.c:283-292) -- checked against the oracle and against the
    single-CTA sequential GPU scan, including r values placed exactly on and next to
    prefix sums."""
    rng = np.random.default_rng(n)
    N = 1 << n
    if kind == "random":
        v = rng.normal(size=N) + 1j * rng.normal(size=N)
    elif kind == "sparse":
        v = np.zeros(N, dtype=np.complex128)
        idx = rng.choice(N, size=64, replace=False)
        v[idx] = rng.normal(size=64) + 1j * rng.normal(size=64)
    elif kind == "ties":
        # equal power-of-two probabilities: every addition is exact or an exact tie
        v = np.zeros(N, dtype=np.complex128)
        v[:: N // 4096] = 1.0
        v[1:: N // 2048] = 0.5
    elif kind == "plateau":
        # four peaks of probability exactly 1/4 over addends around and far below half an ulp of the
        # running sum: the sum sits on the binade boundaries 1/4 and 1/2 for hundreds of chunks
        v = rng.choice([1e-15, 3e-9, 5.2e-9, 7.4e-9, 1.05e-8], size=N, p=[0.9, 0.04, 0.03, 0.02, 0.01]).astype(np.complex128)
        v[N // 8::N // 4] = 0.5
    else:
        v = (rng.normal(size=N) + 1j * rng.normal(size=N)) * np.exp(rng.normal(size=N) * 6)
    if kind != "plateau":
        v /= np.linalg.norm(v)
    o = oracle_built.Restatement(n, 0)
    p = np.abs(v.real) ** 2 + np.abs(v.imag) ** 2
    prefix = np.cumsum(p)
    rs = [0.0, 1e-12, 0.1, 0.25, 0.5, 0.75, 0.9999, 1.0, 1.5]
    rs += [float(rng.uniform()) for _ in range(6)]
    for k in rng.integers(1, N - 1, size=4):
        rs += [float(prefix[k]), float(np.nextafter(prefix[k], 0)), float(np.nextafter(prefix[k], 2))]
    with qcs.Register(n, 0) as reg, qcs.Register(n, 0) as seq:
        seq.set_option(qcs.OPT_MEASURE_SEQUENTIAL, 1)
        for r in rs:
            o.set_state(v)
            reg.set_state(v)
            seq.set_state(v)
            want = o.measure_state(r)
            assert seq.measure_state(r) == want, (kind, r)
            assert reg.measure_state(r) == want, (kind, r)


def test_parallel_measurement_after_qft_at_scale(qcs):
    """n = 26: parallel-exact scan == single-CTA sequential scan on a QFT output state."""
    n = 26
    with qcs.Register(n, 0) as reg:
        reg.fill_synthetic(7)
        reg.scale(1.0 / math.sqrt(reg.norm2()))
        reg.inverse_QFT()
        state = reg.get_state().copy()
        for r in (0.001, 0.37, 0.93):
            reg.set_state(state)
            reg.set_option(qcs.OPT_MEASURE_SEQUENTIAL, 0)
            a = reg.measure_state(r)
            reg.set_state(state)
            reg.set_option(qcs.OPT_MEASURE_SEQUENTIAL, 1)
            b = reg.measure_state(r)
            assert a == b, r


def test_parallel_measurement_on_shor_state(qcs, oracle_built):
    """The state find_period actually measures (peaks of very different height over exact zeros): many
    chunks cross a binade boundary and take the refined (sub-chunk) path of the walk.  Same index as the
    oracle's loop and as the single-CTA sequential scan, for variates across the whole range."""
    L, M, C, a = 13, 9, 511, 5
    n = L + M
    rng = np.random.default_rng(5)
    with qcs.Register(L, M) as reg, qcs.Register(L, M) as seq:
        seq.set_option(qcs.OPT_MEASURE_SEQUENTIAL, 1)
        reg.reset_register()
        reg.quantum_computation(C, a, qcs.POW_VERBATIM)
        v = reg.get_state().copy()
        p = np.abs(v.real) ** 2 + np.abs(v.imag) ** 2
        prefix = np.cumsum(p)
        rs = [1e-9, 0.01, 0.125, 0.25, 0.5, 0.9, 0.999999] + [float(rng.uniform()) for _ in range(8)]
        for k in rng.integers(1, (1 << n) - 1, size=3):
            rs += [float(prefix[k]), float(np.nextafter(prefix[k], 2))]
        o = oracle_built.Restatement(L, M)
        for r in rs:
            o.set_state(v)
            reg.set_state(v)
            seq.set_state(v)
            want = o.measure_state(r)
            assert seq.measure_state(r) == want, r
            assert reg.measure_state(r) == want, r
