"""GPU parity tests of the fused tile sweeps (QCS_OPT_FUSION = 1) through the C ABI.

Bar (north_star): amplitude L2 error <= 1e-12 relative to the reference's
double-precision state; measured indices and factors identical."""
import math

import numpy as np
import pytest

from conftest import load_golden, rel_l2, unhex_c128

pytestmark = pytest.mark.gpu
TOL = 1e-12


def bitrev(k, bits):
    out = 0
    for b in range(bits):
        if k >> b & 1:
            out |= 1 << (bits - 1 - b)
    return out


def test_inverse_qft_matches_reference_golden(qcs):
    for case in load_golden("inverse_qft.json")["cases"]:
        for tile_bits in (0, 8, 9, 11, 13):
            with qcs.Register(case["L"], case["M"]) as reg:
                reg.set_option(qcs.OPT_TILE_BITS, tile_bits)
                reg.set_state(unhex_c128(case["input"]))
                reg.inverse_QFT()
                err = rel_l2(reg.get_state(), unhex_c128(case["output"]))
                assert err <= TOL, (case["L"], case["M"], tile_bits, err)


@pytest.mark.parametrize("L,M,tile_bits", [(14, 0, 0), (14, 0, 8), (13, 3, 0), (16, 2, 10), (18, 0, 0),
                                           (17, 1, 13), (20, 0, 11), (20, 0, 13), (12, 5, 9)])
def test_inverse_qft_matches_oracle_with_strided_sweeps(qcs, oracle_built, L, M, tile_bits):
    _check_iqft_vs_oracle(qcs, oracle_built, L, M, tile_bits, pipeline=1)


@pytest.mark.parametrize("L,M", [(14, 0), (15, 2), (19, 0), (16, 5)])
def test_inverse_qft_direct_kernel_matches_oracle(qcs, oracle_built, L, M):
    """QCS_OPT_PIPELINE = 0: the global<->register sweep kernel at the default tile size."""
    _check_iqft_vs_oracle(qcs, oracle_built, L, M, 0, pipeline=0)


def _check_iqft_vs_oracle(qcs, oracle_built, L, M, tile_bits, pipeline):
    n = L + M
    o = oracle_built.Restatement(L, M)
    o.fill_synthetic(1234 + n)
    o.scale(1.0 / math.sqrt(o.norm2()))
    base = o.get_state().copy()
    o.inverse_QFT()
    with qcs.Register(L, M) as reg:
        reg.set_option(qcs.OPT_TILE_BITS, tile_bits)
        reg.set_option(qcs.OPT_PIPELINE, pipeline)
        reg.set_state(base)
        reg.inverse_QFT()
        err = rel_l2(reg.get_state(), o.get_state())
        assert err <= TOL, err


@pytest.mark.parametrize("L,M,tile_bits", [(5, 2, 0), (10, 0, 8), (14, 1, 0), (16, 0, 9), (18, 0, 13)])
def test_forward_qft_is_the_adjoint_circuit(qcs, L, M, tile_bits):
    """fused forward QFT == gate-by-gate adjoint circuit (reverse order, -theta)."""
    n = L + M
    rng = np.random.default_rng(n)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    v /= np.linalg.norm(v)
    with qcs.Register(L, M) as a, qcs.Register(L, M) as b:
        a.set_option(qcs.OPT_FUSION, 0)
        b.set_option(qcs.OPT_TILE_BITS, tile_bits)
        a.set_state(v)
        b.set_state(v)
        a.QFT()
        b.QFT()
        assert rel_l2(b.get_state(), a.get_state()) <= TOL
        b.inverse_QFT()
        assert rel_l2(b.get_state(), v) <= TOL


def test_ranged_qft(qcs):
    n = 15
    rng = np.random.default_rng(5)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    v /= np.linalg.norm(v)
    with qcs.Register(n, 0) as a, qcs.Register(n, 0) as b:
        a.set_option(qcs.OPT_FUSION, 0)
        for (lo, hi) in [(3, 11), (0, 15), (13, 15), (6, 7)]:
            a.set_state(v)
            b.set_state(v)
            a.inverse_QFT(lo, hi)
            b.inverse_QFT(lo, hi)
            assert rel_l2(b.get_state(), a.get_state()) <= TOL, (lo, hi)


def test_shor_states_fused_match_reference_golden(qcs):
    for case in load_golden("shor_states.json")["cases"]:
        with qcs.Register(case["L"], case["M"]) as reg:
            reg.reset_register()
            reg.quantum_computation(case["C"], case["a"], qcs.POW_VERBATIM)
            want = unhex_c128(case["state"])
            got = reg.get_state()
            assert rel_l2(got, want) <= TOL, case["C"]
            for m in case["measured"]:
                reg.set_state(got)
                assert reg.measure_state(float.fromhex(m["r"])) == m["index"], (case["C"], m)


def test_shor_full_size_n14_intended_power_mode(qcs, oracle_built):
    """BASELINE configs[1] at full size: C=21, a=2, L=9, M=5 (n=14), modular powers."""
    L, M, Cn, a = 9, 5, 21, 2
    o = oracle_built.Restatement(L, M)
    o.reset_register()
    o.quantum_computation(Cn, a, 1)
    want = o.get_state().copy()
    with qcs.Register(L, M) as reg:
        reg.reset_register()
        reg.quantum_computation(Cn, a, qcs.POW_MODULAR)
        got = reg.get_state()
        assert rel_l2(got, want) <= TOL
        g = o.rng(2021)
        for _ in range(20):
            r = g.uniform()
            o.set_state(want)
            reg.set_state(got)
            assert reg.measure_state(r) == o.measure_state(r)


def test_qft_closed_form_and_round_trip_at_scale(qcs):
    """n = 27 (2 GiB): inverse_QFT |k> = e^{2 pi i jk/N}/sqrt(N) at bit-reversed j; QFT undoes it."""
    n = 27
    N = 1 << n
    k = 0x2C0FFEE
    with qcs.Register(n, 0) as reg:
        reg.reset_register()                       # |0...01>
        reg.set_state(np.array([0j, 0j]), first=0)
        reg.set_state(np.array([1 + 0j]), first=k)
        reg.inverse_QFT()
        assert abs(reg.norm2() - 1.0) < 1e-12
        for j in (0, 1, 2, 12345, 0x155555, N - 1, N // 2 + 77):
            got = reg.get_state(bitrev(j, n), 1)[0]
            want = np.exp(2j * math.pi * ((j * k) % N) / N) / math.sqrt(N)
            assert abs(got - want) <= 1e-12 * abs(want), j
        reg.QFT()
        assert abs(reg.get_state(k, 1)[0] - 1.0) < 1e-12
        assert abs(reg.norm2() - 1.0) < 1e-12


def test_fused_sweep_count(qcs):
    """n = 26: the whole 351-gate inverse QFT is a handful of kernel launches."""
    with qcs.Register(26, 0) as reg:
        reg.fill_synthetic(1)
        reg.profile_reset()
        reg.inverse_QFT()
        reg.synchronize()
        launches, _, by = reg.profile()["tile_sweep"]
        # algorithmic bytes = sweeps made x 32 B per amplitude; an L2-paired launch makes two sweeps
        sweeps = by / (32.0 * (1 << 26))
        assert 1 <= launches <= 5 and sweeps == int(sweeps) and launches <= sweeps <= 2 * launches


@pytest.mark.parametrize("L,M,Cn,a,mode", [(3, 4, 15, 7, 0), (5, 5, 21, 2, 0), (6, 5, 21, 4, 0), (3, 4, 15, 6, 0),
                                           (9, 5, 21, 2, 1), (8, 7, 77, 3, 1), (6, 11, 2047, 5, 1), (12, 6, 35, 4, 1),
                                           (7, 4, 12, 10, 1), (5, 12, 4001, 7, 1)])
def test_quantum_computation_on_arbitrary_input_state(qcs, L, M, Cn, a, mode):
    """Fused H^L + modexp sweep + fused inverse QFT vs the gate-by-gate kernels, from a random
    (not reset) state; includes non-bijective multipliers and blocks up to 2^12."""
    n = L + M
    rng = np.random.default_rng(n * 100 + Cn)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    v /= np.linalg.norm(v)
    with qcs.Register(L, M) as exact, qcs.Register(L, M) as fused:
        exact.set_option(qcs.OPT_FUSION, 0)
        exact.set_state(v)
        fused.set_state(v)
        exact.quantum_computation(Cn, a, mode)
        fused.quantum_computation(Cn, a, mode)
        assert rel_l2(fused.get_state(), exact.get_state()) <= TOL
        prof = fused.profile()
        assert prof["modexp_sweep"][0] == 1 and prof["amodc"][0] == 0


@pytest.mark.parametrize("k,n", [(3, 6), (4, 7), (3, 14), (4, 16), (4, 22)])
def test_dense_block_dmma(qcs, oracle_built, k, n):
    """qcs_apply_dense_block (FP64 tensor cores) == the gates it fuses: the k-qubit inverse
    QFT on qubits 0..k-1 multiplied out into one 2^k x 2^k matrix, and a random unitary."""
    R = 1 << k
    # matrix of inverse_QFT on a k-qubit register, column by column, from the oracle
    U = np.zeros((R, R), dtype=np.complex128)
    for j in range(R):
        o = oracle_built.Restatement(k, 0)
        e = np.zeros(R, dtype=np.complex128)
        e[j] = 1.0
        o.set_state(e)
        o.inverse_QFT()
        U[:, j] = o.get_state()
    rng = np.random.default_rng(k * 100 + n)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    v /= np.linalg.norm(v)
    with qcs.Register(n, 0) as dense, qcs.Register(n, 0) as gates:
        gates.set_option(qcs.OPT_FUSION, 0)
        dense.set_state(v)
        gates.set_state(v)
        dense.apply_dense_block(k, U)
        gates.inverse_QFT(0, k)
        assert rel_l2(dense.get_state(), gates.get_state()) <= TOL
        assert dense.profile()["dense_block"][0] == 1
        # random unitary vs numpy
        q_, _ = np.linalg.qr(rng.normal(size=(R, R)) + 1j * rng.normal(size=(R, R)))
        dense.set_state(v)
        dense.apply_dense_block(k, q_)
        want = (v.reshape(-1, R) @ q_.T).reshape(-1)
        assert rel_l2(dense.get_state(), want) <= TOL


@pytest.mark.parametrize("shape", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("L,M,run_bits", [(18, 0, 3), (19, 2, 4), (22, 0, 3)])
def test_every_pipeline_shape_matches_oracle(qcs, oracle_built, shape, L, M, run_bits):
    """QCS_OPT_PIPE_SHAPE: each instantiated shape of the TMA pipeline (tile size, ring depth,
    consumer groups) against the oracle, inverse and forward."""
    n = L + M
    o = oracle_built.Restatement(L, M)
    o.fill_synthetic(900 + n)
    o.scale(1.0 / math.sqrt(o.norm2()))
    base = o.get_state().copy()
    o.inverse_QFT()
    with qcs.Register(L, M) as reg:
        reg.set_option(qcs.OPT_PIPE_SHAPE, shape)
        reg.set_option(qcs.OPT_MIN_RUN_BITS, run_bits)
        reg.set_state(base)
        reg.inverse_QFT()
        assert rel_l2(reg.get_state(), o.get_state()) <= TOL
        reg.QFT()
        assert rel_l2(reg.get_state(), base) <= TOL


def test_full_size_n30_properties(qcs):
    """BASELINE configs[2] at its full size (n = 30, 16 GiB), through size-independent properties:
    closed form of inverse_QFT on a basis state, unitarity (norm), QFT undoing inverse_QFT on the
    synthetic state (probed amplitudes), linearity in a global phase."""
    n = 30
    N = 1 << n
    k = 0x1C0FFEE1 % N
    probes = [0, 1, 4097, 123456789 % N, N - 1, N // 2 + 77, 0x2AAAAAAA % N]
    with qcs.Register(n, 0) as reg:
        reg.reset_register()
        reg.set_state(np.array([0j, 0j]), first=0)
        reg.set_state(np.array([1 + 0j]), first=k)
        reg.inverse_QFT()
        assert abs(reg.norm2() - 1.0) < 1e-12
        for j in probes:
            got = reg.get_state(bitrev(j, n), 1)[0]
            want = np.exp(2j * math.pi * ((j * k) % N) / N) / math.sqrt(N)
            assert abs(got - want) <= 1e-12 * abs(want), j
        reg.fill_synthetic(1234)
        reg.scale(1.0 / math.sqrt(reg.norm2()))
        before = np.array([reg.get_state(i, 1)[0] for i in probes])
        reg.inverse_QFT()
        assert abs(reg.norm2() - 1.0) < 1e-12
        mid = np.array([reg.get_state(i, 1)[0] for i in probes])
        reg.QFT()
        after = np.array([reg.get_state(i, 1)[0] for i in probes])
        assert np.max(np.abs(after - before)) <= 1e-12 * np.max(np.abs(before)) * 10
        # linearity: i * state -> i * transformed state
        reg.apply_gate(0, np.array([[1j, 0], [0, 1j]]))
        reg.inverse_QFT()
        mid2 = np.array([reg.get_state(i, 1)[0] for i in probes])
        assert np.max(np.abs(mid2 - 1j * mid)) <= 1e-12 * np.max(np.abs(mid)) * 10


@pytest.mark.parametrize("L,M,Cn,a,mode", [(3, 4, 15, 7, 0), (5, 5, 21, 2, 0), (6, 5, 21, 4, 0), (3, 4, 15, 6, 0), (7, 4, 15, 7, 0),
                                           (9, 5, 21, 2, 1), (8, 7, 77, 3, 1), (6, 11, 2047, 5, 1), (13, 6, 35, 4, 1),
                                           (7, 4, 12, 10, 1), (5, 12, 4001, 7, 1), (6, 4, 21, 2, 1), (9, 1, 2, 3, 1), (16, 8, 221, 6, 1)])
def test_quantum_computation_from_reset_closed_form(qcs, L, M, Cn, a, mode):
    """reset_register + quantum_computation in fused mode writes the state after the Hadamards and the
    controlled multiplications in closed form (one pass); it must equal the gate-by-gate kernels -- non-bijective
    multipliers, A = 0 from the INT_POW overflow (L = 7, a = 7 verbatim), C > 2^M (row index masked), odd L."""
    with qcs.Register(L, M) as fused, qcs.Register(L, M) as exact:
        exact.set_option(qcs.OPT_FUSION, 0)
        for reg in (fused, exact):
            reg.reset_register()
        before = fused.launch_count
        fused.quantum_computation(Cn, a, mode)
        launches = fused.launch_count - before
        exact.quantum_computation(Cn, a, mode)
        want, got = exact.get_state(), fused.get_state()
        assert rel_l2(got, want) <= TOL
        assert launches <= 1 + 4, launches            # one closed-form pass + the inverse-QFT sweeps
        # the deferred reset is observable as a reset when nothing consumes it
        fused.reset_register()
        assert fused.nonzero_states()[0] == [1]
        # ... and a second computation from a NON-reset state takes the general path
        fused.hadamard_gate(M)
        exact.reset_register()
        exact.hadamard_gate(M)
        fused.quantum_computation(Cn, a, mode)
        exact.quantum_computation(Cn, a, mode)
        assert rel_l2(fused.get_state(), exact.get_state()) <= TOL


@pytest.mark.parametrize("L,M,Cn,a,mode", [(12, 6, 35, 4, 1), (13, 6, 35, 4, 1), (16, 8, 221, 6, 1), (10, 12, 4001, 7, 1), (14, 9, 511, 5, 0),
                                           (15, 5, 21, 2, 1), (11, 9, 300, 10, 1), (18, 4, 15, 7, 1), (12, 12, 5000, 3, 1)])
def test_quantum_computation_from_reset_generating_sweep(qcs, L, M, Cn, a, mode):
    """Registers large enough for the pipelined sweeps: after reset_register the state |x, f(x)> is not written at
    all -- the first sweep of the inverse QFT builds its tiles from the table of f(x) and skips the empty ones.
    Same state as with the closed-form write pass (QCS_OPT_GEN_SWEEP = 0) and as the gate-by-gate kernels, for
    bijective and non-bijective multipliers, C > 2^M and INT_POW overflow; where the launch plan allows the
    generating sweep, no pass over the whole state is accounted to the fill class."""
    N = 1 << (L + M)
    with qcs.Register(L, M) as gen, qcs.Register(L, M) as written, qcs.Register(L, M) as exact:
        written.set_option(qcs.OPT_GEN_SWEEP, 0)
        exact.set_option(qcs.OPT_FUSION, 0)
        for reg in (gen, written, exact):
            reg.set_option(qcs.OPT_PROFILE, 1)
            reg.profile_reset()
            reg.reset_register()
            reg.quantum_computation(Cn, a, mode)
        want = exact.get_state()
        assert rel_l2(gen.get_state(), want) <= TOL
        assert rel_l2(written.get_state(), want) <= TOL
        assert written.profile()["fill"][2] >= 16.0 * N
        generated = gen.profile()["fill"][2] < 16.0 * N
        print(f"L={L} M={M}: generating sweep {'used' if generated else 'not used'}")
        # twice in a row (the table is rebuilt, the pending reset consumed again), then the measured index
        gen.reset_register()
        gen.quantum_computation(Cn, a, mode)
        assert rel_l2(gen.get_state(), want) <= TOL
        exact.set_option(qcs.OPT_FUSION, 1)
        assert gen.measure_state(0.37) == written.measure_state(0.37) == exact.measure_state(0.37)


def test_generating_sweep_is_used_at_scale(qcs):
    """n = 28 (L = 16, M = 12): the plan starts with an unpaired strided sweep, so the generating sweep must
    engage -- and the state must equal the one built by the write pass (probes + norm + measured indices)."""
    L, M, Cn, a = 16, 12, 4087, 7
    N = 1 << (L + M)
    rng = np.random.default_rng(3)
    probes = [int(i) for i in rng.integers(0, N, size=48)]
    with qcs.Register(L, M) as gen, qcs.Register(L, M) as written:
        written.set_option(qcs.OPT_GEN_SWEEP, 0)
        for reg in (gen, written):
            reg.set_option(qcs.OPT_PROFILE, 1)
            reg.profile_reset()
            reg.reset_register()
            reg.quantum_computation(Cn, a, qcs.POW_MODULAR)
        assert written.profile()["fill"][2] >= 16.0 * N
        assert gen.profile()["fill"][2] < 16.0 * N, gen.profile()
        assert abs(gen.norm2() - 1.0) < 1e-12
        got = np.array([gen.get_state(i, 1)[0] for i in probes])
        want = np.array([written.get_state(i, 1)[0] for i in probes])
        assert np.max(np.abs(got - want)) <= 1e-15
        # the heaviest amplitudes agree too: measured indices for several variates
        for r in (0.05, 0.37, 0.62, 0.93):
            for reg in (gen, written):
                reg.reset_register()
                reg.quantum_computation(Cn, a, qcs.POW_MODULAR)
            assert gen.measure_state(r) == written.measure_state(r)
