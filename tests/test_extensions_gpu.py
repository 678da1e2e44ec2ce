"""GPU tests of the rows SURVEY 8(f) ranks as "next": sampling without collapse, state
dump / load, arbitrary (controlled) single-qubit gates.  Where the reference has a
counterpart (Hadamard, controlled phase, measure_state) the check is against the oracle;
otherwise against a dense numpy application of the same 2x2 matrix."""
import math
import os

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def random_state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


def numpy_gate(v, n, q, U, c=None):
    """dense application: amp'[.., b_q, ..] = sum_k U[b_q][k] amp[.., k, ..], where bit c is 1"""
    t = v.reshape([2] * n)                        # axis 0 = qubit n-1 ... axis n-1 = qubit 0
    ax = n - 1 - q
    out = np.moveaxis(np.tensordot(U, np.moveaxis(t, ax, 0), axes=([1], [0])), 0, ax)
    if c is not None:
        sel = [slice(None)] * n
        sel[n - 1 - c] = 0
        out[tuple(sel)] = t[tuple(sel)]
    return out.reshape(-1)


def random_unitary(rng):
    a = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    qm, r = np.linalg.qr(a)
    return qm * (np.diag(r) / np.abs(np.diag(r)))


@pytest.mark.parametrize("n", [3, 9, 14, 20])
def test_arbitrary_gate_matches_numpy(qcs, n):
    rng = np.random.default_rng(n)
    v = random_state(n, 100 + n)
    with qcs.Register(n, 0) as reg:
        reg.set_state(v)
        want = v
        for _ in range(6):
            U = random_unitary(rng)
            q = int(rng.integers(n))
            if rng.random() < 0.5:
                reg.apply_gate(q, U)
                want = numpy_gate(want, n, q, U)
            else:
                c = int(rng.integers(n - 1))
                c = c + 1 if c >= q else c
                reg.apply_controlled_gate(c, q, U)
                want = numpy_gate(want, n, q, U, c)
        assert rel_l2(reg.get_state(), want) <= TOL
        assert abs(reg.norm2() - 1.0) < 1e-12


def test_gate_special_cases_agree_with_reference_gates(qcs, oracle_built):
    """U = H and U = diag(1, e^{i theta}) controlled reproduce hadamard_gate / c_phase_shift_gate
    (qc_shor.c:442-565) to rounding; a non-unitary U is applied as given."""
    n = 12
    o = oracle_built.Restatement(n, 0)
    o.fill_synthetic(3)
    o.scale(1.0 / math.sqrt(o.norm2()))
    base = o.get_state().copy()
    o.hadamard_gate(5)
    o.c_phase_shift_gate(7, 2, 0.7)
    o.hadamard_gate(0)
    H = np.array([[1, 1], [1, -1]]) / math.sqrt(2.0)
    P = np.diag([1.0, np.exp(0.7j)])
    with qcs.Register(n, 0) as reg:
        reg.set_state(base)
        reg.apply_gate(5, H)
        reg.apply_controlled_gate(7, 2, P)
        reg.apply_gate(0, H)
        assert rel_l2(reg.get_state(), o.get_state()) <= TOL
        with pytest.raises(qcs.QcsError):
            reg.apply_gate(n, H)
        with pytest.raises(qcs.QcsError):
            reg.apply_controlled_gate(3, 3, H)
        reg.set_state(base)
        reg.apply_gate(4, np.array([[2.0, 0.0], [0.0, 0.0]]))
        assert abs(reg.norm2() - 4.0 * np.sum(np.abs(base.reshape(-1, 2, 16)[:, 0, :]) ** 2)) < 1e-12


@pytest.mark.parametrize("L,M", [(3, 4), (10, 0), (18, 0), (20, 1), (23, 0), (22, 3)])
def test_sampling_without_collapse_matches_measure_state(qcs, oracle_built, L, M):
    """Every sampled index equals what measure_state returns for the same r (oracle for small
    registers, the engine's own measure_state -- pinned to the oracle elsewhere -- for all)."""
    n = L + M
    rs = [0.0, 1e-9, 0.1234, 0.5, 0.77, 0.999999, 1.0, 1.5]
    with qcs.Register(L, M) as reg:
        reg.fill_synthetic(40 + n)
        reg.scale(1.0 / math.sqrt(reg.norm2()))
        state = reg.get_state().copy()
        if n >= 22:
            # registers of >= 2^21 amplitudes take the single-pass path (boundary sums once, then one
            # 2^20-amplitude scan per variate): probe variates at, just below and just above exact
            # sequential prefix sums, including the ones at the 2^20 boundaries
            p = state.real * state.real + state.imag * state.imag
            cum = np.cumsum(p)                     # close to (not identical with) the sequential sums
            for i in ((1 << 20) - 1, 1 << 20, (1 << 21) + 5, (1 << n) - 3):
                rs += [float(cum[i]), float(np.nextafter(cum[i], 0.0)), float(np.nextafter(cum[i], 2.0))]
            rs += [float(x) for x in np.random.default_rng(n).random(20)]
        got = reg.sample_states(rs)
        assert np.array_equal(reg.get_state().view(np.float64), state.view(np.float64)), "sampling must not collapse"
        want = []
        for r in rs:
            reg.set_state(state)
            want.append(reg.measure_state(r))
        assert got == want
        if n <= 20:
            o = oracle_built.Restatement(L, M)
            ow = []
            for r in rs:
                o.set_state(state)
                ow.append(o.measure_state(r))
            assert got == ow


def test_save_and_load_state(qcs, tmp_path):
    n = 23                                          # 128 MiB: two pieces
    path = os.path.join(tmp_path, "state.qcs")
    with qcs.Register(n - 2, 2) as reg:
        reg.fill_synthetic(5)
        reg.scale(1.0 / math.sqrt(reg.norm2()))
        reg.inverse_QFT()
        want = reg.get_state().copy()
        reg.save_state(path)
        assert os.path.getsize(path) == 64 + 16 * (1 << n)
        # the payload is the interleaved (re, im) array of gsl_vector_complex.data
        raw = np.fromfile(path, dtype=np.float64, offset=64)
        assert np.array_equal(raw, want.view(np.float64))
        reg.reset_register()
        reg.load_state(path)
        assert np.array_equal(reg.get_state().view(np.float64), want.view(np.float64))
    with qcs.Register(n - 1, 1) as other:           # a different register shape refuses the file
        with pytest.raises(qcs.QcsError) as e:
            other.load_state(path)
        assert e.value.code == qcs.BAD_ARGUMENTS
    with qcs.Register(4, 0) as small:
        with pytest.raises(qcs.QcsError):
            small.load_state(os.path.join(tmp_path, "missing.qcs"))


def test_single_caller_handle_on_one_gpu_and_async_upload(qcs, oracle_built):
    """qcs_register_create_multi with n_gpus = 1 is an ordinary register; qcs_set_state_async from a
    pinned buffer allocated next to the device (qcs_host_alloc_near) uploads the same state as qcs_set_state."""
    n = 14
    o = oracle_built.Restatement(n, 0)
    o.fill_synthetic(5)
    o.scale(1.0 / math.sqrt(o.norm2()))
    state = o.get_state().copy()
    o.inverse_QFT()
    want = o.get_state()
    with qcs.Register(n, 0, n_gpus=1) as reg:
        assert reg.num_gpus == 1 and reg.local_states == 1 << n
        pinned = qcs.PinnedBuffer(2 << n, device=0)
        pinned.array[:] = state.view(np.float64)
        reg.set_state_async(pinned.array)
        reg.inverse_QFT()
        got = reg.get_state()
        pinned.close()
        assert rel_l2(got, want) <= 1e-12
    with pytest.raises(qcs.QcsError):
        qcs.Register(n, 0, n_gpus=3)                   # not a power of two


def test_measurement_with_zero_variate_and_leading_zeros(qcs, oracle_built):
    """gsl_rng_uniform can return exactly 0.0: measure_state then returns index 0 whatever amp[0] is
    (qc_shor.c:286-289), also when the parallel scan would skip the all-zero leading chunks."""
    n = 19
    amps = np.zeros(1 << n, dtype=np.complex128)
    amps[(1 << 18) + 5] = 0.6
    amps[(1 << 18) + 77] = 0.8j
    o = oracle_built.Restatement(n, 0)
    o.set_state(amps)
    with qcs.Register(n, 0) as reg:
        reg.set_state(amps)
        for r in (0.0, -0.5):
            assert reg.sample_states([r])[0] == 0
        got = [int(x) for x in reg.sample_states([0.0, 0.2, 0.36, 0.37, 0.999])]
        ref = []
        for r in (0.0, 0.2, 0.36, 0.37, 0.999):
            o.set_state(amps)
            ref.append(int(o.measure_state(r)))
        assert got == ref and got[0] == 0


def test_run_of_general_gates_on_low_qubits_becomes_one_dense_block(qcs):
    """Inside a fuse window a run of arbitrary (controlled) single-qubit gates on qubits 0..3 is multiplied
    up on the host and leaves as ONE DMMA dense block; the result equals the gates applied one by one."""
    n = 16
    rng = np.random.default_rng(31)

    def unitary():
        a = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
        qm, _ = np.linalg.qr(a)
        return qm

    gates = []
    for _ in range(40):
        t = int(rng.integers(4))
        if rng.random() < 0.5:
            gates.append((t, -1, unitary()))
        else:
            c = int(rng.choice([x for x in range(4) if x != t]))
            gates.append((t, c, unitary()))
    with qcs.Register(n, 0) as one, qcs.Register(n, 0) as fused:
        for reg in (one, fused):
            reg.fill_synthetic(3)
            reg.scale(1.0 / math.sqrt(reg.norm2()))
        for t, c, U in gates:
            one.apply_gate(t, U) if c < 0 else one.apply_controlled_gate(c, t, U)
        before = fused.launch_count
        with fused.fused():
            for t, c, U in gates:
                fused.apply_gate(t, U) if c < 0 else fused.apply_controlled_gate(c, t, U)
            assert fused.fuse_pending == len(gates) and fused.launch_count == before     # nothing launched yet
            fused.hadamard_gate(9)                     # ends the run: the dense block goes first
            assert fused.launch_count == before + 1 and fused.fuse_pending == 1
            fused.apply_gate(2, gates[0][2])           # a new run after the Hadamard
        one.hadamard_gate(9)
        one.apply_gate(2, gates[0][2])
        assert fused.launch_count == before + 3        # dense block, Hadamard, dense block
        assert rel_l2(fused.get_state(), one.get_state()) <= 1e-12
        # gates on higher qubits are not recorded: they run at once, after the pending block
        with fused.fused():
            fused.apply_gate(1, gates[1][2])
            fused.apply_gate(7, gates[2][2])
            assert fused.fuse_pending == 0
        one.apply_gate(1, gates[1][2])
        one.apply_gate(7, gates[2][2])
        assert rel_l2(fused.get_state(), one.get_state()) <= 1e-12
