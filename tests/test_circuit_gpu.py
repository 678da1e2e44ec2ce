"""GPU parity tests of the deferred gate stream (qcs_fuse_begin / qcs_fuse_end): arbitrary
runs of hadamard_gate / c_phase_shift_gate scheduled into Walsh-Hadamard tile sweeps with the
diagonal gates riding along, against the oracle applying the same gates one by one
(qc_shor.c:442-565 via operate_matrix, qc_shor.c:370-420).

Bar (north_star): amplitude L2 error <= 1e-12 relative."""
import math

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


from quantumcomputer_b200.workloads import apply_gates as apply, layered_circuit


def start_state(oracle_built, n, seed):
    o = oracle_built.Restatement(n, 0)
    o.fill_synthetic(seed)
    o.scale(1.0 / math.sqrt(o.norm2()))
    return o, o.get_state().copy()


@pytest.mark.parametrize("n,layers,tile_bits,pipeline", [(7, 3, 0, 1), (12, 4, 0, 1), (14, 8, 0, 1), (14, 3, 9, 0),
                                                         (17, 4, 0, 1), (18, 3, 0, 0), (20, 8, 0, 1), (22, 2, 0, 1)])
def test_layered_circuit_matches_oracle(qcs, oracle_built, n, layers, tile_bits, pipeline):
    gates = layered_circuit(n, layers)
    o, base = start_state(oracle_built, n, 77 + n)
    apply(o, gates)
    want = o.get_state()
    with qcs.Register(n, 0) as reg:
        reg.set_option(qcs.OPT_TILE_BITS, tile_bits)
        reg.set_option(qcs.OPT_PIPELINE, pipeline)
        reg.set_state(base)
        before = reg.launch_count
        with reg.fused():
            apply(reg, gates)
            assert reg.fuse_pending == len(gates)
        assert reg.fuse_pending == 0
        launches = reg.launch_count - before
        err = rel_l2(reg.get_state(), want)
        assert err <= TOL, err
        # the point of the scheduler: far fewer passes than gates
        assert launches < len(gates) / 4 or n < 12, (launches, len(gates))


@pytest.mark.parametrize("n,seed", [(3, 1), (6, 2), (11, 3), (13, 4), (15, 5), (16, 6), (19, 7)])
def test_random_gate_stream_matches_oracle(qcs, oracle_built, n, seed):
    """Random H / C-phase streams: repeated H on a qubit, phase gates with c == q, isolated
    Hadamards, diagonal gates before any Hadamard, long diagonal runs."""
    rng = np.random.default_rng(seed)
    gates = []
    for _ in range(int(rng.integers(20, 120))):
        kind = rng.random()
        if kind < 0.45:
            gates.append(("h", int(rng.integers(n))))
        elif kind < 0.55:
            lo = int(rng.integers(n))
            hi = int(rng.integers(lo, n)) + 1
            gates += [("h", q) for q in range(lo, hi)]
        else:
            gates.append(("cp", int(rng.integers(n)), int(rng.integers(n)), float(rng.uniform(-7, 7))))
    if seed % 2:
        gates = [("cp", 0, n - 1, 0.3), ("cp", n // 2, n // 2, 1.1)] + gates
    gates += [("cp", int(rng.integers(n)), int(rng.integers(n)), float(rng.uniform(-7, 7))) for _ in range(70)]
    o, base = start_state(oracle_built, n, 5 + seed)
    apply(o, gates)
    want = o.get_state()
    with qcs.Register(n, 0) as reg:
        reg.set_state(base)
        with reg.fused():
            apply(reg, gates)
        assert rel_l2(reg.get_state(), want) <= TOL


def test_any_other_call_flushes_the_stream(qcs, oracle_built):
    n = 12
    gates = layered_circuit(n, 2)
    o, base = start_state(oracle_built, n, 9)
    apply(o, gates)
    r = 0.4321
    want_index = o.measure_state(r)
    with qcs.Register(n, 0) as reg:
        reg.set_state(base)
        reg.fuse_begin()
        apply(reg, gates)
        assert reg.fuse_pending == len(gates)
        # measure_state is not recordable: it launches what was recorded, then measures
        assert reg.measure_state(r) == want_index
        assert reg.fuse_pending == 0
        reg.hadamard_gate(3)
        assert reg.fuse_pending == 1
        reg.fuse_end()
        with pytest.raises(qcs.QcsError):
            reg.fuse_end()


def test_fusion_off_runs_every_gate_at_once_and_value_identical(qcs, oracle_built):
    from conftest import values_equal
    n = 11
    gates = layered_circuit(n, 2)
    o, base = start_state(oracle_built, n, 10)
    apply(o, gates)
    with qcs.Register(n, 0) as reg:
        reg.set_option(qcs.OPT_FUSION, 0)
        reg.set_state(base)
        with reg.fused():
            apply(reg, gates)
            assert reg.fuse_pending == 0
        assert values_equal(reg.get_state(), o.get_state())


def test_layered_circuit_round_trip_large(qcs):
    """n = 26 (1 GiB): the circuit followed by its inverse (gates reversed, -theta) restores the
    state -- a size-independent property at a size the oracle cannot reach in seconds."""
    n = 26
    gates = layered_circuit(n, 3)
    inverse = [g if g[0] == "h" else ("cp", g[1], g[2], -g[3]) for g in reversed(gates)]
    with qcs.Register(n, 0) as reg:
        reg.fill_synthetic(99)
        reg.scale(1.0 / math.sqrt(reg.norm2()))
        probe = [0, 1, 12345, (1 << n) - 1, (1 << 25) + 17]
        before = np.array([reg.get_state(i, 1)[0] for i in probe])
        with reg.fused():
            apply(reg, gates)
        mid = np.array([reg.get_state(i, 1)[0] for i in probe])
        assert abs(reg.norm2() - 1.0) < 1e-12
        assert np.linalg.norm(mid - before) > 1e-6
        with reg.fused():
            apply(reg, inverse)
        after = np.array([reg.get_state(i, 1)[0] for i in probe])
        assert np.linalg.norm(after - before) <= 1e-12 * math.sqrt(len(probe)) * np.max(np.abs(before)) * 100


def test_more_than_1024_diagonal_gates_in_one_window(qcs, oracle_built):
    """A standalone diagonal list longer than the 1024-slot scratch region must not overwrite the
    in-sweep gate tables that later sweeps read (they live right behind the scratch slots)."""
    n = 13
    rng = np.random.default_rng(2025)
    head = [("cp", int(rng.integers(n)), int(rng.integers(n)), float(rng.uniform(-3, 3))) for _ in range(2500)]
    # a layer whose sweeps carry in-sweep diagonal gates, then another long diagonal run
    tail = layered_circuit(n, 2)
    tail += [("cp", int(rng.integers(n)), int(rng.integers(n)), float(rng.uniform(-3, 3))) for _ in range(1500)]
    gates = head + tail
    o, base = start_state(oracle_built, n, 41)
    apply(o, gates)
    want = o.get_state()
    with qcs.Register(n, 0) as reg:
        reg.set_state(base)
        with reg.fused():
            apply(reg, gates)
        assert rel_l2(reg.get_state(), want) <= TOL
