"""CPU tests of the gate-stream scheduler (csrc/circuit.cu) through its host-only view
qcs_schedule_describe: the passes it emits, replayed with numpy (H on every qubit of a pass in
any order, then its diagonal gates), must reproduce in-order application of the recorded
hadamard_gate / c_phase_shift_gate stream (qc_shor.c:442-565) -- i.e. nothing was moved across
a gate it does not commute with -- and every recorded gate must be placed exactly once."""
import math

import numpy as np
import pytest

from quantumcomputer_b200.workloads import layered_circuit


def apply_h(v, n, q):
    t = v.reshape(1 << (n - 1 - q), 2, 1 << q)
    a, b = t[:, 0, :].copy(), t[:, 1, :].copy()
    t[:, 0, :] = (a + b) / math.sqrt(2.0)
    t[:, 1, :] = (a - b) / math.sqrt(2.0)


def apply_cp(v, n, c, q, theta):
    idx = np.arange(1 << n)
    sel = ((idx >> c) & 1).astype(bool) & ((idx >> q) & 1).astype(bool)
    v[sel] *= np.exp(1j * theta)


def in_order(v, n, gates):
    v = v.copy()
    for g in gates:
        if g[0] == "h":
            apply_h(v, n, g[1])
        else:
            apply_cp(v, n, g[1], g[2], g[3])
    return v


def replay(v, n, gates, passes):
    v = v.copy()
    placed = []
    for ps in passes:
        if ps["type"] != "before":
            for q in ps["h"]:
                apply_h(v, n, q)
        for i in ps["diag"] + ps["after"]:
            g = gates[i]
            assert g[0] == "cp"
            apply_cp(v, n, g[1], g[2], g[3])
            placed.append(i)
    return v, placed


def check(qcs, n, gates, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    v /= np.linalg.norm(v)
    passes = qcs.schedule_describe(n, gates)
    got, placed = replay(v, n, gates, passes)
    want = in_order(v, n, gates)
    assert sorted(placed) == [i for i, g in enumerate(gates) if g[0] == "cp"], "every diagonal gate exactly once"
    n_h = sum(len(ps["h"]) for ps in passes if ps["type"] != "before")
    assert n_h == sum(1 for g in gates if g[0] == "h"), "every Hadamard exactly once"
    assert np.linalg.norm(got - want) <= 1e-12
    return passes


@pytest.mark.parametrize("n,layers", [(5, 3), (12, 4), (14, 3), (16, 2)])
def test_layered_circuit_schedule_is_legal(qcs, n, layers):
    gates = layered_circuit(n, layers)
    passes = check(qcs, n, gates, n)
    # far fewer passes than gates, and the diagonal gates ride in the sweeps
    assert len(passes) <= layers * 3
    assert sum(len(ps["diag"]) for ps in passes) == layers * n


@pytest.mark.parametrize("n,seed", [(4, 1), (9, 2), (13, 3), (14, 4), (15, 5), (16, 6)])
def test_random_stream_schedule_is_legal(qcs, n, seed):
    rng = np.random.default_rng(seed)
    gates = []
    for _ in range(int(rng.integers(30, 90))):
        kind = rng.random()
        if kind < 0.4:
            gates.append(("h", int(rng.integers(n))))
        elif kind < 0.5:
            lo = int(rng.integers(n))
            gates += [("h", q) for q in range(lo, int(rng.integers(lo, n)) + 1)]
        else:
            gates.append(("cp", int(rng.integers(n)), int(rng.integers(n)), float(rng.uniform(-3, 3))))
    if seed % 2:
        gates = [("cp", 0, n - 1, 0.3), ("cp", n // 2, n // 2, 1.1)] + gates
    check(qcs, n, gates, seed)


def test_diagonal_only_and_single_hadamards(qcs):
    n = 13
    gates = [("cp", 1, 2, 0.1), ("cp", 3, 3, 0.2), ("h", 7), ("cp", 7, 1, 0.3), ("h", 7), ("h", 2),
             ("cp", 2, 7, 0.4)] + [("cp", q, (q + 3) % n, 0.01 * q) for q in range(n)]
    passes = check(qcs, n, gates, 7)
    # gates on qubits that no Hadamard ever touches (and the two leading ones) precede every pass
    assert passes[0]["type"] == "before" and passes[0]["after"][:2] == [0, 1]
    assert [ps["type"] for ps in passes[1:]] == ["hadamard", "hadamard", "hadamard"]


def test_sharded_schedule_places_global_hadamards(qcs):
    """world_size 4: Hadamards on the two global qubits become a `global` pass; diagonal gates with
    a global qubit appear on a rank only where its bit is set."""
    n = 16
    gates = [("h", q) for q in range(n)] + [("cp", n - 1, 3, 0.5), ("cp", n - 2, n - 1, 0.25), ("cp", 2, 5, 1.0)]
    for rank in range(4):
        passes = qcs.schedule_describe(n, gates, world_size=4, rank=rank)
        assert passes[0]["type"] == "global" and passes[0]["h"] == [n - 2, n - 1]
        placed = sorted(i for ps in passes for i in ps["diag"] + ps["after"])
        want = [n + 2]
        if rank & 2:
            want.append(n)
        if rank == 3:
            want.append(n + 1)
        assert placed == sorted(want), (rank, placed)


def test_bad_arguments(qcs):
    with pytest.raises(qcs.QcsError):
        qcs.schedule_describe(4, [("h", 4)])
    with pytest.raises(qcs.QcsError):
        qcs.schedule_describe(4, [("h", 1)], world_size=3)


# ---------------------------------------------------------------------------
# the sweep planner (csrc/qft_common.cuh plan_inverse) through qcs_plan_describe
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("n,lo,tile_bits,run_bits", [(7, 4, 0, 0), (10, 5, 0, 0), (14, 0, 0, 0), (20, 3, 11, 4),
                                                     (30, 0, 0, 0), (30, 0, 11, 0), (31, 0, 0, 0), (33, 0, 0, 0),
                                                     (33, 0, 11, 4), (35, 5, 0, 0), (26, 0, 9, 3), (18, 2, 13, 3)])
def test_sweep_plan_covers_every_stage_once_in_order(qcs, n, lo, tile_bits, run_bits):
    """The stages l = n-1 .. lo of inverse_QFT (qc_shor.c:682-689) appear exactly once, top first;
    every tile is a contiguous run [0, a) plus the run [g_lo, g_hi) that holds its stage bits; the
    tiles of a sweep partition the register; the scalings multiply to 2^(-stages/2)."""
    sweeps = qcs.plan_describe(n, lo, n, tile_bits, run_bits)
    t = min(tile_bits or 12, n)
    expect = n - 1
    scale = 1.0
    for sw in sweeps:
        tile = sw["a"] + (sw["g_hi"] - sw["g_lo"])
        assert tile == t and sw["tiles"] == 1 << (n - t)
        assert sw["a"] <= sw["g_lo"] <= sw["g_hi"] <= n
        for low, r in sw["steps"]:
            assert 1 <= r <= 4
            assert low + r - 1 == expect, "stages must come top first without gaps"
            in_run = low >= sw["g_lo"] and low + r <= sw["g_hi"]
            in_low = low + r <= sw["a"]
            assert in_run or in_low, "a step's bits lie inside the tile"
            expect = low - 1
        scale *= sw["scale"]
    assert expect == lo - 1
    assert abs(scale - 2.0 ** (-(n - lo) / 2.0)) <= 1e-15 * scale


def test_sweep_counts_of_the_benchmark_sizes(qcs):
    """n = 30 is three sweeps with 2^12 tiles (128 B runs), four with 2^11; n = 33 four / five."""
    assert len(qcs.plan_describe(30)) == 3
    assert [s["a"] for s in qcs.plan_describe(30)] == [3, 3, 12]
    assert len(qcs.plan_describe(30, tile_bits=11)) == 4
    assert len(qcs.plan_describe(33)) == 4 and qcs.plan_describe(33)[0]["a"] == 5
    assert len(qcs.plan_describe(33, tile_bits=11, min_run_bits=4)) == 5
    with pytest.raises(qcs.QcsError):
        qcs.plan_describe(10, 5, 5)


@pytest.mark.parametrize("n,lo,expect_pairs", [(24, 0, 1), (28, 0, 1), (30, 0, 0), (31, 0, 1), (33, 0, 2), (35, 0, 2), (36, 0, 2),
                                               (33, 5, None), (27, 4, None)])
def test_l2_paired_launch_bookkeeping(qcs, n, lo, expect_pairs):
    """Host-only check of the paired-sweep launch (qft_pipeline.cu): the ticket order hands out every tile
    of both sweeps once and a second-sweep tile only after all first-sweep tiles of its block, and the tiles
    of both sweeps of a block cover the same amplitudes -- for several lags."""
    for lag in (0, 7, 148, 444, 100000):
        got = qcs.lib().qcs_pair_selfcheck(n, lo, n, lag)
        assert got >= 0, (n, lo, lag, got)
        if expect_pairs is not None:
            assert got == expect_pairs, (n, lag, got)
