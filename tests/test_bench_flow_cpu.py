"""CPU dry run of bench.py's GPU arm: its control flow, its in-run parity block and the JSON contract of the
line it prints, with the engine's Python mirror replaced by a stand-in over the CPU oracle (test
infrastructure: bench.py itself never sees the oracle on this arm -- the stand-in is injected by this test).
What it pins: every key the driver reads is present and well-formed, the parity block's own arithmetic
(closed form at bit-reversed indices, round trip, sample_states vs locate vs measure_state) is right, a failed
parity block flags the line and exits 3, and the Shor block's bookkeeping."""
import argparse
import json
import math
import os
import sys
import time
import types

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)


def make_fake_q(oracle, break_transform=False):
    import quantumcomputer_b200 as real          # constants and class names only; no library call is made
    classes = list(real.KERNEL_CLASSES)

    class Register:
        def __init__(self, L_size, M_size, device=-1, rank=0, world_size=1, comm_id=None, n_gpus=1):
            assert world_size == 1
            self.L, self.M, self.n = L_size, M_size, L_size + M_size
            self.o = oracle.Restatement(L_size, M_size)
            self.opts = {real.OPT_FUSION: 1, real.OPT_PIPE_SHAPE: -1, real.OPT_MIN_RUN_BITS: 3}
            self.launches = 0
            self.prof = {c: [0, 0.0, 0.0] for c in classes}
            self.t0 = 0.0

        local_states = property(lambda self: 1 << self.n)
        launch_count = property(lambda self: self.launches)

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            self.close()

        def close(self):
            if self.o is not None:
                self.o.close()
                self.o = None

        def _account(self, cls, launches, nbytes, ms=0.01):
            self.launches += launches
            p = self.prof[cls]
            p[0] += launches
            p[1] += ms * launches
            p[2] += nbytes * launches

        def set_option(self, opt, value):
            self.opts[opt] = value

        def get_option(self, opt):
            return self.opts.get(opt, 0)

        def synchronize(self):
            pass

        def reset_register(self):
            self.o.reset_register()

        def hadamard_gate(self, q):
            self.o.hadamard_gate(q)
            self._account("hadamard", 1, 32.0 * (1 << self.n))

        def c_phase_shift_gate(self, c, q, theta):
            self.o.c_phase_shift_gate(c, q, theta)
            self._account("cphase", 1, 8.0 * (1 << self.n))

        def inverse_QFT(self):
            self.o.inverse_QFT()
            if break_transform:
                self.o.c_phase_shift_gate(0, 1, 0.3)                 # a wrong (but unitary) transform
            self._account("tile_sweep", 3, 32.0 * (1 << self.n))

        def QFT(self):
            if break_transform:
                self.o.c_phase_shift_gate(0, 1, -0.3)
            for l in range(self.M, self.n):                          # the adjoint: reverse order, opposite angles
                for k in range(self.M, l):
                    self.o.c_phase_shift_gate(l, k, -math.pi / float(1 << (l - k)))
                self.o.hadamard_gate(l)
            self._account("tile_sweep", 3, 32.0 * (1 << self.n))

        def quantum_computation(self, Cn, a, pow_mode=0):
            self.o.quantum_computation(Cn, a, pow_mode)
            self._account("tile_sweep", 2, 32.0 * (1 << self.n))

        def measure_state(self, r):
            self._account("reduce", 1, 16.0 * (1 << self.n))
            return int(self.o.measure_state(r))

        def sample_states(self, rs):
            keep = self.o.get_state().copy()
            out = []
            for r in rs:
                out.append(int(self.o.measure_state(r)))
                self.o.set_state(keep)
            return out

        def norm2(self):
            return float(self.o.norm2())

        def fill_synthetic(self, seed):
            self.o.fill_synthetic(seed)

        def scale(self, f):
            self.o.scale(f)

        def get_state(self, first=0, count=None, out=None):
            s = self.o.get_state()
            count = len(s) - first if count is None else count
            piece = s[first:first + count]
            if out is not None:
                out[:2 * count] = piece.view(np.float64)
                return out[:2 * count].view(np.complex128)
            return piece.copy()

        def set_state(self, amps, first=0):
            s = self.o.get_state().copy()
            a = np.asarray(amps)
            if a.dtype == np.float64:
                a = a.view(np.complex128)
            s[first:first + len(a)] = a
            self.o.set_state(s)

        def set_state_async(self, float64_array, first=0):
            self.set_state(float64_array, first)

        def fused(self):
            import contextlib
            return contextlib.nullcontext()

        def timer_start(self):
            self.t0 = time.perf_counter()

        def timer_stop(self):
            return 1e3 * (time.perf_counter() - self.t0) + 1e-3

        def profile_reset(self):
            self.prof = {c: [0, 0.0, 0.0] for c in classes}
            self.launches = 0

        def profile(self):
            return {c: tuple(v) for c, v in self.prof.items()}

    class PinnedBuffer:
        def __init__(self, n_doubles, device=-1):
            self.array = np.zeros(n_doubles, dtype=np.float64)

        def close(self):
            self.array = None

    fake = types.SimpleNamespace(Register=Register, PinnedBuffer=PinnedBuffer)
    for name in dir(real):
        if name.startswith(("OPT_", "POW_")):
            setattr(fake, name, getattr(real, name))
    return fake


def bench_args(**over):
    a = argparse.Namespace(gpus=1, steps=3, warmup=3, impl="ours", qubits=12, no_fusion=False, tile_bits=0, prefetch=-1,
                           workload="iqft", layers=8, pipe_shape=-1, min_run_bits=0, global_run_bits=0, overlap_slices=-1,
                           global_sms=0, no_e2e=False, e2e_steps=2, no_cpu_baseline=True, no_parity=False,
                           no_north_star=False, no_configs=False, north_star_steps=5, l2_pair=-1)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def run_arm(monkeypatch, capsys, fake, args):
    import bench
    monkeypatch.setitem(sys.modules, "quantumcomputer_b200", fake)
    for var in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(var, raising=False)
    code = 0
    try:
        bench.run_ours(args)
    except SystemExit as e:
        code = e.code
    out = capsys.readouterr()
    lines = [l for l in out.out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.out                                  # ONE JSON line on stdout
    return code, json.loads(lines[0]), out.err


def test_gpu_arm_line_has_the_contract_keys(monkeypatch, capsys, oracle_built):
    code, line, _ = run_arm(monkeypatch, capsys, make_fake_q(oracle_built), bench_args())
    assert code == 0 and "error" not in line
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "parity"):
        assert key in line, key
    assert line["metric"] == "qft_gates_per_sec" and line["unit"] == "gates/s" and line["n_gpus"] == 1
    assert line["steps"] == 3 and line["warmup"] == 3 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["scaling"] == "weak"
    assert "inverse_QFT over all n=12 qubits" in line["config"]["workload"] and "model" not in line["config"]
    assert line["config"]["gates_per_step"] == 12 + 66
    assert abs(line["value"] - 78 * 3 / (line["ms_per_step"] * 3e-3)) < 1e-6 * line["value"]
    roof = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in roof, key
    assert roof["bound"] == "hbm" and roof["kernel"] == "tile_sweep" and roof["unit"] == "GB/s"
    assert roof["launches"] == 9 and roof["algorithmic_bytes_per_launch"] == 32.0 * 4096
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    assert line["gpu_launches"] == 9                                  # three sweeps per step, three steps
    e2e = line["e2e"]
    assert e2e["unit"] == "gates/s" and e2e["h2d_bytes_per_step"] == 16 * 4096 and e2e["d2h_bytes_per_step"] == 8
    assert e2e["value"] > 0 and e2e["steps"] == 2
    par = line["parity"]
    assert par["ok"] and par["max_rel_err"] <= 1e-12 and par["closed_form_probes"] >= 10
    assert par["measure"]["ok"] and par["measure"]["sample_states_equals_locate"]
    assert abs(line["config"]["norm_after"] - 1.0) < 1e-12


def test_failed_parity_flags_the_line_and_exits_3(monkeypatch, capsys, oracle_built):
    code, line, err = run_arm(monkeypatch, capsys, make_fake_q(oracle_built, break_transform=True), bench_args(no_e2e=True))
    assert code == 3
    assert line["error"] == "parity check failed" and line["parity"]["ok"] is False
    assert line["parity"]["closed_form_max_rel_err"] > 1e-6          # the closed form catches a wrong unitary
    assert "parity check failed" in err                               # said at once, before anything is timed


def test_shor_block_bookkeeping(monkeypatch, oracle_built):
    import bench
    fake = make_fake_q(oracle_built)                                  # handed in as `q`; the workloads module is the real one
    ranks = types.SimpleNamespace(local_rank=0, rank=0, world=1)
    out = bench.shor_block(fake, ranks, bench_args(steps=5), with_n30=False)
    assert out["parity"]["ok"] and len(out["find_period"]) == 3
    cfg1 = out["find_period"][0]
    assert (cfg1["C"], cfg1["a"], cfg1["L"], cfg1["M"], cfg1["gates"]) == (15, 7, 3, 4, 12)
    assert cfg1["measured_indices_identical"] and cfg1["cpu_kind"] in ("reference", "port") and cfg1["indices"][0] == 55


def test_layered_workload_flow(monkeypatch, capsys, oracle_built):
    """--workload layered (BASELINE configs[3]) through the same arm: the gates are issued one by one through the
    reference's operator names inside a fuse window."""
    code, line, _ = run_arm(monkeypatch, capsys, make_fake_q(oracle_built),
                            bench_args(workload="layered", qubits=9, layers=2, steps=2, no_e2e=True))
    assert code == 0 and "error" not in line
    assert line["metric"] == "layered_circuit_gates_per_sec" and line["config"]["gates_per_step"] == 36
    assert "layered circuit (BASELINE configs[3])" in line["config"]["workload"]
    assert line["gpu_launches"] == 2 * 36 and "parity" not in line and "cpu_baseline" not in line
    assert abs(line["config"]["norm_after"] - 1.0) < 1e-12


def test_north_star_and_layered_blocks(oracle_built):
    """The two blocks the default N = 1 line carries besides the headline (BASELINE's n = 33 transform and
    configs[3]), at a size the oracle finishes in a moment."""
    import bench
    fake = make_fake_q(oracle_built)
    ranks = types.SimpleNamespace(local_rank=0, rank=0, world=1, comm_id=lambda: None, barrier=lambda reg=None: None,
                                  max=lambda x: x)
    ns = bench.north_star_block(fake, ranks, bench_args(qubits=0), 6500.0, "test", n_override=11)
    assert ns["qubits"] == 11 and ns["parity"]["ok"] and ns["parity"]["closed_form_max_rel_err"] <= 1e-12
    assert ns["gates_per_step"] == 11 + 55 and ns["steps"] == 3 and ns["gpu_launches"] == 9
    assert ns["roofline"]["kernel"] == "tile_sweep" and ns["ms_per_qft"] > 0
    lay = bench.layered_block(fake, ranks, bench_args(layers=2), 6500.0, "test", n_override=9)
    assert lay["parity"]["ok"] and lay["parity"]["round_trip_max_rel_err"] <= 1e-11
    assert lay["gates_per_step"] == 36 and lay["qubits"] == 9 and lay["steps"] == 3
    broken = bench.north_star_block(make_fake_q(oracle_built, break_transform=True), ranks, bench_args(qubits=0), 6500.0,
                                    "test", n_override=11)
    assert broken["parity"]["ok"] is False
