"""CPU tests of bench.py: the reference arm's JSON line carries the contract's keys, and the
GPU arm fails loudly (no CPU fallback) when no CUDA device is present."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=300, cwd=ROOT)


def test_reference_arm_line(oracle_built):
    out = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--qubits", "8")
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "qft_gates_per_sec" and line["unit"] == "gates/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and line["config"]["qubits"] == 8 and line["config"]["gates_per_step"] == 36
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_gpu_arm_fails_loudly_without_a_gpu(qcs):
    if qcs.device_count() > 0:
        pytest.skip("a CUDA device is present")
    out = run_bench("--steps", "1", "--warmup", "3", "--no-e2e", "--no-cpu-baseline")
    assert out.returncode != 0
    assert "no usable CUDA device" in out.stderr and "no CPU path" in out.stderr
    assert not out.stdout.strip(), "no bench line may be printed without a GPU"


def test_bench_helpers_on_cpu(oracle_built):
    """The pieces of bench.py that run on the host: the closed-form index helper, the size of the reference
    sample, and the matrix-free CPU baseline (one thread and, when OpenMP is available, all cores)."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.bitrev(0b0011, 4) == 0b1100 and bench.bitrev(1, 30) == 1 << 29
    assert bench.qft_gate_count(30) == 465
    assert bench.reference_n_for(25) == bench.REF_N                       # the driver's 20 + 5 runs keep n = 12
    assert bench.reference_n_for(500) < bench.REF_N
    one = bench.port_gate_rate(12, False, 0.3)
    assert one["kind"] == "port" and one["cores"] == 1 and one["value"] > 0 and one["qubits"] == 12
    if oracle_built.have_restatement_omp():
        many = bench.port_gate_rate(12, True, 0.3)
        assert many["cores"] >= 1 and many["value"] > 0
    traffic, src = bench.profiled_traffic(30)
    assert traffic is None or (3.3e10 < traffic < 3.6e10 and src.startswith("profiles/"))


def test_workload_generators_follow_the_reference_rng(oracle_built):
    """BASELINE configs[3] draws its angles from gsl_rng_mt19937 / gsl_rng_uniform: the Python
    generator behind bench.py --workload layered against the golden draws of the unmodified reference's RNG
    (KAT-3) and, where oracle/_ref exists, against the reference's RNG live; then the circuit's definition."""
    import math
    from conftest import load_golden
    from quantumcomputer_b200.workloads import layered_circuit, mt19937_uniforms
    g = load_golden("scalars.json")["mt19937"]
    assert [x.hex() for x in mt19937_uniforms(5489, 4)] == g["seed5489_first_uniform"]
    assert [x.hex() for x in mt19937_uniforms(0, 2)] == g["seed0_first_uniform"] == g["seed4357_first_uniform"]
    if oracle_built.have_reference():
        ref = oracle_built.Reference(1, 1)
        for seed in (33, 34, 40, 12345):
            ref.seed(seed)
            assert mt19937_uniforms(seed, 700) == [ref.rng_uniform() for _ in range(700)]     # past one state refill
        ref.close()
    n, layers = 33, 8
    gates = layered_circuit(n, layers)
    assert len(gates) == 528                                              # 66 gates per layer
    for d in range(layers):
        layer = gates[d * 2 * n:(d + 1) * 2 * n]
        u = mt19937_uniforms(n + d, n)
        assert layer[:n] == [("h", q) for q in range(n)]
        assert layer[n:] == [("cp", q, (q + 1 + d) % n, 2.0 * math.pi * u[q]) for q in range(n)]
