"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_register_matches_single_gpu(qcs, world):
    if qcs.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout[-4000:])
    sys.stderr.write(out.stderr[-2000:])
    assert out.returncode == 0 and "DIST_OK" in out.stdout


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_single_process_multi_gpu_register(qcs, n_gpus):
    """qcs_register_create_multi: the whole sharded register behind ONE handle in ONE process (what
    the C host uses, qc_shor.c:1316-1324 / 922-928) against a single-GPU register running the same calls."""
    import math
    import numpy as np
    if qcs.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    # full-size cfg2 (BASELINE configs[1]), intended pow mode: state, norm, measured index
    L, M, Cn, a = 10, 5, 21, 2
    with qcs.Register(L, M) as one, qcs.Register(L, M, n_gpus=n_gpus) as many:
        assert many.num_gpus == n_gpus and many.local_states == many.num_states == 1 << (L + M)
        for reg in (one, many):
            reg.reset_register()
            reg.quantum_computation(Cn, a, qcs.POW_MODULAR)
        want, got = one.get_state(), many.get_state()
        assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)
        assert abs(many.norm2() - one.norm2()) < 1e-13
        for r in (0.1, 0.6059782775118947, 0.95):
            assert many.sample_states([r])[0] == one.sample_states([r])[0]
        assert many.measure_state(0.6059782775118947) == one.measure_state(0.6059782775118947)
        assert many.nonzero_states()[0] == one.nonzero_states()[0]
    # synthetic n = 26 inverse QFT + forward: the peer-memory sweeps from one process
    n = 26
    with qcs.Register(n, 0) as one, qcs.Register(n, 0, n_gpus=n_gpus) as many:
        for reg in (one, many):
            reg.fill_synthetic(11)
            reg.scale(1.0 / math.sqrt(reg.norm2()))
            reg.inverse_QFT()
        probes = [0, 1, (1 << n) - 1, (1 << (n - 1)) + 12345, 0x1555555, 777777]
        for i in probes:
            w, g = one.get_state(i, 1)[0], many.get_state(i, 1)[0]
            assert abs(g - w) <= 1e-12 * 2.0 ** (-n / 2), i
        many.QFT()
        one.QFT()
        for i in probes:
            assert abs(many.get_state(i, 1)[0] - one.get_state(i, 1)[0]) <= 1e-12 * 2.0 ** (-n / 2)


def test_c_host_driver_sharded(qcs):
    """qc_shor_b200 -g 2: the reference's main() flow (qc_shor.c:1284-1348) on a register sharded over
    two GPUs gives the factors of the single-GPU run."""
    import re
    if qcs.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    binary = os.path.join(ROOT, "quantumcomputer_b200", "bin", "qc_shor_b200")
    outs = []
    for extra in ([], ["-g", "2"]):
        out = subprocess.run([binary, "-C", "21", "-L", "10", "-M", "5", "-a", "2", "-s", "2021", "-r"] + extra,
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        outs.append(re.search(r"Factors of 21 found: \((\d+), (\d+)\)", out.stdout).groups())
    assert outs[0] == outs[1] and sorted(map(int, outs[0])) == [3, 7]
