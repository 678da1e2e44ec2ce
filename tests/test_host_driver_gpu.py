"""GPU: the C host driver (qc_shor_b200) end to end against the reference's own runs
(tests/golden/shor_runs.json: shors_algorithm of the unmodified reference, same seeds)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "quantumcomputer_b200", "bin", "qc_shor_b200")


def run_driver(run, extra=()):
    cmd = [BIN, "-C", str(run["C"]), "-L", str(run["L"]), "-M", str(run["M"]), "-s", str(run["seed"])]
    if run["a"]:
        cmd += ["-a", str(run["a"])]
    out = subprocess.run(cmd + list(extra), capture_output=True, text=True, timeout=120)
    m = re.search(r"Factors of (\d+) found: \((\d+), (\d+)\)", out.stdout)
    return out.returncode, (int(m.group(2)), int(m.group(3))) if m else None, out.stdout


@pytest.mark.parametrize("extra", [(), ("-x",)])
def test_factors_identical_to_reference_runs(extra):
    """-x = gate-by-gate reference-order kernels, default = fused sweeps."""
    assert os.path.exists(BIN), "build the host driver with `make host`"
    for run in load_golden("shor_runs.json")["runs"]:
        rc, factors, stdout = run_driver(run, extra)
        assert rc == run["error"], (run, stdout)
        assert list(factors) == run["factors"], (run, stdout)


def test_cli_behaviour():
    out = subprocess.run([BIN, "-L", "3", "-M", "4"], capture_output=True, text=True)
    assert out.returncode == 2 and "not given" in out.stderr          # BAD_ARGUMENTS, qc_shor.c:1222-1236
    out = subprocess.run([BIN, "-C", "15", "-L", "3", "-M", "4", "-f", "7", "-s", "1", "-v"],
                         capture_output=True, text=True)               # -f: the documented spelling of -a
    assert out.returncode == 0 and "Forced trial integer a = 7" in out.stdout and "(5, 3)" in out.stdout
    out = subprocess.run([BIN, "-C", "15", "-L", "3", "-M", "4", "-a", "7", "-s", "1", "-r", "-V"],
                         capture_output=True, text=True)               # robust post-processing
    assert out.returncode == 0 and "(5, 3)" in out.stdout and "Measuring state" in out.stdout
