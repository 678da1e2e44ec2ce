"""Every line the C host driver prints, against the unmodified reference's own output
(tests/golden/shor_stdout.json: shors_algorithm, qc_shor.c:1003-1134, same seed, same -v / -V flags;
main's closing lines, qc_shor.c:1335-1341).

CPU: the driver's calls into libqcs.so are answered by tests/mock/mock_qcs.c (the CPU oracle behind the
same C ABI, injected with LD_PRELOAD -- test infrastructure; the product library has no CPU path), which
isolates the host-side logic.  GPU: the same comparison with the real library on a device."""
import os
import re
import subprocess

import pytest

from conftest import ROOT, load_golden

BIN = os.path.join(ROOT, "quantumcomputer_b200", "bin", "qc_shor_b200")
MOCK_SRC = os.path.join(ROOT, "tests", "mock", "mock_qcs.c")
MOCK_SO = os.path.join(ROOT, "tests", "mock", "_build", "libqcs_mock.so")


@pytest.fixture(scope="module")
def mock_so(oracle_built):
    orc_dir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(os.path.dirname(MOCK_SO), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "oracle"), "-o", MOCK_SO, MOCK_SRC,
                    "-L" + orc_dir, "-lqcsoracle", "-Wl,-rpath," + orc_dir], check=True)
    return MOCK_SO


def expected_stdout(case):
    """shors_algorithm's output, then main's closing lines (qc_shor.c:1335-1341)."""
    want = case["stdout"]
    if case["error"] == 0:
        want += " --- Factors of %d found: (%d, %d).\n" % (case["C"], case["factors"][0], case["factors"][1])
        if case["C"] // case["factors"][0] != case["factors"][1]:
            want += " --- These factors are incorrect. Consider increasing register sizes as per the warnings.\n"
    return want


def run_case(case, env=None):
    cmd = [BIN, "-C", str(case["C"]), "-L", str(case["L"]), "-M", str(case["M"]), "-s", str(case["seed"])]
    if case["a"]:
        cmd += ["-a", str(case["a"])]
    if case["very_verbose"]:
        cmd += ["-V"]
    elif case["verbose"]:
        cmd += ["-v"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=env)
    got = re.sub(r"Algorithm: [0-9.]+s\.", "Algorithm: <t>s.", out.stdout)
    # issue_warnings comes first (main, qc_shor.c:1305; compared in test_host_classical.py)
    got = "".join(line for line in got.splitlines(keepends=True) if "*WARNING*" not in line)
    return out.returncode, got, out.stderr


def check_all(env=None):
    assert os.path.exists(BIN), "build the host driver with `make host`"
    cases = load_golden("shor_stdout.json")["cases"]
    assert len(cases) >= 30
    for case in cases:
        rc, got, stderr = run_case(case, env)
        assert rc == case["error"], (case, stderr)                      # the process exit code is the ErrorCode
        assert got == expected_stdout(case), (case, stderr)


def test_host_driver_stdout_is_the_references_cpu(mock_so):
    env = dict(os.environ, LD_PRELOAD=mock_so)
    check_all(env)


@pytest.mark.gpu
def test_host_driver_stdout_is_the_references_gpu():
    check_all()


@pytest.mark.parametrize("args,factors", [
    (["-C", "35", "-L", "11", "-M", "6", "-a", "2"], {5, 7}),        # period 12
    (["-C", "33", "-L", "11", "-M", "6"], {3, 11}),                  # loop: a = 2 (period 10, 2^5 = -1: rejected), a = 3 ...
    (["-C", "55", "-L", "12", "-M", "6", "-a", "2"], {5, 11}),       # period 20
    (["-C", "77", "-L", "13", "-M", "7", "-a", "2"], {7, 11}),       # period 30
])
def test_robust_mode_factors_numbers_whose_period_needs_the_continued_fraction(mock_so, args, factors):
    """with -r (modular a^p, terminated expansion, initialised flags) the host driver factors
    numbers whose period is > 10, i.e. where the period really has to come out of the measured index through
    read_omega and the continued fraction (the reference's verbatim arithmetic wraps INT_POW there, Appendix B
    #3) -- end to end on the CPU over the mock ABI, three seeds each."""
    env = dict(os.environ, LD_PRELOAD=mock_so)
    for seed in ("1", "2", "3"):
        out = subprocess.run([BIN] + args + ["-s", seed, "-r"], capture_output=True, text=True, timeout=300, env=env)
        m = re.search(r"Factors of (\d+) found: \((\d+), (\d+)\)", out.stdout)
        assert out.returncode == 0 and m, (args, seed, out.stdout, out.stderr)
        assert {int(m.group(2)), int(m.group(3))} == factors and "incorrect" not in out.stdout, (args, seed, out.stdout)
