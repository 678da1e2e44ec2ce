"""The drop-in proof of INTEGRATION.md, executed: the reference PROGRAM (its own CLI, shors_algorithm,
find_period, read_omega, continued fractions -- and, in the first variant, its own quantum_computation and
inverse_QFT) with only the edits INTEGRATION.md section 1 lists, linked against libqcs.so
(oracle/make_dropin.py builds it from the reference where it lies; binaries in oracle/_ref/), run side by side
with the unmodified reference program: same stdout, same stderr, same exit code.

The reference seeds its RNG with time(NULL) (qc_shor.c:1299), so the cases are the ones whose output does not
depend on the draw: a period <= 10 is found from denominators[0] = 1 whatever is measured.

CPU: libqcs.so's entry points are answered by tests/mock/mock_qcs.c over the CPU oracle (LD_PRELOAD; test
infrastructure -- the product has no CPU path).  GPU: the real library on the device."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_BIN = os.path.join(REF_DIR, "qc_shor_ref")
STAGE_LINES = ("         - Applying Hadamard matrices.\n", "         - Applying a^x mod (C) gates.\n",
               "         - Performing inverse quantum Fourier transform.\n")

CASES = [
    ["-C", "15", "-L", "3", "-M", "4", "-a", "7"],
    ["-C", "15", "-L", "3", "-M", "4", "-a", "7", "-v"],
    ["-C", "15", "-L", "3", "-M", "4", "-a", "7", "-V"],
    ["-C", "21", "-L", "5", "-M", "5", "-a", "2", "-v"],
    ["-C", "21", "-L", "5", "-M", "5", "-a", "2", "-V"],
    ["-C", "15", "-L", "3", "-M", "4", "-v"],                     # the loop over trial integers
    ["-C", "21", "-L", "4", "-M", "5", "-V"],
    ["-C", "15", "-L", "8", "-M", "4", "-a", "2", "-v"],          # no warnings: 2^L >= C^2
    ["-C", "15", "-L", "3", "-M", "4", "-a", "14", "-v"],         # period 2, a = C - 1: rejected, exit code 3
    ["-C", "21", "-L", "5", "-M", "5", "-a", "20"],
    ["-C", "21", "-L", "5", "-M", "5", "-a", "4", "-V"],          # odd period (and an A = 0 gate on the way)
    ["-L", "3", "-M", "4"],                                       # BAD_ARGUMENTS, qc_shor.c:1222-1238
    ["-C", "15", "-M", "4"],
    ["-C", "15", "-L", "3"],
    ["-C", "15", "-L", "3", "-M", "4", "-z"],                     # getopt's own message, then the usage line
]


def run(binary, args, env=None):
    out = subprocess.run([binary] + args, capture_output=True, text=True, timeout=300, env=env)
    stdout = re.sub(r"Algorithm: [0-9.]+s\.", "Algorithm: <t>s.", out.stdout)
    stderr = out.stderr.replace(os.path.basename(binary), "qc_shor")      # getopt names argv[0]
    stderr = stderr.replace(binary, "qc_shor")
    return out.returncode, stdout, stderr


def side_by_side(env=None):
    for name in ("qc_shor_ref", "qc_shor_dropin", "qc_shor_dropin_fused"):
        if not os.path.exists(os.path.join(REF_DIR, name)):
            pytest.skip("oracle/_ref/%s not built (python oracle/make_dropin.py, needs /root/reference)" % name)
    for args in CASES:
        want = run(REF_BIN, args)
        got = run(os.path.join(REF_DIR, "qc_shor_dropin"), args, env)
        assert got == want, args
        # quantum_computation forwarded as ONE call: the three stage lines it printed on the way are gone
        got = run(os.path.join(REF_DIR, "qc_shor_dropin_fused"), args, env)
        want_fused = (want[0], "".join(l for l in want[1].splitlines(keepends=True) if l not in STAGE_LINES), want[2])
        assert got == want_fused, args


def test_reference_program_over_the_abi_cpu(oracle_built):
    from test_host_stdout import MOCK_SO, MOCK_SRC
    orc_dir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(os.path.dirname(MOCK_SO), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "oracle"), "-o", MOCK_SO, MOCK_SRC,
                    "-L" + orc_dir, "-lqcsoracle", "-Wl,-rpath," + orc_dir], check=True)
    side_by_side(dict(os.environ, LD_PRELOAD=MOCK_SO))


@pytest.mark.gpu
def test_reference_program_over_the_abi_gpu():
    side_by_side()
