"""The reference's two development helpers, display_state and check_normalisation (testing_and_debug.c:7-37),
as the host library offers them on top of the C ABI (quantumcomputer_b200/host/state_debug.c): stdout against
the unmodified helpers' own (tests/golden/debug_stdout.json, via oracle/ref_bridge.c) on the states after
reset_register + quantum_computation, including the ones with 1e-17 residues and a non-bijective a^x mod C.

CPU: the ABI answered by tests/mock (LD_PRELOAD, the oracle; its norm is the sequential sum, so the total
matches to the last digit).  GPU: the real library, gate-by-gate mode (value-identical amplitudes, so the
listing is identical; the device's norm is a parallel reduction: compared to 1e-14)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, load_golden

HOSTLIB = os.path.join(ROOT, "quantumcomputer_b200", "lib", "libqcshost.so")

CHILD = r"""
import ctypes as C, sys
abi = C.CDLL(sys.argv[1])                  # libqcs.so, or the mock standing in for it
host = C.CDLL(sys.argv[2])
Cn, a, L, M, OPT_FUSION, POW_VERBATIM = (int(x) for x in sys.argv[3:9])
reg = C.c_void_p()
abi.qcs_register_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
assert abi.qcs_register_create(C.byref(reg), L, M, -1) == 0
abi.qcs_set_option.argtypes = [C.c_void_p, C.c_int, C.c_longlong]
assert abi.qcs_set_option(reg, OPT_FUSION, 0) == 0          # reference-order kernels
abi.qcs_reset_register.argtypes = [C.c_void_p]
abi.qcs_quantum_computation.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int]
assert abi.qcs_reset_register(reg) == 0 and abi.qcs_quantum_computation(reg, Cn, a, POW_VERBATIM) == 0
host.qcsh_display_state.argtypes = [C.c_void_p]
host.qcsh_check_normalisation.argtypes = [C.c_void_p]
assert host.qcsh_display_state(reg) == 0 and host.qcsh_check_normalisation(reg) == 0
abi.qcs_register_destroy.argtypes = [C.c_void_p]
abi.qcs_register_destroy(reg)
"""


def helper_stdout(abi_so, case, env=None):
    import quantumcomputer_b200 as q                 # the constants only; the child talks to the C ABI itself
    out = subprocess.run([sys.executable, "-c", CHILD, abi_so, HOSTLIB] + [str(case[k]) for k in ("C", "a", "L", "M")] +
                         [str(q.OPT_FUSION), str(q.POW_VERBATIM)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr
    return out.stdout


def test_debug_helpers_print_what_the_references_print_cpu(oracle_built):
    from test_host_stdout import MOCK_SO, MOCK_SRC
    orc_dir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(os.path.dirname(MOCK_SO), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "oracle"), "-o", MOCK_SO, MOCK_SRC,
                    "-L" + orc_dir, "-lqcsoracle", "-Wl,-rpath," + orc_dir, "-lm"], check=True)
    env = dict(os.environ, LD_PRELOAD=MOCK_SO)
    for case in load_golden("debug_stdout.json")["cases"]:
        assert helper_stdout(MOCK_SO, case, env) == case["stdout"], case["C"]


@pytest.mark.gpu
def test_debug_helpers_print_what_the_references_print_gpu():
    abi = os.path.join(ROOT, "quantumcomputer_b200", "lib", "libqcs.so")
    for case in load_golden("debug_stdout.json")["cases"]:
        got, want = helper_stdout(abi, case).splitlines(), case["stdout"].splitlines()
        assert got[:-1] == want[:-1], case["C"]                          # display_state: line for line
        label, _, total = got[-1].rpartition(" ")
        assert label == "Total Probability:" and abs(float(total) - float(want[-1].rpartition(" ")[2])) <= 1e-14
