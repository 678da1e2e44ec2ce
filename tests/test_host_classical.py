"""CPU tests of the C host driver's classical half (quantumcomputer_b200/host),
against vectors produced by the unmodified reference (tests/golden/scalars.json)."""
import ctypes as C
import os

import pytest

from conftest import ROOT, load_golden


@pytest.fixture(scope="module")
def host(qcs):
    lib = C.CDLL(os.path.join(ROOT, "quantumcomputer_b200", "lib", "libqcshost.so"))
    lib.qcsh_gcd.restype = C.c_uint
    lib.qcsh_gcd.argtypes = [C.c_uint, C.c_uint]
    lib.qcsh_read_omega.restype = C.c_double
    lib.qcsh_read_omega.argtypes = [C.c_ulonglong, C.c_int, C.c_int]
    lib.qcsh_continued_fraction_denominators.restype = C.c_uint
    lib.qcsh_continued_fraction_denominators.argtypes = [C.c_double, C.c_uint, C.POINTER(C.c_uint), C.c_int]
    lib.qcsh_period_is_valid.restype = C.c_int
    lib.qcsh_period_is_valid.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_int]
    lib.qcsh_modpow.restype = C.c_ulonglong
    lib.qcsh_modpow.argtypes = [C.c_uint, C.c_ulonglong, C.c_uint]
    lib.qcsh_rng_seed.argtypes = [C.c_void_p, C.c_ulong]
    lib.qcsh_rng_uniform.restype = C.c_double
    lib.qcsh_rng_uniform.argtypes = [C.c_void_p]
    lib.qcsh_rng_u32.restype = C.c_uint32
    lib.qcsh_rng_u32.argtypes = [C.c_void_p]
    return lib


def new_rng(host, seed):
    buf = C.create_string_buffer(624 * 4 + 16)
    host.qcsh_rng_seed(buf, seed)
    return buf


def test_mt19937_kat3(host):
    g = new_rng(host, 5489)
    assert [host.qcsh_rng_u32(g), host.qcsh_rng_u32(g)] == [3499211612, 581869302]
    gold = load_golden("scalars.json")["mt19937"]
    for key, seed in (("seed5489_first_uniform", 5489), ("seed0_first_uniform", 0), ("seed4357_first_uniform", 4357)):
        g = new_rng(host, seed)
        assert [host.qcsh_rng_uniform(g).hex() for _ in gold[key]] == gold[key]
    g = new_rng(host, 12345)
    assert host.qcsh_rng_uniform(g) == 0.92961608665063977      # KAT-1's r


def test_workload_generator_uses_the_same_stream():
    """quantumcomputer_b200.workloads draws the layered circuit's angles from the same MT19937."""
    from quantumcomputer_b200.workloads import layered_circuit, mt19937_uniforms
    gold = load_golden("scalars.json")["mt19937"]
    for key, seed in (("seed5489_first_uniform", 5489), ("seed0_first_uniform", 0), ("seed4357_first_uniform", 4357)):
        assert [u.hex() for u in mt19937_uniforms(seed, len(gold[key]))] == gold[key]
    assert mt19937_uniforms(12345, 1)[0] == 0.92961608665063977
    assert len(mt19937_uniforms(7, 1500)) == 1500          # crosses two state regenerations
    gates = layered_circuit(33, 8)
    assert len(gates) == 528 and gates[0] == ("h", 0) and gates[33][:3] == ("cp", 0, 1)
    assert gates[66 * 7 + 33 + 32][:3] == ("cp", 32, (32 + 1 + 7) % 33)


def test_gcd_and_read_omega(host):
    g = load_golden("scalars.json")
    for e in g["gcd"]:
        assert host.qcsh_gcd(e["a"], e["b"]) == e["g"], e
    for e in g["read_omega"]:
        assert host.qcsh_read_omega(e["state"], e["L"], e["M"]).hex() == e["omega"], e


def test_continued_fractions_verbatim(host):
    for e in load_golden("scalars.json")["continued_fractions"]:
        out = (C.c_uint * 15)()
        n = host.qcsh_continued_fraction_denominators(float.fromhex(e["omega"]), 15, out, 0)
        assert n == 15 and list(out) == e["den"], e


def test_continued_fractions_robust_terminate(host):
    out = (C.c_uint * 15)()
    n = host.qcsh_continued_fraction_denominators(0.75, 15, out, 1)
    assert list(out[:n])[:3] == [1, 1, 4] and n <= 15
    n0 = host.qcsh_continued_fraction_denominators(0.0, 15, out, 1)
    assert n0 == 1 and out[0] == 1
    n = host.qcsh_continued_fraction_denominators(5.0 / 32.0, 15, out, 1)     # KAT-2's omega = 0.15625
    assert 6 in [out[i] * m for i in range(n) for m in range(1, 11)]


def test_period_validity(host):
    assert host.qcsh_period_is_valid(7, 4, 15, 0) == 1 and host.qcsh_period_is_valid(7, 4, 15, 1) == 1
    assert host.qcsh_period_is_valid(7, 2, 15, 0) == 0
    # 2^60 wraps INT_POW (verbatim says 0 % 21 != 1) but is 1 mod 21
    assert host.qcsh_period_is_valid(2, 60, 21, 0) == 0 and host.qcsh_period_is_valid(2, 60, 21, 1) == 1
    assert host.qcsh_modpow(7, 123456789, 1000003) == pow(7, 123456789, 1000003)


def test_register_size_warnings_are_the_references():
    """The host driver prints issue_warnings' text (qc_shor.c:340-351) before it touches a device: stdout
    compared with the unmodified reference's (tests/golden/warnings.json, and live when oracle/_ref exists).
    -d 9999 makes register creation fail at once on any box, so only the warnings are printed."""
    import subprocess
    import oracle
    binary = os.path.join(ROOT, "quantumcomputer_b200", "bin", "qc_shor_b200")
    assert os.path.exists(binary), "build the host driver with `make host`"
    for case in load_golden("warnings.json")["cases"]:
        out = subprocess.run([binary, "-C", str(case["C"]), "-L", str(case["L"]), "-M", str(case["M"]), "-d", "9999"],
                             capture_output=True, text=True, timeout=60)
        assert out.returncode != 0 and "Error" in out.stderr          # no device 9999: fails loudly, no CPU path
        assert out.stdout == case["stdout"], case
        if oracle.have_reference():
            assert oracle.Reference.warnings_text(case["C"], case["L"], case["M"]) == case["stdout"]
