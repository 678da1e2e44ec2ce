"""GPU: SURVEY 8(d) cfg1 / cfg2 across seeds -- the C host driver's find_period and
shors_algorithm (quantumcomputer_b200/host, the three calls of qc_shor.c:922-928 made into
libqcs.so) against the oracle's restatement of the same functions with the same MT19937 seed:
error code, period, measured index and factors must be identical for every seed."""
import ctypes as C
import os

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


class Options(C.Structure):
    _fields_ = [("mode", C.c_int), ("verbose", C.c_int), ("very_verbose", C.c_int),
                ("last_measured", C.c_ulonglong), ("last_omega", C.c_double)]


@pytest.fixture(scope="module")
def host(qcs):
    lib = C.CDLL(os.path.join(ROOT, "quantumcomputer_b200", "lib", "libqcshost.so"))
    lib.qcsh_rng_seed.argtypes = [C.c_void_p, C.c_ulong]
    lib.qcsh_find_period.restype = C.c_int
    lib.qcsh_find_period.argtypes = [C.POINTER(C.c_uint), C.c_uint, C.c_uint, C.c_void_p, C.c_void_p,
                                     C.POINTER(Options)]
    lib.qcsh_shors_algorithm.restype = C.c_int
    lib.qcsh_shors_algorithm.argtypes = [C.POINTER(C.c_uint), C.c_uint, C.c_uint, C.c_void_p, C.c_void_p,
                                         C.POINTER(Options)]
    return lib


def new_rng(host, seed):
    buf = C.create_string_buffer(624 * 4 + 16)
    host.qcsh_rng_seed(buf, seed)
    return buf


@pytest.mark.parametrize("Cn,a,L,M,seeds,fusion", [
    (15, 7, 3, 4, [12345] + list(range(1, 101)), 1),
    (15, 7, 3, 4, list(range(1, 26)), 0),
    (21, 2, 5, 5, list(range(1, 41)), 1),
    (15, 2, 3, 4, list(range(1, 21)), 1),
])
def test_find_period_identical_to_oracle_across_seeds(qcs, oracle_built, host, Cn, a, L, M, seeds, fusion):
    o = oracle_built.Restatement(L, M)
    with qcs.Register(L, M) as reg:
        reg.set_option(qcs.OPT_FUSION, fusion)
        for seed in seeds:
            want_err, want_period, want_measured = o.find_period(Cn, a, o.rng(seed), 0)
            opt = Options(0, 0, 0, 0, 0.0)
            period = C.c_uint(0)
            err = host.qcsh_find_period(C.byref(period), Cn, a, reg._h, new_rng(host, seed), C.byref(opt))
            assert err == want_err, (seed, err, want_err)
            assert opt.last_measured == want_measured, (seed, opt.last_measured, want_measured)
            if err == 0:
                assert period.value == want_period, (seed, period.value, want_period)


@pytest.mark.parametrize("Cn,forced_a,L,M,seeds", [(15, 7, 3, 4, list(range(1, 31))), (21, 2, 5, 5, list(range(1, 16))),
                                                   (15, 0, 3, 4, list(range(1, 21)))])
def test_shors_algorithm_identical_to_oracle_across_seeds(qcs, oracle_built, host, Cn, forced_a, L, M, seeds):
    """forced_a = 0: the trial integers are drawn from the same stream as the measurement variates."""
    o = oracle_built.Restatement(L, M)
    with qcs.Register(L, M) as reg:
        for seed in seeds:
            want_err, want_factors = o.shors_algorithm(Cn, forced_a, o.rng(seed), 0)
            opt = Options(0, 0, 0, 0, 0.0)
            factors = (C.c_uint * 2)(0, 0)
            err = host.qcsh_shors_algorithm(factors, Cn, forced_a, reg._h, new_rng(host, seed), C.byref(opt))
            assert err == want_err, (seed, err, want_err)
            if err == 0:
                assert (factors[0], factors[1]) == want_factors, (seed, tuple(factors), want_factors)
