"""ctypes bindings for the two CPU checkers (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "libqcsoracle.so")
_ORACLE_OMP_SO = os.path.join(_HERE, "_build", "libqcsoracle_omp.so")
_REF_SO = os.path.join(_HERE, "_ref", "libqcref.so")

_u = C.c_uint
_ull = C.c_ulonglong
_dp = C.POINTER(C.c_double)


def build():
    """(Re)build both checkers with oracle/Makefile."""
    subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True,
                   stdout=subprocess.DEVNULL)


def have_restatement():
    return os.path.exists(_ORACLE_SO)


def have_reference():
    return os.path.exists(_REF_SO)


def have_restatement_omp():
    return os.path.exists(_ORACLE_OMP_SO)


def _as_dp(a):
    return a.ctypes.data_as(_dp)


class _MT(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int)]


class MT19937:
    """gsl_rng_mt19937 semantics (seed 0 -> 4357, uniform = u32 / 2^32)."""

    def __init__(self, lib, seed):
        self._lib = lib
        self._g = _MT()
        lib.orc_mt_seed(C.byref(self._g), C.c_ulong(seed))

    def next_u32(self):
        return int(self._lib.orc_mt_next(C.byref(self._g)))

    def uniform(self):
        return float(self._lib.orc_mt_uniform(C.byref(self._g)))


def _load_restatement(path=None):
    lib = C.CDLL(path or _ORACLE_SO)
    vp = C.c_void_p
    lib.orc_create.restype = vp
    lib.orc_create.argtypes = [C.c_int, C.c_int]
    lib.orc_destroy.argtypes = [vp]
    lib.orc_num_states.restype = C.c_uint64
    lib.orc_num_states.argtypes = [vp]
    lib.orc_get_state.argtypes = [vp, _dp]
    lib.orc_set_state.argtypes = [vp, _dp]
    lib.orc_reset_register.argtypes = [vp]
    lib.orc_hadamard_gate.argtypes = [vp, _u]
    lib.orc_c_phase_shift_gate.argtypes = [vp, _u, _u, C.c_double]
    lib.orc_c_amodc_gate.argtypes = [vp, _u, _ull, _u]
    lib.orc_inverse_QFT.argtypes = [vp]
    lib.orc_quantum_computation.argtypes = [vp, _u, _u, C.c_int]
    lib.orc_measure_state.restype = C.c_uint64
    lib.orc_measure_state.argtypes = [vp, C.c_double]
    lib.orc_norm2.restype = C.c_double
    lib.orc_norm2.argtypes = [vp]
    lib.orc_int_pow.restype = _u
    lib.orc_int_pow.argtypes = [_u, _u]
    lib.orc_gcd.restype = _u
    lib.orc_gcd.argtypes = [_u, _u]
    lib.orc_read_omega.restype = C.c_double
    lib.orc_read_omega.argtypes = [C.c_uint64, C.c_int, C.c_int]
    lib.orc_continued_fraction_denominators.argtypes = [C.c_double, _u, C.POINTER(_u)]
    lib.orc_mt_seed.argtypes = [C.POINTER(_MT), C.c_ulong]
    lib.orc_mt_next.restype = C.c_uint32
    lib.orc_mt_next.argtypes = [C.POINTER(_MT)]
    lib.orc_mt_uniform.restype = C.c_double
    lib.orc_mt_uniform.argtypes = [C.POINTER(_MT)]
    lib.orc_find_period.restype = C.c_int
    lib.orc_find_period.argtypes = [vp, _u, _u, C.c_int, C.POINTER(_MT), C.POINTER(_u),
                                    C.POINTER(C.c_uint64)]
    lib.orc_shors_algorithm.restype = C.c_int
    lib.orc_shors_algorithm.argtypes = [vp, _u, _u, C.c_int, C.POINTER(_MT), C.POINTER(_u)]
    lib.orc_synthetic_u.restype = C.c_double
    lib.orc_synthetic_u.argtypes = [C.c_uint64, C.c_uint64]
    lib.orc_fill_synthetic.argtypes = [vp, C.c_uint64]
    lib.orc_scale.argtypes = [vp, C.c_double]
    lib.orc_threads.restype = C.c_int
    return lib


class Restatement:
    """The matrix-free CPU restatement (qcs_oracle.c)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            cls._lib = _load_restatement()
        return cls._lib

    @classmethod
    def threads(cls):
        return int(cls.lib().orc_threads())

    def __init__(self, L_size, M_size):
        self.L, self.M = L_size, M_size
        self.n = L_size + M_size
        self._h = self.lib().orc_create(L_size, M_size)
        if not self._h:
            raise MemoryError("orc_create failed")

    def close(self):
        if self._h:
            self.lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_states(self):
        return 1 << self.n

    def get_state(self):
        out = np.empty(2 * self.num_states, dtype=np.float64)
        self.lib().orc_get_state(self._h, _as_dp(out))
        return out.view(np.complex128)

    def set_state(self, amps):
        a = np.ascontiguousarray(np.asarray(amps, dtype=np.complex128)).view(np.float64)
        assert a.size == 2 * self.num_states
        self.lib().orc_set_state(self._h, _as_dp(a))

    def reset_register(self):
        self.lib().orc_reset_register(self._h)

    def hadamard_gate(self, q):
        self.lib().orc_hadamard_gate(self._h, q)

    def c_phase_shift_gate(self, c, q, theta):
        self.lib().orc_c_phase_shift_gate(self._h, c, q, theta)

    def c_amodc_gate(self, Cn, atox, c):
        self.lib().orc_c_amodc_gate(self._h, Cn, atox, c)

    def inverse_QFT(self):
        self.lib().orc_inverse_QFT(self._h)

    def quantum_computation(self, Cn, a, pow_mode=0):
        self.lib().orc_quantum_computation(self._h, Cn, a, pow_mode)

    def measure_state(self, r):
        return int(self.lib().orc_measure_state(self._h, r))

    def norm2(self):
        return float(self.lib().orc_norm2(self._h))

    def fill_synthetic(self, seed):
        self.lib().orc_fill_synthetic(self._h, seed)

    def scale(self, s):
        self.lib().orc_scale(self._h, s)

    def rng(self, seed):
        return MT19937(self.lib(), seed)

    def find_period(self, Cn, a, rng, pow_mode=0):
        p = _u(0)
        m = C.c_uint64(0)
        err = self.lib().orc_find_period(self._h, Cn, a, pow_mode, C.byref(rng._g),
                                         C.byref(p), C.byref(m))
        return err, int(p.value), int(m.value)

    def shors_algorithm(self, Cn, forced_a, rng, pow_mode=0):
        f = (_u * 2)(0, 0)
        err = self.lib().orc_shors_algorithm(self._h, Cn, forced_a, pow_mode,
                                             C.byref(rng._g), f)
        return err, (int(f[0]), int(f[1]))

    # scalar helpers
    @classmethod
    def int_pow(cls, b, p):
        return int(cls.lib().orc_int_pow(b, p))

    @classmethod
    def gcd(cls, a, b):
        return int(cls.lib().orc_gcd(a, b))

    @classmethod
    def read_omega(cls, state, L, M):
        return float(cls.lib().orc_read_omega(state, L, M))

    @classmethod
    def cf_denominators(cls, omega, n=15):
        out = (_u * n)()
        cls.lib().orc_continued_fraction_denominators(omega, n, out)
        return [int(x) for x in out]

    @classmethod
    def synthetic_u(cls, seed, k):
        return float(cls.lib().orc_synthetic_u(seed, k))


class RestatementAllCores(Restatement):
    """The same restatement compiled with -fopenmp: the row loops of hadamard_gate and
    c_phase_shift_gate run on all host cores (bit-identical results; bench.py's cpu_baseline)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            cls._lib = _load_restatement(_ORACLE_OMP_SO)
        return cls._lib


def _load_reference():
    lib = C.CDLL(_REF_SO)
    vp = C.c_void_p
    lib.qcref_create.restype = vp
    lib.qcref_create.argtypes = [C.c_int, C.c_int]
    lib.qcref_destroy.argtypes = [vp]
    lib.qcref_num_states.restype = _ull
    lib.qcref_num_states.argtypes = [vp]
    lib.qcref_set_verbosity.argtypes = [C.c_int, C.c_int]
    lib.qcref_seed.argtypes = [vp, C.c_ulong]
    lib.qcref_rng_uniform.restype = C.c_double
    lib.qcref_rng_uniform.argtypes = [vp]
    lib.qcref_get_state.argtypes = [vp, _dp]
    lib.qcref_set_state.argtypes = [vp, _dp]
    lib.qcref_reset_register.argtypes = [vp]
    lib.qcref_hadamard_gate.argtypes = [vp, _u]
    lib.qcref_c_phase_shift_gate.argtypes = [vp, _u, _u, C.c_double]
    lib.qcref_c_amodc_gate.argtypes = [vp, _u, _ull, _u]
    lib.qcref_inverse_QFT.argtypes = [vp]
    lib.qcref_quantum_computation.argtypes = [vp, _u, _u]
    lib.qcref_int_pow.restype = _u
    lib.qcref_int_pow.argtypes = [_u, _u]
    lib.qcref_measure_state.restype = _ull
    lib.qcref_measure_state.argtypes = [vp]
    lib.qcref_measure_state_r.restype = _ull
    lib.qcref_measure_state_r.argtypes = [vp, C.c_double]
    lib.qcref_read_omega.restype = C.c_double
    lib.qcref_read_omega.argtypes = [vp, _ull]
    lib.qcref_continued_fraction_denominators.argtypes = [C.c_double, _u, C.POINTER(_u)]
    lib.qcref_gcd.restype = _u
    lib.qcref_gcd.argtypes = [_u, _u]
    lib.qcref_find_period.restype = C.c_int
    lib.qcref_find_period.argtypes = [vp, _u, _u, C.POINTER(_u)]
    lib.qcref_shors_algorithm.restype = C.c_int
    lib.qcref_shors_algorithm.argtypes = [vp, _u, _u, C.POINTER(_u)]
    lib.qcref_norm2.restype = C.c_double
    lib.qcref_norm2.argtypes = [vp]
    lib.qcref_issue_warnings.argtypes = [_u, C.c_int, C.c_int]
    return lib


class Reference:
    """The unmodified reference program, compiled in place (ref_bridge.c)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            cls._lib = _load_reference()
        return cls._lib

    def __init__(self, L_size, M_size):
        self.L, self.M = L_size, M_size
        self.n = L_size + M_size
        self._h = self.lib().qcref_create(L_size, M_size)
        if not self._h:
            raise MemoryError("qcref_create failed")

    def close(self):
        if self._h:
            self.lib().qcref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_states(self):
        return 1 << self.n

    def seed(self, s):
        self.lib().qcref_seed(self._h, s)

    def rng_uniform(self):
        return float(self.lib().qcref_rng_uniform(self._h))

    def get_state(self):
        out = np.empty(2 * self.num_states, dtype=np.float64)
        self.lib().qcref_get_state(self._h, _as_dp(out))
        return out.view(np.complex128)

    def set_state(self, amps):
        a = np.ascontiguousarray(np.asarray(amps, dtype=np.complex128)).view(np.float64)
        assert a.size == 2 * self.num_states
        self.lib().qcref_set_state(self._h, _as_dp(a))

    def reset_register(self):
        self.lib().qcref_reset_register(self._h)

    def hadamard_gate(self, q):
        self.lib().qcref_hadamard_gate(self._h, q)

    def c_phase_shift_gate(self, c, q, theta):
        self.lib().qcref_c_phase_shift_gate(self._h, c, q, theta)

    def c_amodc_gate(self, Cn, atox, c):
        self.lib().qcref_c_amodc_gate(self._h, Cn, atox, c)

    def inverse_QFT(self):
        self.lib().qcref_inverse_QFT(self._h)

    def quantum_computation(self, Cn, a):
        self.lib().qcref_quantum_computation(self._h, Cn, a)

    def measure_state(self):
        return int(self.lib().qcref_measure_state(self._h))

    def measure_state_r(self, r):
        return int(self.lib().qcref_measure_state_r(self._h, r))

    def norm2(self):
        return float(self.lib().qcref_norm2(self._h))

    def find_period(self, Cn, a):
        p = _u(0)
        err = self.lib().qcref_find_period(self._h, Cn, a, C.byref(p))
        return err, int(p.value)

    def shors_algorithm(self, Cn, forced_a):
        f = (_u * 2)(0, 0)
        err = self.lib().qcref_shors_algorithm(self._h, Cn, forced_a, f)
        return err, (int(f[0]), int(f[1]))

    @staticmethod
    def shor_stdout(Cn, L_size, M_size, forced_a, seed, verbose=0, very_verbose=0):
        """(ErrorCode, factors, stdout) of the reference's shors_algorithm (qc_shor.c:1003-1134) with
        gsl_rng_set(seed) and the -v / -V flags, run in a child process (the program prints as it goes)."""
        import json
        import sys
        code = ("import sys, json; sys.path.insert(0, %r); from oracle.bindings import Reference; "
                "ref = Reference(%d, %d); Reference.lib().qcref_set_verbosity(%d, %d); ref.seed(%d); "
                "err, f = ref.shors_algorithm(%d, %d); sys.stderr.write(json.dumps([err, list(f)]))"
                % (os.path.dirname(_HERE), L_size, M_size, verbose, very_verbose, seed, Cn, forced_a))
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True, timeout=600)
        err, f = json.loads(out.stderr.strip().splitlines()[-1])
        return err, f, out.stdout

    @staticmethod
    def debug_stdout(Cn, a, L_size, M_size):
        """stdout of display_state, then of check_normalisation (testing_and_debug.c:7-37), on the state after
        reset_register + quantum_computation(C, a); child process, the helpers print."""
        import sys
        code = ("import sys; sys.path.insert(0, %r); from oracle.bindings import Reference; "
                "ref = Reference(%d, %d); ref.reset_register(); ref.quantum_computation(%d, %d); "
                "Reference.lib().qcref_display_state.argtypes = [__import__('ctypes').c_void_p]; "
                "Reference.lib().qcref_check_normalisation.argtypes = [__import__('ctypes').c_void_p]; "
                "Reference.lib().qcref_display_state(ref._h); Reference.lib().qcref_check_normalisation(ref._h)"
                % (os.path.dirname(_HERE), L_size, M_size, Cn, a))
        return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True, timeout=600).stdout

    @staticmethod
    def warnings_text(Cn, L_size, M_size):
        """What issue_warnings (qc_shor.c:340-351) prints for these sizes: run in a child process, because
        the reference writes to the C stdout."""
        import sys
        code = ("import sys; sys.path.insert(0, %r); from oracle.bindings import Reference; "
                "Reference.lib().qcref_issue_warnings(%d, %d, %d)" % (os.path.dirname(_HERE), Cn, L_size, M_size))
        return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True, timeout=120).stdout

    @classmethod
    def int_pow(cls, b, p):
        return int(cls.lib().qcref_int_pow(b, p))

    @classmethod
    def gcd(cls, a, b):
        return int(cls.lib().qcref_gcd(a, b))

    def read_omega(self, state):
        return float(self.lib().qcref_read_omega(self._h, state))

    @classmethod
    def cf_denominators(cls, omega, n=15):
        out = (_u * n)()
        cls.lib().qcref_continued_fraction_denominators(omega, n, out)
        return [int(x) for x in out]
