"""quantumcomputer_b200 -- B200 (sm_100a) state-vector engine for the gate path
of adamalderton/QuantumComputer's ``qc_shor.c``.

The product is ``lib/libqcs.so`` (hand-written CUDA behind the C ABI of
``include/qcs.h``) and the C host driver ``bin/qc_shor_b200``.  This Python
module is a thin ctypes mirror of that ABI used by the tests and by
``bench.py``; method names and argument order are the reference's
(``reset_register``, ``hadamard_gate``, ``c_phase_shift_gate``, ``c_amodc_gate``,
``inverse_QFT``, ``quantum_computation``, ``measure_state``; qc_shor.c:272-737).
"""
import ctypes as C

import numpy as np

from . import _lib

NO_ERROR, INSUFFICIENT_MEMORY, BAD_ARGUMENTS, PERIOD_NOT_FOUND, UNKNOWN_ERROR = range(5)
POW_VERBATIM, POW_MODULAR = 0, 1
OPT_FUSION, OPT_PROFILE, OPT_TILE_BITS, OPT_MEASURE_SEQUENTIAL, OPT_PIPELINE, OPT_PREFETCH_TILES = 1, 2, 3, 4, 5, 6
OPT_PIPE_SHAPE, OPT_MIN_RUN_BITS, OPT_GLOBAL_RUN_BITS, OPT_OVERLAP_SLICES, OPT_GLOBAL_SMS = 7, 8, 9, 10, 11
OPT_L2_PAIR, OPT_L2_PAIR_LAG, OPT_L2_PAIR_MAX_BLOCK, OPT_L2_PAIR_HINTS = 12, 13, 14, 15
OPT_SPLIT3 = 17
OPT_GEN_SWEEP = 18
KERNEL_CLASSES = ["hadamard", "cphase", "amodc", "fill", "reduce", "tile_sweep",
                  "modexp_sweep", "exchange", "scale", "dense_block", "diag_multi", "global_sweep", "gate_1q"]


class QcsError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        name = _lib.load().qcs_error_string(code).decode()
        super().__init__(f"{where}: {name} ({code})")


def _check(code, where):
    if code != NO_ERROR:
        raise QcsError(code, where)


def lib():
    return _lib.load()


def device_count():
    return int(lib().qcs_device_count())


def int_pow(base, power):
    """INT_POW of qc_shor.c:158-159 as it behaves on x86-64."""
    return int(lib().qcs_int_pow(base, power))


def modpow2k(a, k, Cn):
    return int(lib().qcs_modpow2k(a, k, Cn))


def schedule_describe(n_qubits, gates, world_size=1, rank=0):
    """Passes the gate-stream scheduler would launch for `gates` (the tuples of
    quantumcomputer_b200.workloads: ("h", q) / ("cp", c, q, theta)); host only, no GPU needed.
    Returns a list of dicts: {"type": "sweep"|"hadamard"|"global"|"before", "h": [qubits],
    "diag": [stream positions], "after": [stream positions]}."""
    kinds = np.array([0 if g[0] == "h" else 1 for g in gates], dtype=np.int32)
    q0 = np.array([g[1] for g in gates], dtype=np.uint32)
    q1 = np.array([g[2] if g[0] != "h" else g[1] for g in gates], dtype=np.uint32)
    cap = 1 << 16
    while True:
        buf = C.create_string_buffer(cap)
        rc = lib().qcs_schedule_describe(n_qubits, world_size, rank, len(gates), kinds.ctypes.data,
                                         q0.ctypes.data, q1.ctypes.data, buf, cap)
        if rc == INSUFFICIENT_MEMORY:
            cap *= 4
            continue
        _check(rc, "qcs_schedule_describe")
        break
    passes = []
    for line in buf.value.decode().splitlines():
        fields = line.split()
        entry = {"type": fields[0].split("=")[0], "h": [], "diag": [], "after": []}
        for f in fields:
            key, _, val = f.partition("=")
            if key == "h":
                mask = int(val, 16)
                entry["h"] = [b for b in range(64) if (mask >> b) & 1]
            elif key == "q":
                entry["h"] = [int(val)]
            elif key in ("diag", "after"):
                entry[key] = [int(v) for v in val.split(",")]
            elif key == "before":
                entry["after"] = [int(v) for v in val.split(",")]
        passes.append(entry)
    return passes


def plan_describe(n_qubits, lo=0, hi=None, tile_bits=0, min_run_bits=0):
    """Sweeps of the fused inverse transform on qubits [lo, hi); host only, no GPU needed.
    Returns a list of dicts {"a", "g_lo", "g_hi", "tiles", "steps": [(lowest bit, radix bits)], "scale"}."""
    hi = n_qubits if hi is None else hi
    buf = C.create_string_buffer(1 << 14)
    _check(lib().qcs_plan_describe(n_qubits, lo, hi, tile_bits, min_run_bits, buf, len(buf)), "qcs_plan_describe")
    sweeps = []
    for line in buf.value.decode().splitlines():
        f = dict(kv.split("=", 1) for kv in line.split()[1:])
        g_lo, g_hi = f["g"].strip("[)").split(",")
        sweeps.append({"a": int(f["a"]), "g_lo": int(g_lo), "g_hi": int(g_hi), "tiles": int(f["tiles"]),
                       "steps": [tuple(int(v) for v in st.split("+")) for st in f["steps"].split(",")],
                       "scale": float(f["scale"])})
    return sweeps


def comm_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().qcs_comm_unique_id(buf), "qcs_comm_unique_id")
    return buf.raw


class PinnedBuffer:
    """Page-locked host memory viewed as a numpy float64 array."""

    def __init__(self, n_doubles, device=-1):
        """device >= 0: pages on the NUMA node next to that GPU (qcs_host_alloc_near)."""
        self._ptr = C.c_void_p()
        if device >= 0:
            _check(lib().qcs_host_alloc_near(C.byref(self._ptr), n_doubles * 8, device), "qcs_host_alloc_near")
        else:
            _check(lib().qcs_host_alloc(C.byref(self._ptr), n_doubles * 8), "qcs_host_alloc")
        arr_t = C.c_double * n_doubles
        self.array = np.frombuffer(arr_t.from_address(self._ptr.value), dtype=np.float64)

    def close(self):
        if self._ptr:
            self.array = None
            lib().qcs_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Register:
    """Device register: the replacement of the reference's ``Register`` struct
    (qc_shor.c:194-203) plus the operators that act on it."""

    def __init__(self, L_size, M_size, device=-1, rank=0, world_size=1, comm_id=None, n_gpus=1):
        """world_size > 1: this process's shard of a sharded register (one process per GPU);
        n_gpus > 1: the whole register sharded over the first n_gpus devices, driven from this
        one process (qcs_register_create_multi)."""
        self._l = lib()
        self._h = C.c_void_p()
        if n_gpus > 1:
            rc = self._l.qcs_register_create_multi(C.byref(self._h), L_size, M_size, n_gpus)
        elif world_size == 1:
            rc = self._l.qcs_register_create(C.byref(self._h), L_size, M_size, device)
        else:
            rc = self._l.qcs_register_create_sharded(C.byref(self._h), L_size, M_size, device,
                                                     rank, world_size, comm_id)
        _check(rc, "qcs_register_create")

    def close(self):
        if getattr(self, "_h", None):
            self._l.qcs_register_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- info
    @property
    def L_size(self):
        return self._l.qcs_L_size(self._h)

    @property
    def M_size(self):
        return self._l.qcs_M_size(self._h)

    @property
    def num_qubits(self):
        return self._l.qcs_num_qubits(self._h)

    @property
    def num_states(self):
        return self._l.qcs_num_states(self._h)

    @property
    def local_states(self):
        return self._l.qcs_local_states(self._h)

    @property
    def num_gpus(self):
        return self._l.qcs_num_gpus(self._h)

    @property
    def peer_memory(self):
        """True when the shards of a sharded register are stitched into one address range."""
        return bool(self._l.qcs_peer_memory(self._h))

    def set_option(self, opt, value):
        _check(self._l.qcs_set_option(self._h, opt, value), "qcs_set_option")

    def get_option(self, opt):
        return self._l.qcs_get_option(self._h, opt)

    def synchronize(self):
        _check(self._l.qcs_synchronize(self._h), "qcs_synchronize")

    # ---- gate path (reference names and argument order)
    def reset_register(self):
        _check(self._l.qcs_reset_register(self._h), "reset_register")

    def hadamard_gate(self, qubit_num):
        _check(self._l.qcs_hadamard_gate(self._h, qubit_num), "hadamard_gate")

    def c_phase_shift_gate(self, c_qubit_num, qubit_num, theta):
        _check(self._l.qcs_c_phase_shift_gate(self._h, c_qubit_num, qubit_num, theta),
               "c_phase_shift_gate")

    def c_amodc_gate(self, Cn, atox, c_qubit_num):
        _check(self._l.qcs_c_amodc_gate(self._h, Cn, atox, c_qubit_num), "c_amodc_gate")

    def inverse_QFT(self, lo=None, hi=None):
        if lo is None:
            _check(self._l.qcs_inverse_QFT(self._h), "inverse_QFT")
        else:
            _check(self._l.qcs_inverse_QFT_range(self._h, lo, hi), "inverse_QFT_range")

    def QFT(self, lo=None, hi=None):
        if lo is None:
            _check(self._l.qcs_QFT(self._h), "QFT")
        else:
            _check(self._l.qcs_QFT_range(self._h, lo, hi), "QFT_range")

    def quantum_computation(self, Cn, a, pow_mode=POW_VERBATIM):
        _check(self._l.qcs_quantum_computation(self._h, Cn, a, pow_mode), "quantum_computation")

    def measure_state(self, r):
        out = C.c_ulonglong(0)
        _check(self._l.qcs_measure_state(self._h, r, C.byref(out)), "measure_state")
        return int(out.value)

    def sample_states(self, rs):
        """Indices measure_state would return for each variate in rs, without collapsing."""
        r = np.ascontiguousarray(np.asarray(rs, dtype=np.float64))
        out = np.zeros(r.size, dtype=np.uint64)
        _check(self._l.qcs_sample_states(self._h, r.size, r.ctypes.data, out.ctypes.data), "sample_states")
        return [int(v) for v in out]

    def apply_gate(self, qubit_num, U):
        u = np.ascontiguousarray(np.asarray(U, dtype=np.complex128).reshape(2, 2)).view(np.float64)
        _check(self._l.qcs_apply_gate(self._h, qubit_num, u.ctypes.data), "apply_gate")

    def apply_controlled_gate(self, c_qubit_num, qubit_num, U):
        u = np.ascontiguousarray(np.asarray(U, dtype=np.complex128).reshape(2, 2)).view(np.float64)
        _check(self._l.qcs_apply_controlled_gate(self._h, c_qubit_num, qubit_num, u.ctypes.data),
               "apply_controlled_gate")

    def save_state(self, path):
        _check(self._l.qcs_save_state(self._h, str(path).encode()), "save_state")

    def load_state(self, path):
        _check(self._l.qcs_load_state(self._h, str(path).encode()), "load_state")

    def norm2(self):
        out = C.c_double(0.0)
        _check(self._l.qcs_norm2(self._h, C.byref(out)), "norm2")
        return float(out.value)

    def nonzero_states(self, capacity=1 << 16):
        idx = (C.c_ulonglong * capacity)()
        mag = (C.c_double * capacity)()
        cnt = C.c_ulonglong(0)
        _check(self._l.qcs_nonzero_states(self._h, capacity, idx, mag, C.byref(cnt)), "nonzero_states")
        k = min(capacity, cnt.value)
        return list(idx[:k]), list(mag[:k]), int(cnt.value)

    # ---- bulk access
    def get_state(self, first=0, count=None, out=None):
        if count is None:
            count = self.local_states - first
        if out is None:
            out = np.empty(2 * count, dtype=np.float64)
        assert out.dtype == np.float64 and out.size >= 2 * count and out.flags.c_contiguous
        _check(self._l.qcs_get_state(self._h, first, count, out.ctypes.data), "get_state")
        return out[:2 * count].view(np.complex128)

    def set_state(self, amps, first=0):
        a = np.asarray(amps)
        if a.dtype != np.float64:
            a = np.ascontiguousarray(a, dtype=np.complex128).view(np.float64)
        a = np.ascontiguousarray(a)
        _check(self._l.qcs_set_state(self._h, first, a.size // 2, a.ctypes.data), "set_state")

    def set_state_async(self, float64_array, first=0):
        """Host -> device copy without the trailing synchronise (pinned source)."""
        _check(self._l.qcs_set_state_async(self._h, first, float64_array.size // 2,
                                           float64_array.ctypes.data), "set_state_async")

    # ---- deferred gate stream
    def fuse_begin(self):
        _check(self._l.qcs_fuse_begin(self._h), "fuse_begin")

    def fuse_end(self):
        _check(self._l.qcs_fuse_end(self._h), "fuse_end")

    @property
    def fuse_pending(self):
        return int(self._l.qcs_fuse_pending(self._h))

    def fused(self):
        """``with reg.fused(): ...`` records H / C-phase gates and launches them fused on exit."""
        reg = self

        class _Ctx:
            def __enter__(self):
                reg.fuse_begin()
                return reg

            def __exit__(self, *exc):
                reg.fuse_end()
        return _Ctx()

    def apply_dense_block(self, k, U):
        """U: (2^k, 2^k) complex matrix applied to qubits 0..k-1 (k = 3 or 4), DMMA kernel."""
        u = np.ascontiguousarray(np.asarray(U, dtype=np.complex128)).view(np.float64)
        assert u.size == 2 * (1 << k) ** 2
        _check(self._l.qcs_apply_dense_block(self._h, k, u.ctypes.data), "apply_dense_block")

    def fill_synthetic(self, seed):
        _check(self._l.qcs_fill_synthetic(self._h, seed), "fill_synthetic")

    def scale(self, factor):
        _check(self._l.qcs_scale(self._h, factor), "scale")

    # ---- engine measurement
    def timer_start(self):
        _check(self._l.qcs_timer_start(self._h), "timer_start")

    def timer_stop(self):
        ms = C.c_double(0.0)
        _check(self._l.qcs_timer_stop(self._h, C.byref(ms)), "timer_stop")
        return float(ms.value)

    @property
    def launch_count(self):
        return int(self._l.qcs_launch_count(self._h))

    def profile_reset(self):
        _check(self._l.qcs_profile_reset(self._h), "profile_reset")

    def profile(self):
        """{class name: (launches, milliseconds, algorithmic bytes)}"""
        out = {}
        for k, name in enumerate(KERNEL_CLASSES):
            n = C.c_ulonglong(0)
            ms = C.c_double(0.0)
            by = C.c_double(0.0)
            _check(self._l.qcs_profile_get(self._h, k, C.byref(n), C.byref(ms), C.byref(by)),
                   "profile_get")
            out[name] = (int(n.value), float(ms.value), float(by.value))
        return out
