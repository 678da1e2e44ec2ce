"""CPU tests of the drop-in boundary: libqcs.so loads, exports every symbol that
include/qcs.h declares, its host-side scalar helpers match the reference, and it
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, load_golden


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qcs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qcs_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(qcs):
    lib = qcs.lib()
    names = declared_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/qcs.h but not exported by libqcs.so"


def test_python_binding_covers_header(qcs):
    from quantumcomputer_b200 import _lib
    assert sorted(n for n, _, _ in _lib.SIGNATURES) == declared_symbols()


def test_error_strings_follow_reference_enum(qcs):
    # ErrorCode, qc_shor.c:164-170
    lib = qcs.lib()
    assert [lib.qcs_error_string(i).decode() for i in range(5)] == [
        "NO_ERROR", "INSUFFICIENT_MEMORY", "BAD_ARGUMENTS", "PERIOD_NOT_FOUND", "UNKNOWN_ERROR"]


def test_int_pow_matches_reference(qcs):
    for e in load_golden("scalars.json")["int_pow"]:
        assert qcs.int_pow(e["base"], e["power"]) == e["value"], e


def test_modpow2k(qcs):
    for a, k, Cn in [(7, 0, 15), (7, 1, 15), (7, 4, 15), (2, 9, 21), (5, 3, 33), (4, 5, 21)]:
        assert qcs.modpow2k(a, k, Cn) == pow(a, 2 ** k, Cn)


def test_no_cpu_fallback(qcs):
    if qcs.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(qcs.QcsError) as e:
        qcs.Register(3, 4)
    assert e.value.code == qcs.UNKNOWN_ERROR


def test_product_does_not_link_the_oracle():
    """The shipped libraries must not reference oracle code."""
    for name in ("libqcs.so", "libqcshost.so"):
        blob = open(os.path.join(ROOT, "quantumcomputer_b200", "lib", name), "rb").read()
        assert b"orc_" not in blob and b"qcref_" not in blob and b"libqcsoracle" not in blob
    for dirpath, _, files in os.walk(os.path.join(ROOT, "quantumcomputer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".c", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "qcs_oracle" not in src, f


def test_host_library_exports_its_headers(qcs):
    """libqcshost.so (the classical half + the debug helpers, host C) exports what its headers declare, and the
    product never mentions the test-only mock."""
    lib = C.CDLL(os.path.join(ROOT, "quantumcomputer_b200", "lib", "libqcshost.so"))
    host = os.path.join(ROOT, "quantumcomputer_b200", "host")
    names = set()
    for header in ("shor_classical.h", "state_debug.h", "mt19937.h"):
        text = re.sub(r"/\*.*?\*/", "", open(os.path.join(host, header)).read(), flags=re.S)
        names |= set(re.findall(r"\b(qcsh_[A-Za-z0-9_]+)\s*\(", text))
    assert {"qcsh_shors_algorithm", "qcsh_find_period", "qcsh_display_state", "qcsh_check_normalisation"} <= names
    for name in sorted(names):
        assert hasattr(lib, name), f"{name} declared under quantumcomputer_b200/host but not exported by libqcshost.so"
    for dirpath, _, files in os.walk(os.path.join(ROOT, "quantumcomputer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".c", ".h", ".cuh")):
                assert "mock_qcs" not in open(os.path.join(dirpath, f)).read(), f
    assert "mock" not in open(os.path.join(ROOT, "bench.py")).read()


def test_documents_name_only_entry_points_that_exist():
    """Every qcs_* / qcsh_* call INTEGRATION.md, README.md and DESIGN.md mention is declared in a header."""
    declared = set(declared_symbols())
    host = os.path.join(ROOT, "quantumcomputer_b200", "host")
    for header in os.listdir(host):
        if header.endswith(".h"):
            declared |= set(re.findall(r"\b(qcsh_[A-Za-z0-9_]+)\s*\(", open(os.path.join(host, header)).read()))
    internal = {"qcs_k_", "qcs_dist_", "qcs_peer_", "qcs_group_", "qcs_fused_", "qcs_pipeline_", "qcs_launch_", "qcs_map_",
                "qcs_materialise_", "qcs_sharded_", "qcs_profile_resolve", "qcs_fuse_flush", "qcs_internal", "qcs_oracle",
                "qcs_register_create_", "qcs_standin"}
    for doc in ("INTEGRATION.md", "README.md", "DESIGN.md"):
        text = open(os.path.join(ROOT, doc)).read()
        for name in set(re.findall(r"\b(qcsh?_[a-z][A-Za-z0-9_]*)\b", text)):
            if name in declared or any(name.startswith(p) for p in internal) or name in ("qcs_register", "qcs_group", "qcsh_rng", "qcsh_options"):
                continue
            raise AssertionError(f"{doc} mentions {name}, which no header declares")
