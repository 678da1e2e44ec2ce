#!/usr/bin/env python
"""make_dropin.py -- TEST INFRASTRUCTURE ONLY: the drop-in proof of INTEGRATION.md, executed.

Takes the reference program where it lies (/root/reference/qc_shor.c; nothing is copied into this
repository: the edited translation unit lives in a temporary directory for the duration of the compile),
applies the edits INTEGRATION.md section 1 lists -- and nothing else -- and links the result against
libqcs.so.  Outputs, next to the other binaries built from the reference (oracle/_ref/, git-ignored):

    oracle/_ref/qc_shor_ref            the UNMODIFIED reference program (GSL stand-in), for side-by-side runs
    oracle/_ref/qc_shor_dropin         the reference with its three primitive gates, reset, measurement and the
                                       allocation block forwarded to the C ABI; quantum_computation, inverse_QFT,
                                       find_period, shors_algorithm, the CLI: the reference's own code
    oracle/_ref/qc_shor_dropin_fused   the same, with quantum_computation forwarded as one call (the fast path)

tests/test_dropin.py runs them side by side (CPU: libqcs.so's entry points answered by tests/mock over the
oracle; GPU: the real library).  The RNG stays the host's (gsl_rng, qc_shor.c:281,1296-1299)."""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("QC_REF_SOURCE", "/root/reference/qc_shor.c")
OUT = os.path.join(HERE, "_ref")
CC = os.environ.get("CC", "gcc")

REGISTER = """typedef struct {
    int L_size;
    int M_size;
    unsigned int num_qubits;
    unsigned long int num_states;
    qcs_register *dev;      /* the state lives on the device(s) behind this handle */
} Register;"""

# function name -> new body (signatures, and with them every call site, stay as they are)
PRIMITIVES = {
    "swap_states": "    (void) reg;",
    "operate_matrix": "    (void) matrix; (void) reg;",
    "reset_register": "    qcs_reset_register(reg.dev);",
    "hadamard_gate": "    (void) matrix;\n    qcs_hadamard_gate(reg->dev, qubit_num);",
    "c_phase_shift_gate": "    (void) matrix;\n    qcs_c_phase_shift_gate(reg->dev, c_qubit_num, qubit_num, theta);",
    "c_amodc_gate": "    (void) matrix;\n    qcs_c_amodc_gate(reg->dev, C, atox, c_qubit_num);",
    "measure_state": ("    unsigned long long state_num = 0;\n"
                      "    qcs_measure_state(reg.dev, gsl_rng_uniform(rng), &state_num);   /* the draw of qc_shor.c:281 */\n"
                      "    return (unsigned long int) state_num;"),
}
COMPOSITES = {
    "inverse_QFT": "    (void) matrix;\n    qcs_inverse_QFT(reg->dev);",
    "quantum_computation": "    (void) matrix;\n    qcs_quantum_computation(reg->dev, C, a, QCS_POW_VERBATIM);",
}


def replace_body(src, name, body):
    m = re.search(r"^static [^\n;{]*\b%s\([^)]*\)\n\{\n" % re.escape(name), src, flags=re.M)
    assert m, name
    end = src.index("\n}\n", m.end() - 1)
    return src[:m.end()] + body + src[end:]


def replace_once(src, pattern, new, flags=re.S):
    out, n = re.subn(pattern, lambda _m: new, src, count=1, flags=flags)
    assert n == 1, pattern
    return out


def patched(src, fused):
    src = replace_once(src, r"#include <gsl/gsl_math\.h>\n", '#include <gsl/gsl_math.h>\n#include "qcs.h"\n')
    src = replace_once(src, r"typedef struct \{\n    int L_size;.*?\} Register;", REGISTER)
    for name, body in PRIMITIVES.items():
        src = replace_body(src, name, body)
    if fused:
        for name, body in COMPOSITES.items():
            src = replace_body(src, name, body)
    # main, qc_shor.c:1316-1324: one create instead of two vectors and the COO matrix
    src = replace_once(src, r"    reg\.state_a = gsl_vector_complex_alloc.*?reg\.new_state = &reg\.state_b;\n",
                       "    matrix = NULL;\n"
                       "    error = qcs_register_create(&reg.dev, reg.L_size, reg.M_size, -1);\n"
                       "    ERROR_CHECK(error);      /* the same ErrorCode values, qc_shor.c:164-170 */\n")
    # main, qc_shor.c:1330-1332: one destroy
    src = replace_once(src, r"    gsl_vector_complex_free\(reg\.state_a\);\n    gsl_vector_complex_free\(reg\.state_b\);\n"
                            r"    gsl_spmatrix_complex_free\(matrix\);\n",
                       "    qcs_register_destroy(reg.dev);\n")
    return src


def cc(args):
    subprocess.run([CC] + args, check=True)


def main():
    if not os.path.exists(REF):
        print("make_dropin: %s not present, keeping prebuilt oracle/_ref/ binaries (if any)" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    with open(REF) as f:
        src = f.read()
    shim = os.path.join(HERE, "gsl_shim")
    common = ["-O2", "-w", "-ffp-contract=off", "-I" + shim]
    cc(common + [REF, "-lm", "-o", os.path.join(OUT, "qc_shor_ref")])
    lib_dir = os.path.join(ROOT, "quantumcomputer_b200", "lib")
    with tempfile.TemporaryDirectory() as tmp:
        for fused, name in ((False, "qc_shor_dropin"), (True, "qc_shor_dropin_fused")):
            path = os.path.join(tmp, name + ".c")
            with open(path, "w") as f:
                f.write(patched(src, fused))
            cc(common + ["-I" + os.path.join(ROOT, "include"), path, "-L" + lib_dir, "-lqcs",
                         "-Wl,-rpath,$ORIGIN/../../quantumcomputer_b200/lib", "-lm", "-o", os.path.join(OUT, name)])
    print("make_dropin: built qc_shor_ref, qc_shor_dropin, qc_shor_dropin_fused in", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
