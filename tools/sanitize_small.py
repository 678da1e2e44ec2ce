"""Small end-to-end pass over every kernel family, meant to run under compute-sanitizer
(one tool per gpurun call): gates, fused QFT on every pipeline shape, the gate stream,
fused modular exponentiation, measurement, sampling, general gates, dense block."""
import math
import sys

import numpy as np

sys.path.insert(0, ".")
import quantumcomputer_b200 as q
from quantumcomputer_b200.workloads import apply_gates, layered_circuit

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
with q.Register(n, 0) as reg:
    reg.fill_synthetic(1)
    reg.scale(1.0 / math.sqrt(reg.norm2()))
    reg.hadamard_gate(3)
    reg.hadamard_gate(n - 1)
    reg.c_phase_shift_gate(2, n - 2, 0.3)
    for shape in range(6):
        reg.set_option(q.OPT_PIPE_SHAPE, shape)
        reg.inverse_QFT()
        reg.QFT()
    reg.set_option(q.OPT_PIPE_SHAPE, -1)
    reg.set_option(q.OPT_PIPELINE, 0)
    reg.inverse_QFT()
    reg.set_option(q.OPT_PIPELINE, 1)
    with reg.fused():
        apply_gates(reg, layered_circuit(n, 3))
    reg.apply_gate(5, np.array([[0.6, 0.8], [-0.8, 0.6]]))
    reg.apply_controlled_gate(1, n - 1, np.array([[0, 1], [1, 0]]))
    reg.apply_dense_block(3, np.eye(8))
    print("norm", reg.norm2(), "samples", reg.sample_states([0.1, 0.9]), "measured", reg.measure_state(0.5))
with q.Register(n - 5, 5) as reg:
    reg.reset_register()
    reg.quantum_computation(21, 2, q.POW_MODULAR)
    print("shor state norm", reg.norm2(), "measured", reg.measure_state(0.4))
    reg.set_option(q.OPT_FUSION, 0)
    reg.reset_register()
    reg.quantum_computation(21, 2, q.POW_MODULAR)
    print("gate by gate norm", reg.norm2())
print("SANITIZE_RUN_OK")
