"""Small driver used under ncu: the layered H / C-phase circuit (BASELINE configs[3]) at n qubits
through the deferred gate stream."""
import math
import sys

sys.path.insert(0, ".")
import quantumcomputer_b200 as q
from quantumcomputer_b200.workloads import apply_gates, layered_circuit

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
gates = layered_circuit(n, layers)
with q.Register(n, 0) as reg:
    reg.fill_synthetic(1234)
    reg.scale(1.0 / math.sqrt(reg.norm2()))
    for _ in range(2):
        with reg.fused():
            apply_gates(reg, gates)
    reg.synchronize()
    print("norm", reg.norm2(), "passes per layer", len(q.schedule_describe(n, gates)) / layers)
