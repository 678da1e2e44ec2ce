"""Small driver used under ncu: one fused inverse QFT at n qubits (default 28)."""
import math
import sys

sys.path.insert(0, ".")
import quantumcomputer_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
tb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
shape = int(sys.argv[4]) if len(sys.argv) > 4 else -1
prefetch = int(sys.argv[5]) if len(sys.argv) > 5 else 0
with q.Register(n, 0) as reg:
    reg.set_option(q.OPT_TILE_BITS, tb)
    reg.set_option(q.OPT_PIPE_SHAPE, shape)
    reg.set_option(q.OPT_PREFETCH_TILES, prefetch)
    reg.fill_synthetic(1234)
    reg.scale(1.0 / math.sqrt(reg.norm2()))
    reg.inverse_QFT()
    reg.synchronize()
    reg.timer_start()
    for _ in range(reps):
        reg.inverse_QFT()
    print("ms per inverse_QFT", reg.timer_stop() / reps)
    print("norm", reg.norm2())
