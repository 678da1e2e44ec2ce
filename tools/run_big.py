"""n = 33 on one B200 (128 GiB in place): fused inverse QFT / QFT round trip, closed form, norm,
exact measurement.  Prints timings; run on the GPU box."""
import json
import math
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import quantumcomputer_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 33
out = {"n": n}
N = 1 << n
with q.Register(n, 0) as reg:
    # closed form: inverse_QFT |k> = e^{2 pi i jk/N}/sqrt(N) at bit-reversed j
    k = 0x1C0FFEE1 % N
    reg.reset_register()
    reg.set_state(np.array([0j, 0j]), first=0)
    reg.set_state(np.array([1 + 0j]), first=k)
    reg.timer_start(); reg.inverse_QFT(); out["iqft_basis_ms"] = reg.timer_stop()
    def bitrev(j):
        return int(format(j, f"0{n}b")[::-1], 2)
    errs = []
    for j in (0, 1, 5, 123456789 % N, N - 1, N // 2 + 77):
        got = reg.get_state(bitrev(j), 1)[0]
        want = np.exp(2j * math.pi * ((j * k) % N) / N) / math.sqrt(N)
        errs.append(abs(got - want) / abs(want))
    out["closed_form_max_rel_err"] = max(errs)
    out["norm_after_iqft"] = reg.norm2()
    reg.timer_start(); reg.QFT(); out["qft_ms"] = reg.timer_stop()
    out["round_trip_amp_k"] = abs(reg.get_state(k, 1)[0])
    # random state: timing + exact measurement
    reg.fill_synthetic(1234)
    s = reg.norm2()
    reg.scale(1.0 / math.sqrt(s))
    for _ in range(2):
        reg.inverse_QFT()
    reg.synchronize()
    reg.timer_start()
    for _ in range(3):
        reg.inverse_QFT()
    ms = reg.timer_stop() / 3
    gates = n + n * (n - 1) // 2
    out["iqft_random_ms"] = ms
    out["gates_per_s"] = gates / (ms * 1e-3)
    out["sweeps_GBps"] = None
    prof = reg.profile()
    out["tile_sweep_launches_total"] = prof["tile_sweep"][0]
    out["norm_random_after"] = reg.norm2()
    t0 = time.time(); idx = reg.measure_state(0.6180339887); out["measure_s"] = time.time() - t0
    out["measured_index"] = idx
print(json.dumps(out))
