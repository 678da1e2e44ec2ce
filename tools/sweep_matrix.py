"""Tuning matrix for the pipelined sweep kernel: inverse QFT over all n qubits for every
(pipeline shape, direct store, shortest run) combination, with a closed-form correctness probe
per configuration.  Run on the GPU box; prints one JSON line per configuration."""
import json
import math
import subprocess
import sys

import numpy as np

sys.path.insert(0, ".")
import quantumcomputer_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
shapes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3, 4, 5, 6]
directs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
runs = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [3]
N = 1 << n
if len(shapes) > 1:
    # one process per shape under a timeout: an experimental shape that hangs must not take
    # the whole matrix (or the GPU box) with it
    for sh in shapes:
        try:
            subprocess.run([sys.executable, __file__, str(n), str(reps), str(sh), ",".join(map(str, directs)),
                            ",".join(map(str, runs))], timeout=90)
        except subprocess.TimeoutExpired:
            print(json.dumps({"n": n, "shape": sh, "error": "timeout"}), flush=True)
    sys.exit(0)


def bitrev(j):
    return int(format(j, f"0{n}b")[::-1], 2)


with q.Register(n, 0) as reg:
    for shape in shapes:
        for direct in directs:
            for run_bits in runs:
                reg.set_option(q.OPT_PIPE_SHAPE, shape)
                reg.set_option(q.OPT_MIN_RUN_BITS, run_bits)
                # correctness: inverse_QFT |k> = e^{2 pi i jk/N}/sqrt(N) at bit-reversed j
                k = 0x1C0FFEE1 % N
                reg.reset_register()
                reg.set_state(np.array([0j, 0j]), first=0)
                reg.set_state(np.array([1 + 0j]), first=k)
                reg.inverse_QFT()
                errs = []
                for j in (0, 1, 5, 123456789 % N, N - 1, N // 2 + 77, 0x2AAAAAAA % N, 4097):
                    got = reg.get_state(bitrev(j), 1)[0]
                    want = np.exp(2j * math.pi * ((j * k) % N) / N) / math.sqrt(N)
                    errs.append(abs(got - want) / abs(want))
                reg.fill_synthetic(1234)
                reg.scale(1.0 / math.sqrt(reg.norm2()))
                for _ in range(2):
                    reg.inverse_QFT()
                reg.synchronize()
                reg.set_option(q.OPT_PROFILE, 1)
                reg.profile_reset()
                reg.timer_start()
                for _ in range(reps):
                    reg.inverse_QFT()
                ms = reg.timer_stop() / reps
                prof = reg.profile()
                reg.set_option(q.OPT_PROFILE, 0)
                launches, kms, kbytes = prof["tile_sweep"]
                print(json.dumps({"n": n, "shape": shape, "direct_store": direct, "min_run_bits": run_bits,
                                  "ms_per_iqft": round(ms, 3), "sweeps_per_iqft": launches // reps,
                                  "sweep_GBps": round(kbytes / (kms * 1e-3) / 1e9, 1),
                                  "gates_per_s": round((n + n * (n - 1) // 2) / (ms * 1e-3), 1),
                                  "closed_form_max_rel_err": float(max(errs)),
                                  "norm": reg.norm2()}), flush=True)
