"""Time the inverse QFT at n qubits under a few option settings (one JSON line each).

    python tools/run_qft_variants.py [n] [steps] [lag,lag,...]      (QCS_NO_SPLIT3=1: plain layout of the contiguous sweep)
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import quantumcomputer_b200 as q  # noqa: E402


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    lags = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [444]
    variants = [("pair off", {q.OPT_L2_PAIR: 0, q.OPT_SPLIT3: 1})]
    variants += [(f"pair lag {lag}", {q.OPT_L2_PAIR: 1, q.OPT_L2_PAIR_LAG: lag}) for lag in lags]
    variants += [("pair off, plain swizzle", {q.OPT_L2_PAIR: 0, q.OPT_SPLIT3: 0})]
    N = 1 << n
    with q.Register(n, 0) as reg:
        for name, opts in variants:
            for k, v in opts.items():
                reg.set_option(k, v)
            # closed-form check
            kk = (N - 1) - 0x12345
            reg.reset_register()
            reg.set_state(np.array([0j]), first=1)
            reg.set_state(np.array([1 + 0j]), first=kk)
            reg.inverse_QFT()
            err = 0.0
            for i in (0, 1, 77, N // 2 + 3, N - 1, 123456789 % N, (N // 3) | 1):
                j = bitrev(i, n)
                want = np.exp(2j * math.pi * ((j * kk) % N) / N) / math.sqrt(N)
                err = max(err, abs(reg.get_state(i, 1)[0] - want) * math.sqrt(N))
            reg.QFT()
            back = abs(reg.get_state(kk, 1)[0] - 1.0)
            reg.fill_synthetic(1234)
            reg.scale(1.0 / math.sqrt(reg.norm2()))
            for _ in range(3):
                reg.inverse_QFT()
            reg.set_option(q.OPT_PROFILE, 1)
            reg.profile_reset()
            reg.timer_start()
            for _ in range(steps):
                reg.inverse_QFT()
            ms = reg.timer_stop()
            prof = reg.profile()
            reg.set_option(q.OPT_PROFILE, 0)
            reg.timer_start()
            for _ in range(steps):
                reg.QFT()
            ms_f = reg.timer_stop()
            print(json.dumps({"n": n, "variant": name, "ms_per_iqft": ms / steps, "ms_per_qft": ms_f / steps,
                              "closed_form_err": err, "forward_back_err": back, "norm": reg.norm2(),
                              "tile_sweep": prof["tile_sweep"]}), flush=True)


if __name__ == "__main__":
    main()
