"""Numpy prototype of the fused inverse-QFT sweep schedule (same pass/step/twiddle
decomposition as csrc/qft_fused.cu), used to validate the index math on the CPU.
Not part of the product."""
import math
import sys

import numpy as np


def plan_passes(n_local, lo, hi, T, a):
    """Returns a list of passes; each pass = dict(a, g_lo, g_hi, steps=[(s_tile, r)])
    tile = physical bits [0,a) U [g_lo,g_hi).  Stage bits processed descending."""
    t = min(T, n_local)
    passes = []
    top = hi          # exclusive
    if n_local <= t:
        a_eff = n_local
    else:
        a_eff = a
    # strided passes for stage bits >= t
    first_hi = max(lo, t)
    c_hi = max(0, top - first_hi)
    if c_hi > 0:
        cap = t - a_eff
        m = -(-c_hi // cap)
        sizes = [c_hi // m + (1 if i < c_hi % m else 0) for i in range(m)]
        for g in sizes:
            g_hi, g_lo = top, top - g
            passes.append(dict(a=a_eff, g_lo=g_lo, g_hi=g_hi, stage_lo=g_lo, stage_hi=g_hi))
            top = g_lo
    if top > lo:
        passes.append(dict(a=t, g_lo=t, g_hi=t, stage_lo=lo, stage_hi=top))   # contiguous [0,t)
    for p in passes:
        g = p["stage_hi"] - p["stage_lo"]
        k = -(-g // 4)
        sizes = [g // k + (1 if i < g % k else 0) for i in range(k)]
        steps = []
        l = p["stage_hi"]
        for r in sizes:
            phys_low = l - r
            # tile-local position of physical bit phys_low
            s_tile = phys_low if phys_low < p["a"] else p["a"] + (phys_low - p["g_lo"])
            steps.append((s_tile, r, l - 1))      # (tile-local low bit, radix bits, physical top bit)
            l -= r
        p["steps"] = steps
    return passes


def dft_dif(x, r):
    """in-register radix-2^r DIF with + sign, output in place (bit-reversed frequency), unscaled."""
    R = 1 << r
    x = list(x)
    span = R >> 1
    while span >= 1:
        for start in range(0, R, 2 * span):
            for m in range(span):
                u, v = x[start + m], x[start + m + span]
                x[start + m] = u + v
                x[start + m + span] = (u - v) * np.exp(1j * math.pi * m / span)
        span >>= 1
    return x


def bitrev(k, r):
    return int(format(k, f"0{r}b")[::-1], 2) if r else 0


def fused_iqft(state, n_local, lo, hi, T=6, a=2):
    st = state.copy()
    for p in plan_passes(n_local, lo, hi, T, a):
        a_, g_lo, g_hi = p["a"], p["g_lo"], p["g_hi"]
        t = a_ + (g_hi - g_lo)

        def spread(e):
            return (e & ((1 << a_) - 1)) | ((e >> a_) << g_lo)

        n_outer = 1 << (n_local - t)
        gbits = p["stage_hi"] - p["stage_lo"]
        scale = (1 / math.sqrt(2)) ** gbits
        for o in range(n_outer):
            base = ((o >> (g_lo - a_)) << g_hi) | ((o & ((1 << (g_lo - a_)) - 1)) << a_) if g_lo > a_ else (o << t)
            tile = np.array([st[base | spread(e)] for e in range(1 << t)])
            for si, (s, r, l) in enumerate(p["steps"]):
                j = l - lo
                R = 1 << r
                low_phys = l - r + 1
                y_base = (base & ((1 << low_phys) - 1)) >> lo
                wb = np.exp(1j * math.pi * y_base / (1 << j)) if j > 0 else 1.0
                last = si == len(p["steps"]) - 1
                for c in range(1 << (t - r)):
                    e_base = ((c >> s) << (s + r)) | (c & ((1 << s) - 1))
                    y_col = (spread(e_base) & ((1 << low_phys) - 1)) >> lo
                    wc = np.exp(1j * math.pi * y_col / (1 << j)) if j > 0 else 1.0
                    w = wb * wc
                    x = [tile[e_base + (d << s)] for d in range(R)]
                    x = dft_dif(x, r)
                    for d in range(R):
                        k = bitrev(d, r)
                        v = x[d] * (w ** k)
                        if last:
                            v *= scale
                        tile[e_base + (d << s)] = v
            for e in range(1 << t):
                st[base | spread(e)] = tile[e]
    return st


if __name__ == "__main__":
    sys.path.insert(0, ".")
    from oracle import Restatement
    rng = np.random.default_rng(1)
    for (L, M, T, a) in [(3, 4, 12, 4), (5, 5, 6, 2), (10, 0, 6, 2), (9, 2, 5, 2), (11, 0, 7, 3), (8, 3, 4, 1), (12, 0, 6, 3)]:
        n = L + M
        v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        v /= np.linalg.norm(v)
        o = Restatement(L, M)
        o.set_state(v)
        o.inverse_QFT()
        want = o.get_state()
        got = fused_iqft(v, n, M, n, T, a)
        print(L, M, T, a, [(p["a"], p["g_lo"], p["g_hi"], p["steps"]) for p in plan_passes(n, M, n, T, a)],
              "err", np.linalg.norm(got - want) / np.linalg.norm(want))
