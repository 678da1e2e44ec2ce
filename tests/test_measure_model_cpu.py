"""CPU model of the exact parallel measurement scan (quantumcomputer_b200/csrc/measure.cu).

The reference's measure_state (qc_shor.c:283-292) adds |amp_i|^2 into one double in index order and stops
at the first i whose running sum reaches r; which i that is depends on every rounding on the way.  The
CUDA path reproduces the roundings in parallel: inside one binade the running sum is an integer number a
of ulps and one addend is the map a -> a + (a even ? de : do); such maps compose associatively, so chunks
collapse into one (de, do) pair.  This file restates that scheme step by step in plain Python (same
formulas as element_map / compose / classify / apply_map and the walk, including the DUAL class of chunks
astride one binade boundary and the refinement into sub-chunks) and checks it against the sequential loop
on the inputs that are hard for it: ties, plateaus on binade boundaries, addends below half an ulp, wide
dynamic range, variates on and next to prefix sums.  It pins the arithmetic argument; the kernels
themselves are compared with the oracle in tests/test_gates_gpu.py::test_parallel_measurement_*."""
import math
import struct

import numpy as np
import pytest

CHUNK, SUB = 64, 8                 # the kernels use 4096 and 32; the argument does not depend on the sizes
SEQ, ZERO = "seq", "zero"
SUB_DELTA = 2.0 ** -36


def sequential(p, s, r):
    """qc_shor.c:283-292 on probabilities p from running sum s: (found, index, sum)."""
    for i, x in enumerate(p):
        s = s + x
        if s >= r:
            return True, i, s
    return False, 0, s


def binade_of(x):
    return ((struct.unpack("<Q", struct.pack("<d", x))[0] >> 52) & 0x7FF) - 1023


def element_map(p, e):
    """(de, do) of one addend in binade e: x = p / ulp exactly; the tie goes by the parity of a."""
    x = p * 2.0 ** (52 - e)
    y = x + 2.0 ** 52
    t = x - (y - 2.0 ** 52)
    de = int(y) - (1 << 52)
    return de, de + (1 if t == 0.5 else 0) - (1 if t == -0.5 else 0)


def compose(f1, f2):
    de = f1[0] + (f2[1] if f1[0] & 1 else f2[0])
    od = f1[1] + (f2[1] if (1 + f1[1]) & 1 else f2[0])
    return de, od


def chunk_map(p, e):
    f = (0, 0)
    for x in p:
        f = compose(f, element_map(x, e))
    return f


def apply_map(s, f, e):
    """None when an invariant fails (the CUDA path then falls back to the sequential scan)."""
    if binade_of(s) != e:
        return None
    a = int(s / 2.0 ** (e - 52))                                  # exact: s is a multiple of its ulp
    a2 = a + (f[1] if a & 1 else f[0])
    if not (1 << 52) <= a2 < (1 << 53):
        return None
    return a2 * 2.0 ** (e - 52)


def classify(lo, hi, biggest, r, dual_ok):
    e = binade_of(lo)
    if biggest < 2.0 ** (e - 53):
        return ZERO
    if hi < r and e > -960:
        if hi < 2.0 ** (e + 1):
            return ("clean", e)
        if dual_ok and hi < 2.0 ** (e + 2):
            return ("dual", e)
    return SEQ


def model_scan(p, s0, r, n_qubits, stats, s0_approx=None):
    """The four passes of measure.cu on probabilities p: (found, index, sum) or None (invariant failed).
    s0_approx: the running sum before p as passes 1-3 know it on a shard of a sharded register (the shards'
    approximate totals, all-gathered) -- the walk alone carries the exact one (api.cu locate_state)."""
    if s0 >= r:                                                   # qcs_k_measure_scan: the sum only grows, so a running sum
        return sequential(p[:1], s0, r)                           # that already reaches r stops at the first index
    delta = 2.0 ** (n_qubits + 3 - 53)
    chunks = [p[c:c + CHUNK] for c in range(0, len(p), CHUNK)]
    csum = [float(np.sum(np.asarray(c))) for c in chunks]         # pass 1: any summation order
    cmax = [max(c) for c in chunks]
    codes, before = [], (s0 if s0_approx is None else s0_approx)
    for v, big in zip(csum, cmax):                                # pass 2: approximate prefix + classes
        lo, hi = before * (1.0 - delta), (before + v) * (1.0 + delta)
        codes.append(ZERO if v == 0.0 else classify(lo, hi, big, r, True) if lo > 0.0 else SEQ)
        before += v
    maps, upper = {}, {}
    for c, cd in enumerate(codes):                                # pass 3: maps (both binades of a DUAL chunk)
        if isinstance(cd, tuple):
            maps[c] = chunk_map(chunks[c], cd[1])
            if cd[0] == "dual":
                upper[c] = chunk_map(chunks[c], cd[1] + 1)
    s = s0
    for c, cd in enumerate(codes):                                # pass 4: the walk carries the exact sum
        if cd == ZERO:
            continue
        if isinstance(cd, tuple) and cd[0] == "clean":
            s = apply_map(s, maps[c], cd[1])
        elif isinstance(cd, tuple) and s >= 2.0 ** (cd[1] + 1):
            stats["dual_upper"] += 1
            s = apply_map(s, upper[c], cd[1] + 1)
        elif isinstance(cd, tuple) and (s + csum[c]) * (1.0 + SUB_DELTA) < 2.0 ** (cd[1] + 1):
            stats["dual_lower"] += 1
            s = apply_map(s, maps[c], cd[1])
        else:                                                     # refine: sub-chunks from the exact sum
            stats["refined"] += 1
            run, base = s, c * CHUNK
            for j in range(0, len(chunks[c]), SUB):
                sub = chunks[c][j:j + SUB]
                end = run + float(np.sum(np.asarray(sub)))
                lo, hi = run * (1.0 - SUB_DELTA), end * (1.0 + SUB_DELTA)
                sd = ZERO if max(sub) == 0.0 else classify(lo, hi, max(sub), r, False) if lo > 0.0 else SEQ
                run = end
                if sd == ZERO:
                    continue
                if isinstance(sd, tuple):
                    s = apply_map(s, chunk_map(sub, sd[1]), sd[1])
                    if s is None:
                        return None
                    continue
                found, i, s = sequential(sub, s, r)
                if found:
                    return True, base + j + i, s
        if s is None:
            return None
    return False, 0, s


def variates(p, rng):
    prefix = np.cumsum(p)
    rs = [1e-12, 0.1, 0.25, 0.5, 0.75, 0.9999, 1.0, 1.5] + [float(x) for x in rng.uniform(size=6)]
    for k in rng.integers(1, len(p) - 1, size=4):
        rs += [float(prefix[k]), float(np.nextafter(prefix[k], 0)), float(np.nextafter(prefix[k], 2))]
    return rs


def make_state(kind, n, rng):
    N = 1 << n
    if kind == "random":
        p = rng.normal(size=N) ** 2 + rng.normal(size=N) ** 2
    elif kind == "ties":                      # equal powers of two: every addition is exact or an exact tie
        p = np.zeros(N)
        p[::N // 256] = 2.0 ** -9
        p[1::N // 128] = 2.0 ** -10
        p[3::N // 512] = 2.0 ** -62
        return p
    elif kind == "plateau":                   # peaks of exactly 1/4 over addends around and below half an ulp
        p = rng.choice([1e-30, 9e-18, 2.7e-17, 5.5e-17, 1.1e-16], size=N, p=[0.9, 0.04, 0.03, 0.02, 0.01])
        p[N // 8::N // 4] = 0.25
        return p
    elif kind == "below":                     # the same, parked just BELOW 1/4: a DUAL chunk that runs in the lower binade
        p = rng.choice([1e-30, 9e-18, 2.7e-17, 5.5e-17, 1.1e-16], size=N, p=[0.9, 0.04, 0.03, 0.02, 0.01])
        p[N // 8::N // 4] = 0.25
        p[N // 8] = 0.25 * (1.0 - 2.0 ** -30)
        return p
    elif kind == "shor":                      # narrow peaks of very different height over exact zeros
        p = np.zeros(N)
        idx = rng.choice(N, size=48, replace=False)
        p[idx] = rng.uniform(size=48) ** 8
    else:                                     # wide dynamic range
        p = (rng.normal(size=N) ** 2) * np.exp(rng.normal(size=N) * 12)
    return p / p.sum()


def test_element_map_is_round_to_nearest_even():
    rng = np.random.default_rng(0)
    for e in (-3, 0, 5):
        ulp = 2.0 ** (e - 52)
        addends = [0.0, ulp / 2, ulp / 4, 3 * ulp / 2, 5 * ulp / 2, ulp, 2.0 ** (e - 1)] + \
                  [float(x) * 2.0 ** e for x in rng.uniform(size=40) ** 6]
        for a in [1 << 52, (1 << 52) + 1, (1 << 52) + 12345, (1 << 52) + 12346, (1 << 53) - (1 << 51)]:
            s = a * ulp
            for p in addends:
                want = s + p
                if binade_of(want) != e:
                    continue
                de, od = element_map(p, e)
                assert (a + (od if a & 1 else de)) * ulp == want, (e, a, p)


@pytest.mark.parametrize("kind", ["random", "ties", "plateau", "below", "shor", "wide"])
def test_model_scan_returns_the_sequential_index(kind):
    n = 13
    rng = np.random.default_rng(11)
    p = [float(x) for x in make_state(kind, n, rng)]
    stats = {"dual_upper": 0, "dual_lower": 0, "refined": 0}
    # the coarse margin of an n = 30 register (2^-20; any margin >= the rigorous one is valid): the band in which
    # a chunk is DUAL is then wider than the walk's own 2^-36 margin, as it is at the sizes the kernels run at
    margin_qubits = 30 if kind in ("plateau", "below") else n
    for r in variates(np.asarray(p), rng):
        got = model_scan(p[:-1], 0.0, r, margin_qubits, stats)   # index N-1 is the fall-through (qc_shor.c:283)
        assert got is not None, (kind, r)
        assert got == sequential(p[:-1], 0.0, r), (kind, r)
    assert stats["refined"] > 0
    if kind == "plateau":                                         # the cases the DUAL class exists for
        assert stats["dual_upper"] > 0
    if kind == "below":
        assert stats["dual_lower"] > 0


def test_compose_is_associative_and_matches_the_loop():
    rng = np.random.default_rng(3)
    e = -2
    p = [float(x) * 2.0 ** (e - 40) for x in rng.integers(0, 1 << 30, size=300)] + [2.0 ** (e - 53)] * 20 + [2.0 ** (e - 52) * 1.5] * 20
    rng.shuffle(p)
    for a in ((1 << 52) + 7, (1 << 52) + 8):
        s = a * 2.0 ** (e - 52)
        want = sequential(p, s, math.inf)[2]
        assert binade_of(want) == e
        whole = chunk_map(p, e)
        for cut in (1, 17, 150, 299):
            assert compose(chunk_map(p[:cut], e), chunk_map(p[cut:], e)) == whole
        assert apply_map(s, whole, e) == want


@pytest.mark.parametrize("kind", ["random", "plateau", "below", "shor", "wide"])
@pytest.mark.parametrize("shards", [2, 8])
def test_model_scan_sharded_hand_off(kind, shards):
    """Sharded registers (api.cu locate_state): every shard classifies its chunks from the all-gathered APPROXIMATE
    totals of the shards before it, the exact running sum is handed from shard to shard by the walk.  Same index
    as one sequential loop over the whole register."""
    n = 13
    rng = np.random.default_rng(23)
    p = [float(x) for x in make_state(kind, n, rng)]
    scanned = p[:-1]
    per = len(p) // shards
    stats = {"dual_upper": 0, "dual_lower": 0, "refined": 0}
    margin_qubits = 30 if kind in ("plateau", "below") else n
    for r in variates(np.asarray(p), rng):
        want = sequential(scanned, 0.0, r)
        s, got = 0.0, None
        totals = [float(np.sum(np.asarray(scanned[k * per:(k + 1) * per])[::-1])) for k in range(shards)]   # another order
        for k in range(shards):
            part = scanned[k * per:(k + 1) * per]
            res = model_scan(part, s, r, margin_qubits, stats, s0_approx=float(sum(totals[:k])))
            assert res is not None, (kind, r, k)
            found, i, s = res
            if found:
                got = (True, k * per + i, s)
                break
        if got is None:
            got = (False, 0, s)
        assert got == want, (kind, r)
