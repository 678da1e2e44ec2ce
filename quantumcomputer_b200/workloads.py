"""Gate lists of the benchmark workloads, host side, no device work.

``layered_circuit`` is BASELINE configs[3]: D layers, layer d = hadamard_gate on every qubit in
ascending order, then c_phase_shift_gate on the pairs (q, (q+1+d) mod n) with
theta = 2 pi * gsl_rng_uniform drawn from gsl_rng_mt19937 seeded with n + d -- the generator the
reference draws its measurement variate from (qc_shor.c:281, 1296-1299)."""
import math


def mt19937_uniforms(seed, count):
    """gsl_rng_mt19937 + gsl_rng_uniform: seed 0 is replaced by 4357, output = u32 / 2^32."""
    mt = [0] * 624
    mt[0] = (seed or 4357) & 0xFFFFFFFF
    for i in range(1, 624):
        mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
    out, idx = [], 624
    for _ in range(count):
        if idx >= 624:
            for k in range(624):
                y = (mt[k] & 0x80000000) | (mt[(k + 1) % 624] & 0x7FFFFFFF)
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            idx = 0
        y = mt[idx]
        idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        out.append(y / 4294967296.0)
    return out


def layered_circuit(n, layers):
    """[("h", q) | ("cp", c, q, theta)] in program order; n + n gates per layer."""
    gates = []
    for d in range(layers):
        u = mt19937_uniforms(n + d, n)
        for q in range(n):
            gates.append(("h", q))
        for q in range(n):
            gates.append(("cp", q, (q + 1 + d) % n, 2.0 * math.pi * u[q]))
    return gates


def apply_gates(target, gates):
    """Issue the gates through the reference's operator names (works on a Register and on
    anything else that mirrors qc_shor.c's hadamard_gate / c_phase_shift_gate)."""
    for g in gates:
        if g[0] == "h":
            target.hadamard_gate(g[1])
        else:
            target.c_phase_shift_gate(g[1], g[2], g[3])
