#!/usr/bin/env python
"""bench.py -- headline benchmark of the gate-application path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2]): the reference's inverse_QFT (qc_shor.c:678-690)
over all n qubits (L = n, M = 0) of a synthetic random state, n = 30 on one GPU:
30 Hadamards + 435 controlled phase rotations = 465 gates per step.  With N > 1
GPUs (one process per GPU, torchrun) the state is sharded by its top log2(N)
qubits and n = 30 + log2(N) (weak scaling: 2^30 amplitudes = 16 GiB per GPU).

One JSON line is printed by rank 0; see README/DESIGN.md for the fields.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 1234
CLASS_BYTES_NOTE = {
    "diag_multi": "32*2^n B per launch (read+write every amplitude, any number of diagonal gates)",
    "hadamard": "32*2^n B per launch (read+write every amplitude)",
    "cphase": "8*2^n B per launch (read+write the |11> quarter)",
    "tile_sweep": "32*2^n B per launch (read+write every amplitude once per sweep)",
}


def qft_gate_count(n):
    return n + n * (n - 1) // 2


def profiled_traffic(n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the tile-sweep kernel from the
    committed `ncu --set full` capture of the same workload (profiles/, summarised by
    tools/ncu_summary.py), or None.  Not measured in this run: the line says so."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_qft_sweeps_final_n30.csv")
    if n != 30 or not os.path.exists(path):
        return None, None
    try:
        import csv
        with open(path) as f:
            rows = list(csv.reader(f))
        hdr, units = rows[0], rows[1]
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        per = [float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]] for r in rows[2:] if len(r) > wr]
        return sum(per) / len(per), os.path.relpath(path, ROOT)
    except Exception:
        return None, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                p = json.load(f)
            return float(p["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------
# CPU arms (the only places that execute oracle/)
# --------------------------------------------------------------------------
REF_N = 12          # the reference arm and cpu_baseline time the same n (4^n index scan per gate: 3-7 s per IQFT)


def reference_iqft_seconds(n):
    """One inverse_QFT over all n qubits with the unmodified reference
    (oracle/_ref) if it was compiled, else the oracle restatement."""
    import oracle
    if oracle.have_reference():
        kind, obj = "reference", oracle.Reference(n, 0)
    else:
        if not oracle.have_restatement():
            oracle.build()
        kind, obj = "port", oracle.Restatement(n, 0)
    gen = oracle.Restatement(n, 0)
    gen.fill_synthetic(SEED)
    gen.scale(1.0 / math.sqrt(gen.norm2()))
    obj.set_state(gen.get_state())
    t0 = time.perf_counter()
    obj.inverse_QFT()
    dt = time.perf_counter() - t0
    obj.close()
    gen.close()
    return kind, dt


def reference_n_for(total_runs):
    """n of the reference runs: REF_N unless (steps + warmup) of them would take much more than a few
    minutes (cost grows ~4.5x per qubit, BASELINE.md section 2: 3-7 s at n = 12)."""
    n = REF_N
    while n > 8 and total_runs * 7.0 * 4.5 ** (n - REF_N) > 240.0:
        n -= 1
    return n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    n = args.qubits if args.qubits else reference_n_for(total)
    kind = None
    for _ in range(args.warmup):
        kind, _dt = reference_iqft_seconds(n)
    times = []
    for _ in range(args.steps):
        kind, dt = reference_iqft_seconds(n)
        times.append(dt)
    gates = qft_gate_count(n)
    total_s = sum(times)
    value = gates * len(times) / total_s
    sample = (f"inverse_QFT over all n={n} qubits ({gates} gates) of the synthetic state, seed {SEED}; "
              f"the reference builds each gate as a 4^n-scan COO matrix, so n=30 is out of reach; "
              f"serial program: 1 thread of {os.cpu_count()} host cores")
    line = {
        "impl": "reference", "metric": "qft_gates_per_sec", "value": value, "unit": "gates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_s / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"inverse_QFT n={n} (bounded sample of the n=30 workload)", "qubits": n,
                   "gates_per_step": gates},
        "cpu_baseline": {"value": value, "unit": "gates/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def port_gate_rate(n, all_cores, budget_s):
    """The matrix-free restatement (oracle/qcs_oracle.c, O(2^n) per gate, the reference's arithmetic
    order) on the inverse-QFT gate sequence of an n-qubit synthetic state, in program order
    (qc_shor.c:682-689), stopped after budget_s seconds: gates done / time."""
    import oracle
    cls = oracle.RestatementAllCores if all_cores else oracle.Restatement
    obj = cls(n, 0)
    threads = obj.threads()
    obj.fill_synthetic(SEED)
    obj.scale(1.0 / math.sqrt(obj.norm2()))
    done, t0 = 0, time.perf_counter()
    for l in range(n - 1, -1, -1):
        obj.hadamard_gate(l)
        done += 1
        for k in range(l - 1, -1, -1):
            obj.c_phase_shift_gate(l, k, math.pi / float(1 << (l - k)))
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    obj.close()
    return {"value": done / dt, "unit": "gates/s", "cores": threads, "kind": "port", "qubits": n,
            "sample": f"first {done} of the {qft_gate_count(n)} gates of inverse_QFT at n={n} in program order "
                      f"({dt:.1f} s), matrix-free restatement, {threads} thread(s) of {os.cpu_count()} host cores",
            "n30_equivalent": done / dt * 2.0 ** (n - 30)}


def cpu_baseline():
    """(1) the unmodified reference at the reference arm's n; (2) the fair O(2^n)-per-gate port on one
    thread (n = 24) and on all cores (n = 26), each a bounded sample."""
    import oracle
    n = REF_N
    kind, dt = reference_iqft_seconds(n)
    gates = qft_gate_count(n)
    out = {"value": gates / dt, "unit": "gates/s", "cores": 1, "kind": kind,
           "sample": f"one inverse_QFT over all n={n} qubits ({gates} gates, {dt:.2f} s) of the synthetic "
                     f"state, seed {SEED}; serial program, 1 thread of {os.cpu_count()} host cores"}
    try:
        out["port_1_thread"] = port_gate_rate(24, False, 8.0)
        if oracle.have_restatement_omp():
            out["port_all_cores"] = port_gate_rate(26, True, 8.0)
    except Exception as exc:                       # the baseline must never take the bench line down
        out["port_error"] = repr(exc)
    return out


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
NV_PEER_GBS = 770.0      # measured peer copy per direction per GPU (B200_PROFILING.md; nominal 900)


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


class Ranks:
    """torch.distributed plumbing of the bench (rendezvous, barrier, max / all-ok over ranks)."""

    def __init__(self, q):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.q = q
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
            self.torch = torch

    def comm_id(self):
        if self.dist is None:
            return None
        ids = [self.q.comm_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(ids, src=0)
        return ids[0]

    def barrier(self, reg=None):
        if reg is not None:
            reg.synchronize()
        if self.dist is not None:
            self.dist.barrier()

    def max(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def parity_block(q, reg, n, ranks):
    """In-run correctness evidence on the register the timed region uses (every rank takes part):
      * inverse_QFT of a basis state |k> against the closed form e^{2 pi i jk/N}/sqrt(N) at
        bit-reversed j (SURVEY KAT-4, qc_shor.c:678-690), 16 probes in every rank's shard;
      * QFT after inverse_QFT restores the synthetic state (64 probes per shard) and the norm;
      * sample_states (one scan for all variates) against one locate per variate (the code path
        measure_state runs), then measure_state itself: same index, collapsed state has norm 1.
    Errors are relative to the amplitude scale 1/sqrt(N); the bar is north_star's 1e-12."""
    import numpy as np
    rank, world = ranks.rank, ranks.world
    N, nl = 1 << n, reg.local_states
    scale = 1.0 / math.sqrt(N)
    rng = np.random.default_rng(97 + rank)
    probes = sorted(set([0, 1, nl - 1, nl // 2, nl // 2 + 77] + [int(x) for x in rng.integers(0, nl, size=11)]))

    # ---- closed form
    k = (N - 1) - 0x12345 if n > 20 else N - 3           # lives in the last rank's shard: all global bits set
    reg.reset_register()                                  # |0...01>
    if rank == 0:
        reg.set_state(np.array([0j]), first=1)
    if k // nl == rank:
        reg.set_state(np.array([1 + 0j]), first=k % nl)
    reg.inverse_QFT()
    err_closed = 0.0
    for li in probes:
        j = bitrev(rank * nl + li, n)
        want = np.exp(2j * math.pi * ((j * k) % N) / N) * scale
        err_closed = max(err_closed, abs(reg.get_state(li, 1)[0] - want) / scale)
    norm_closed = reg.norm2()
    reg.QFT()
    back = abs(reg.get_state(k % nl, 1)[0] - 1.0) if k // nl == rank else 0.0

    # ---- round trip on the synthetic state
    reg.fill_synthetic(SEED)
    reg.scale(1.0 / math.sqrt(reg.norm2()))
    probes2 = sorted(set(probes + [int(x) for x in rng.integers(0, nl, size=48)]))
    before = np.array([reg.get_state(li, 1)[0] for li in probes2])
    reg.inverse_QFT()
    mid = np.array([reg.get_state(li, 1)[0] for li in probes2])
    reg.QFT()
    after = np.array([reg.get_state(li, 1)[0] for li in probes2])
    err_round = float(np.max(np.abs(after - before))) / scale
    moved = float(np.max(np.abs(mid - before))) / scale    # the transform did something
    norm_round = reg.norm2()

    # ---- measurement (state: the synthetic one again, to rounding)
    rs = [0.123456789, 0.5, 0.987654321]
    many = [int(x) for x in reg.sample_states(rs)]
    single = [int(reg.sample_states([r])[0]) for r in rs]
    # i.i.d. random amplitudes: the cumulative probability is linear in the index to ~1/sqrt(N)
    linear = max(abs(idx / N - r) for idx, r in zip(many, rs))
    measured = int(reg.measure_state(rs[0]))
    norm_collapsed = reg.norm2()
    # (1e-3 at the sizes the bench runs at, n >= 24; wider for the small registers of --qubits runs)
    measure_ok = many == single and measured == many[0] and norm_collapsed == 1.0 and linear < max(1e-3, 4.0 / math.sqrt(N))

    err_closed = ranks.max(max(err_closed, back))       # `back`: |amp[k] - 1| after the forward transform
    err_round = ranks.max(err_round)
    moved = ranks.max(moved)
    bad = ranks.max(0.0 if measure_ok else 1.0)
    ok = (err_closed <= 1e-12 and err_round <= 1e-12 and moved > 1e-3 and bad == 0.0 and
          abs(norm_closed - 1.0) < 1e-12 and abs(norm_round - 1.0) < 1e-12)
    return {"ok": bool(ok), "max_rel_err": max(err_closed, err_round),
            "closed_form_max_rel_err": err_closed, "closed_form_probes": len(probes) * world,
            "round_trip_max_rel_err": err_round, "round_trip_probes": len(probes2) * world,
            "norm_after_closed_form": norm_closed, "norm_after_round_trip": norm_round,
            "measure": {"ok": bool(bad == 0.0), "indices": many, "variates": rs,
                        "sample_states_equals_locate": many == single, "measure_state_index": measured,
                        "norm_after_collapse": norm_collapsed},
            "tolerance": 1e-12,
            "what": "inverse_QFT|k> vs closed form at bit-reversed probes in every shard; QFT(inverse_QFT(state)) "
                    "vs state; sample_states vs per-variate locate vs measure_state"}


def roofline_from_profile(prof, n, world, circuit_is_qft, peak, peak_src):
    dom = max((k for k in prof if prof[k][0] > 0), key=lambda k: prof[k][1])
    d_launches, d_ms, d_bytes = prof[dom]
    achieved = d_bytes / (d_ms * 1e-3) / 1e9 if d_ms > 0 else 0.0
    kernel_ms = sum(v[1] for v in prof.values())
    if dom == "global_sweep":
        # multi-GPU: the sweep over the global qubits is bound by NVLink; the denominator is the
        # measured peer copy per direction per GPU stated in B200_PROFILING.md (nominal 900)
        return {"bound": "nvlink", "kernel": dom, "achieved": achieved, "peak": NV_PEER_GBS, "unit": "GB/s",
                "frac": achieved / NV_PEER_GBS,
                "peak_source": "B200_PROFILING.md measured peer copy per direction per GPU (of measured)",
                "traffic": None, "launches": d_launches, "avg_launch_ms": d_ms / d_launches,
                "algorithmic_bytes_per_launch": d_bytes / d_launches,
                "share_of_kernel_time": d_ms / kernel_ms if kernel_ms else None,
                "bytes_model": "NVLink bytes per direction per GPU and launch (remote reads, or remote writes, of the "
                               "rank's share of the tiles of the global sweep)",
                "local_sweeps_GBps": (prof["tile_sweep"][2] / (prof["tile_sweep"][1] * 1e-3) / 1e9
                                      if prof["tile_sweep"][1] > 0 else None)}
    traffic, traffic_src = (profiled_traffic(n) if (dom == "tile_sweep" and circuit_is_qft and world == 1)
                            else (None, None))
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src + " (of measured)",
            "traffic": traffic,
            "traffic_source": (traffic_src + " (committed ncu --set full capture of this workload, not this run)"
                               if traffic_src else None),
            "launches": d_launches, "avg_launch_ms": d_ms / d_launches,
            "algorithmic_bytes_per_launch": d_bytes / d_launches,
            "share_of_kernel_time": d_ms / kernel_ms if kernel_ms else None,
            "bytes_model": CLASS_BYTES_NOTE.get(dom, "")}


def timed_steps(q, reg, ranks, one_step, steps, warmup, sample_clocks):
    """W warm-up steps, then K steps between barriers, device-timed, max over ranks."""
    for _ in range(warmup):
        one_step()
    ranks.barrier(reg)
    reg.set_option(q.OPT_PROFILE, 1)
    reg.profile_reset()
    sampler = ClockSampler(ranks.local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
    ranks.barrier(reg)
    reg.timer_start()
    for _ in range(steps):
        one_step()
    ms = reg.timer_stop()
    ranks.barrier(reg)
    clocks = sampler.stop() if sampler else None
    ms = ranks.max(ms)
    prof = reg.profile()
    launches = reg.launch_count
    reg.set_option(q.OPT_PROFILE, 0)
    return ms, prof, launches, clocks


def apply_tuning(q, reg, args):
    reg.set_option(q.OPT_FUSION, 0 if args.no_fusion else 1)
    if args.tile_bits:
        reg.set_option(q.OPT_TILE_BITS, args.tile_bits)
    if args.prefetch >= 0:
        reg.set_option(q.OPT_PREFETCH_TILES, args.prefetch)
    if args.pipe_shape >= 0:
        reg.set_option(q.OPT_PIPE_SHAPE, args.pipe_shape)
    if args.min_run_bits > 0:
        reg.set_option(q.OPT_MIN_RUN_BITS, args.min_run_bits)
    if args.global_run_bits > 0:
        reg.set_option(q.OPT_GLOBAL_RUN_BITS, args.global_run_bits)
    if args.overlap_slices >= 0:
        reg.set_option(q.OPT_OVERLAP_SLICES, args.overlap_slices)
    if args.global_sms > 0:
        reg.set_option(q.OPT_GLOBAL_SMS, args.global_sms)
    if args.l2_pair >= 0:
        reg.set_option(q.OPT_L2_PAIR, args.l2_pair)


def north_star_block(q, ranks, args, peak, peak_src, n_override=None):
    """BASELINE `metric`: QFT time at n = 33 (1 GPU), 34 (2), 35 (4 and 8), timed in this run.
    (n_override: the CPU dry run of tests/test_bench_flow_cpu.py.)"""
    world = ranks.world
    n = n_override or {1: 33, 2: 34, 4: 35, 8: 35}.get(world)
    if n is None:
        return None
    p = int(math.log2(world))
    reg = q.Register(n, 0, device=ranks.local_rank, rank=ranks.rank, world_size=world, comm_id=ranks.comm_id())
    apply_tuning(q, reg, args)
    try:
        import numpy as np
        # parity at this size: closed form of a basis state at 8 probes per shard
        N, nl = 1 << n, reg.local_states
        k = (N - 1) - 0x54321 if n > 20 else N - 3
        reg.reset_register()
        if ranks.rank == 0:
            reg.set_state(np.array([0j]), first=1)
        if k // nl == ranks.rank:
            reg.set_state(np.array([1 + 0j]), first=k % nl)
        reg.inverse_QFT()
        scale = 1.0 / math.sqrt(N)
        err = 0.0
        for li in (0, 1, nl - 1, nl // 2 + 5, nl // 3, nl // 5, nl // 7, (nl // 11) * 3):
            j = bitrev(ranks.rank * nl + li, n)
            want = np.exp(2j * math.pi * ((j * k) % N) / N) * scale
            err = max(err, abs(reg.get_state(li, 1)[0] - want) / scale)
        err = ranks.max(err)
        norm_closed = reg.norm2()
        reg.fill_synthetic(SEED)
        reg.scale(1.0 / math.sqrt(reg.norm2()))
        steps = max(1, min(args.steps, args.north_star_steps))
        ms, prof, launches, _ = timed_steps(q, reg, ranks, reg.inverse_QFT, steps, 3, False)
        norm_out = reg.norm2()
        gates = qft_gate_count(n)
        out = {"qubits": n, "n_gpus": world, "state_bytes_per_gpu": int(16 * nl), "steps": steps, "warmup": 3,
               "ms_per_qft": ms / steps, "gates_per_step": gates, "gates_per_sec": gates * steps / (ms * 1e-3),
               "n30_equivalent_gates_per_sec": gates * 2.0 ** (n - 30) * steps / (ms * 1e-3),
               "parallelism": f"top {p} qubits global", "gpu_launches": int(launches),
               "parity": {"ok": bool(err <= 1e-12 and abs(norm_closed - 1.0) < 1e-12 and abs(norm_out - 1.0) < 1e-10),
                          "closed_form_max_rel_err": err, "closed_form_probes": 8 * world,
                          "norm_after_closed_form": norm_closed, "norm_after_timed_steps": norm_out},
               "roofline": roofline_from_profile(prof, n, world, False, peak, peak_src),
               "kernels": {kk: {"launches": v[0], "ms": round(v[1], 4),
                                "GBps": round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None}
                           for kk, v in prof.items() if v[0]}}
    finally:
        reg.close()
    return out


def shor_block(q, ranks, args, with_n30=True):
    """BASELINE configs[0] / [1]: wall time of the quantum half of find_period (qc_shor.c:922-928:
    reset_register, quantum_computation, measure_state) through the C ABI, beside the reference's
    own time for the same calls on the host, the measured index compared run by run; plus one
    n = 30 quantum_computation (L = 18, M = 12) for the throughput of the modular-exponentiation sweep."""
    from quantumcomputer_b200.workloads import mt19937_uniforms
    import numpy as np
    import oracle
    if not oracle.have_restatement():
        oracle.build()
    peak, peak_src = measured_peaks()
    runs = max(5, min(args.steps, 40))
    cases = []
    specs = [("cfg1 (BASELINE configs[0])", 15, 7, 3, 4, q.POW_VERBATIM, 12345),
             ("cfg2 reference-safe (BASELINE configs[1])", 21, 2, 5, 5, q.POW_VERBATIM, 2021),
             ("cfg2 full size, intended a^x mod C", 21, 2, 10, 5, q.POW_MODULAR, 2021)]
    for name, Cn, a, L, M, mode, seed in specs:
        n = L + M
        rs = mt19937_uniforms(seed, runs)
        with q.Register(L, M, device=ranks.local_rank) as reg:
            for _ in range(3):                                   # warm-up (module load, allocations)
                reg.reset_register(); reg.quantum_computation(Cn, a, mode); reg.measure_state(0.5)
            before = reg.launch_count
            got, t_gpu = [], []
            for r in rs:
                t0 = time.perf_counter()
                reg.reset_register()
                reg.quantum_computation(Cn, a, mode)
                got.append(int(reg.measure_state(r)))
                t_gpu.append(time.perf_counter() - t0)
            launches = (reg.launch_count - before) / runs
        # the CPU side: the unmodified reference where it finishes in seconds and the pow mode is its own,
        # else the matrix-free restatement
        use_ref = oracle.have_reference() and mode == q.POW_VERBATIM and n <= 10
        cpu = oracle.Reference(L, M) if use_ref else oracle.Restatement(L, M)
        want, t_cpu = [], []
        cpu_runs = runs if n <= 7 else min(runs, 5)
        for r in rs[:cpu_runs]:
            t0 = time.perf_counter()
            cpu.reset_register()
            if use_ref:
                cpu.quantum_computation(Cn, a)
                want.append(int(cpu.measure_state_r(r)))
            else:
                cpu.quantum_computation(Cn, a, 1 if mode == q.POW_MODULAR else 0)
                want.append(int(cpu.measure_state(r)))
            t_cpu.append(time.perf_counter() - t0)
        cpu.close()
        cases.append({"case": name, "C": Cn, "a": a, "L": L, "M": M, "qubits": n,
                      "gates": 3 * L + L * (L - 1) // 2,
                      "gpu_ms_per_find_period": 1e3 * float(np.median(t_gpu)), "gpu_runs": runs,
                      "gpu_launches_per_find_period": launches,
                      "cpu_ms_per_find_period": 1e3 * float(np.median(t_cpu)), "cpu_runs": cpu_runs,
                      "cpu_kind": "reference" if use_ref else "port", "cpu_cores": 1,
                      "measured_indices_identical": got[:cpu_runs] == want, "indices": got[:8]})
    if not with_n30:
        return {"find_period": cases, "parity": {"ok": bool(all(c["measured_indices_identical"] for c in cases))}}
    # ---- n = 30: (a) find_period's own sequence reset_register -> quantum_computation, which the engine runs as a
    # closed-form write of the state after the Hadamards and the controlled multiplications + the inverse QFT;
    # (b) quantum_computation on an arbitrary (synthetic) state: Walsh-Hadamard sweeps, the modular exponentiation
    # as one block-local sweep (its GB/s is the roofline entry), inverse QFT
    L, M, Cn, a = 18, 12, 4087, 7                                 # 4087 = 61 * 67 < 2^12
    big_runs = 5
    with q.Register(L, M, device=ranks.local_rank) as reg:
        reg.reset_register(); reg.quantum_computation(Cn, a, q.POW_MODULAR); reg.synchronize()
        reg.set_option(q.OPT_PROFILE, 1)
        reg.profile_reset()
        reg.timer_start()
        for _ in range(big_runs):
            reg.reset_register()
            reg.quantum_computation(Cn, a, q.POW_MODULAR)
        ms = reg.timer_stop()
        prof_reset = reg.profile()
        launches = reg.launch_count
        norm = reg.norm2()
        # measure_state on the state find_period measures (collapses it: rebuilt before each variate)
        measure_ms = {}
        reg.set_option(q.OPT_PROFILE, 0)                           # per-launch events would be timed too
        for r in (0.9, 0.25, 0.6180339887, 0.9):                   # the first call allocates the scan's scratch
            t0 = time.perf_counter()                               # a synchronous call: the host clock is the cost.  (The
            idx = int(reg.measure_state(r))                        # collapse stays pending and the next reset drops it.)
            measure_ms[f"r={r:.3g}"] = round(1e3 * (time.perf_counter() - t0), 3)
            reg.reset_register(); reg.quantum_computation(Cn, a, q.POW_MODULAR)
            reg.norm2()                                            # nothing deferred is left for the next timing
        # the parallel exact scan against the single-CTA sequential scan (the reference's loop, one thread) on this state
        scan_check = {}
        for r in (0.25, 0.6180339887):
            par = int(reg.measure_state(r))
            reg.reset_register(); reg.quantum_computation(Cn, a, q.POW_MODULAR)
            reg.set_option(q.OPT_MEASURE_SEQUENTIAL, 1)
            seq = int(reg.measure_state(r))
            reg.set_option(q.OPT_MEASURE_SEQUENTIAL, 0)
            reg.reset_register(); reg.quantum_computation(Cn, a, q.POW_MODULAR)
            scan_check[f"r={r:.3g}"] = {"parallel": par, "sequential": seq, "same": par == seq}
        reg.set_option(q.OPT_PROFILE, 1)
        idx = int(reg.measure_state(0.6180339887))
        # (b) the general path
        reg.fill_synthetic(SEED)
        reg.scale(1.0 / math.sqrt(reg.norm2()))
        reg.quantum_computation(Cn, a, q.POW_MODULAR)
        reg.synchronize()
        reg.profile_reset()
        reg.timer_start()
        for _ in range(big_runs):
            reg.quantum_computation(Cn, a, q.POW_MODULAR)
        ms_general = reg.timer_stop()
        prof = reg.profile()
        norm_general = reg.norm2()
    per_class = {k: {"launches": v[0], "ms": round(v[1], 4),
                     "GBps": round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None}
                 for k, v in prof.items() if v[0]}
    per_class_reset = {k: {"launches": v[0], "ms": round(v[1], 4),
                           "GBps": round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None}
                       for k, v in prof_reset.items() if v[0]}
    mx = prof["modexp_sweep"]
    mx_gbps = mx[2] / (mx[1] * 1e-3) / 1e9 if mx[1] > 0 else 0.0
    gates30 = 3 * L + L * (L - 1) // 2
    ok = (all(c["measured_indices_identical"] for c in cases) and abs(norm - 1.0) < 1e-10 and abs(norm_general - 1.0) < 1e-10
          and all(v["same"] for v in scan_check.values()))
    line = {"metric": "shor_quantum_computation_gates_per_sec", "value": gates30 * big_runs / (ms * 1e-3), "unit": "gates/s",
            "n_gpus": 1, "steps": big_runs, "warmup": 1, "ms_per_step": ms / big_runs, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 amplitudes, u32 index arithmetic", "data": "synthetic",
            "config": {"workload": f"reset_register + quantum_computation(C={Cn}, a={a}), L={L}, M={M} (n=30, {gates30} gates "
                                   f"per step: {L} H, {L} controlled a^(2^k) mod C, inverse QFT on the L register), "
                                   f"the sequence of find_period (qc_shor.c:922-923)",
                       "qubits": 30, "norm_after": norm, "measured_index": idx,
                       "measure_state_ms": measure_ms, "measure_parallel_vs_sequential_scan": scan_check,
                       "kernels_from_reset": per_class_reset,
                       "general_state": {"ms_per_quantum_computation": ms_general / big_runs, "norm_after": norm_general,
                                         "what": "the same quantum_computation on a synthetic (non-reset) state: "
                                                 "Walsh-Hadamard sweeps + modexp sweep + inverse QFT"}},
            "roofline": {"bound": "hbm", "kernel": "modexp_sweep", "achieved": mx_gbps, "peak": peak, "unit": "GB/s",
                         "frac": mx_gbps / peak, "peak_source": peak_src + " (of measured)", "traffic": None,
                         "launches": mx[0], "avg_launch_ms": mx[1] / max(mx[0], 1),
                         "algorithmic_bytes_per_launch": mx[2] / max(mx[0], 1),
                         "bytes_model": "32 * 2^n * C / 2^M B per launch (read + write the rows f < C of every block); "
                                        "measured on the general-state run"},
            "kernels": per_class, "gpu_launches": int(launches),
            "find_period": cases, "parity": {"ok": bool(ok)}}
    return line


def run_shor(q, ranks, args):
    if ranks.world != 1:
        raise SystemExit("--workload shor is a single-GPU workload")
    line = shor_block(q, ranks, args)
    print(json.dumps(line), flush=True)
    ranks.close()
    if not line["parity"]["ok"]:
        raise SystemExit(3)


def layered_block(q, ranks, args, peak, peak_src, n_override=None):
    """BASELINE configs[3] in the default run: the layered H / C-phase circuit at n = 33 on one GPU,
    issued gate by gate inside qcs_fuse_begin / qcs_fuse_end; checked by applying the inverse circuit.
    (n_override: the CPU dry run of tests/test_bench_flow_cpu.py.)"""
    import numpy as np
    from quantumcomputer_b200.workloads import apply_gates, layered_circuit
    n, layers = n_override or 33, args.layers
    circuit = layered_circuit(n, layers)
    inverse = [g if g[0] == "h" else ("cp", g[1], g[2], -g[3]) for g in reversed(circuit)]
    with q.Register(n, 0, device=ranks.local_rank) as reg:
        apply_tuning(q, reg, args)
        reg.fill_synthetic(SEED)
        reg.scale(1.0 / math.sqrt(reg.norm2()))
        nl = reg.local_states
        probes = [0, 1, nl - 1, nl // 2 + 9, nl // 3, (nl // 7) * 5, min(123456789, nl - 2), nl // 5 + 1]
        before = np.array([reg.get_state(i, 1)[0] for i in probes])

        def one_step():
            with reg.fused():
                apply_gates(reg, circuit)

        steps = max(1, min(args.steps, 3))
        ms, prof, launches, _ = timed_steps(q, reg, ranks, one_step, steps, 1, False)
        norm_out = reg.norm2()
        # undo the (steps + 1) applications: the round trip restores the state
        for _ in range(steps + 1):
            with reg.fused():
                apply_gates(reg, inverse)
        after = np.array([reg.get_state(i, 1)[0] for i in probes])
        err = float(np.max(np.abs(after - before))) * math.sqrt(float(1 << n))
    gates = len(circuit)
    return {"workload": f"layered circuit (BASELINE configs[3]): {layers} layers of H on every qubit + C-phase on "
                        f"(q, (q+1+d) mod n), n={n}, {gates} gates per step, gate by gate inside qcs_fuse_begin/end",
            "qubits": n, "gates_per_step": gates, "steps": steps, "warmup": 1, "ms_per_step": ms / steps,
            "gates_per_sec": gates * steps / (ms * 1e-3), "gpu_launches": int(launches),
            "parity": {"ok": bool(err <= 1e-11 and abs(norm_out - 1.0) < 1e-10),
                       "round_trip_max_rel_err": err, "norm_after": norm_out,
                       "what": "circuit^steps followed by its inverse^steps restores 8 probed amplitudes"},
            "roofline": roofline_from_profile(prof, n, 1, False, peak, peak_src),
            "kernels": {kk: {"launches": v[0], "ms": round(v[1], 4),
                             "GBps": round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None}
                        for kk, v in prof.items() if v[0]}}


def run_ours(args):
    import quantumcomputer_b200 as q

    ranks = Ranks(q)
    rank, world, local_rank = ranks.rank, ranks.world, ranks.local_rank
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    if args.workload == "shor":
        return run_shor(q, ranks, args)

    p = int(math.log2(world))
    if args.workload == "layered":
        # BASELINE configs[3]: random Hadamard / controlled-phase layered circuit, n = 33 on one GPU
        from quantumcomputer_b200.workloads import apply_gates, layered_circuit
        n = args.qubits if args.qubits else 33 + p
        circuit = layered_circuit(n, args.layers)
        gates = len(circuit)
        workload = (f"layered circuit (BASELINE configs[3]): {args.layers} layers of H on every qubit + C-phase on "
                    f"(q, (q+1+d) mod n), n={n}, {gates} gates per step, issued gate by gate through "
                    f"qcs_hadamard_gate / qcs_c_phase_shift_gate inside qcs_fuse_begin/end")
    else:
        n = args.qubits if args.qubits else 30 + p
        circuit = None
        gates = qft_gate_count(n)
        workload = (f"inverse_QFT over all n={n} qubits of a synthetic random state "
                    f"(BASELINE configs[2]; {gates} gates per step)")
    reg = q.Register(n, 0, device=local_rank, rank=rank, world_size=world, comm_id=ranks.comm_id())
    apply_tuning(q, reg, args)

    if circuit is not None:
        args.no_cpu_baseline = True                  # the CPU arm times the headline (iqft) workload
        args.no_north_star = True
        if 16 * reg.local_states > 32 * 2 ** 30:
            args.no_e2e = True                       # no 128 GiB pinned host mirror of an n = 33 state

    def one_step():
        if circuit is None:
            reg.inverse_QFT()
        else:
            with reg.fused():
                apply_gates(reg, circuit)

    # ---- in-run parity on this register (all ranks), before anything is timed
    parity = None
    if circuit is None and not args.no_parity:
        parity = parity_block(q, reg, n, ranks)
        if not parity["ok"] and rank == 0:
            # said at once (should a later step hang) and again in the line, which then carries "error" and
            # the process exits 3: a number next to a failed check is not a result
            print(json.dumps({"error": "parity check failed", "parity": parity}), file=sys.stderr, flush=True)

    # synthetic state, generated on the device, normalised
    reg.fill_synthetic(SEED)
    reg.scale(1.0 / math.sqrt(reg.norm2()))
    norm_in = reg.norm2()

    # ---- device-resident throughput ("value")
    ms, prof, launches, clocks = timed_steps(q, reg, ranks, one_step, args.steps, args.warmup, True)
    norm_out = reg.norm2()

    # ---- end to end through the C ABI with host buffers ("e2e")
    e2e = None
    if not args.no_e2e:
        local = reg.local_states
        pinned = q.PinnedBuffer(2 * local, device=local_rank)       # pages next to this rank's GPU
        reg.get_state(0, local, out=pinned.array)          # a normalised host-resident input
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        ranks.barrier(reg)
        t_ms, h2d_ms = 0.0, 0.0
        for step_i in range(e2e_steps):
            reg.timer_start()
            reg.set_state_async(pinned.array)               # H2D from pinned memory
            one_step()
            # the step's result, as in find_period (qc_shor.c:923-928): the measured index (8 bytes D2H)
            result = reg.measure_state(((step_i * 2654435761 + 12345) % 2 ** 32) / 2.0 ** 32)
            t_ms += reg.timer_stop()
        # the upload alone, for the PCIe share of the number above
        reg.timer_start()
        reg.set_state_async(pinned.array)
        h2d_ms = ranks.max(reg.timer_stop())
        t_ms = ranks.max(t_ms)
        e2e = {"value": gates * e2e_steps / (t_ms * 1e-3), "unit": "gates/s",
               "h2d_bytes_per_step": int(16 * local * world), "d2h_bytes_per_step": 8 * world,
               "steps": e2e_steps, "ms_per_step": t_ms / e2e_steps,
               "h2d_ms": h2d_ms, "h2d_GBps_per_gpu": 16 * local / (h2d_ms * 1e-3) / 1e9,
               "result": "measure_state index read back each step (qc_shor.c:928)", "last_result": result}
        pinned.close()

    peak, peak_src = measured_peaks()
    line = None
    if rank == 0:
        per_class = {k: {"launches": v[0], "ms": round(v[1], 4),
                         "GBps": round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None}
                     for k, v in prof.items() if v[0]}
        gates_per_s = gates * args.steps / (ms * 1e-3)
        # Weak scaling grows n with the GPU count, and one gate on n+1 qubits is twice the work, so for
        # N > 1 the value is the n = 30-equivalent gate rate (gates/s x 2^(n-30)): extensive in the GPU
        # count, identical to gates/s at N = 1 (n = 30).
        value = gates_per_s * 2.0 ** (n - 30) if (circuit is None and world > 1) else gates_per_s
        line = {
            "metric": "qft_gates_per_sec" if circuit is None else "layered_circuit_gates_per_sec",
            "value": value, "unit": "gates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload,
                       "qubits": n, "gates_per_step": gates, "state_bytes_per_gpu": int(16 * reg.local_states),
                       "value_definition": ("gates/s at n=30" if world == 1 or circuit is not None else
                                            f"n=30-equivalent gates/s = gates/s at n={n} x 2^(n-30) "
                                            f"(= amplitude-gate updates/s / 2^30), so that it can scale with the GPU count"),
                       "gates_per_sec_at_n": gates_per_s,
                       "fusion": int(reg.get_option(q.OPT_FUSION)), "parallelism": f"top {p} qubits global",
                       "pipe_shape": int(reg.get_option(q.OPT_PIPE_SHAPE)),
                       "min_run_bits": int(reg.get_option(q.OPT_MIN_RUN_BITS)),
                       "l2": f"state ({16 * reg.local_states / 2 ** 30:.0f} GiB per GPU) is far larger than the "
                             f"126 MB L2; no flush needed",
                       "norm_before": norm_in, "norm_after": norm_out},
            "work_rate": {"value": gates * float(1 << n) * args.steps / (ms * 1e-3), "unit": "amplitude-gate updates/s"},
            "roofline": roofline_from_profile(prof, n, world, circuit is None, peak, peak_src),
            "kernels": per_class,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if parity is not None:
            line["parity"] = parity
            if not parity["ok"]:
                line["error"] = "parity check failed"
        if e2e is not None:
            line["e2e"] = e2e
    reg.close()

    # ---- BASELINE's multi-GPU metric at its own sizes, same run
    if circuit is None and not args.no_north_star and not args.qubits:
        ns = north_star_block(q, ranks, args, peak, peak_src)
        if rank == 0 and ns is not None:
            line["north_star"] = ns
            if not ns["parity"]["ok"]:
                line["error"] = "north_star parity check failed"

    # ---- the other single-GPU configs of BASELINE.json, in the same run (N = 1 only)
    if circuit is None and world == 1 and not args.qubits and not args.no_configs:
        blocks = {}
        try:
            blocks["shor"] = shor_block(q, ranks, args, with_n30=False)
            blocks["layered_n33"] = layered_block(q, ranks, args, peak, peak_src)
        except Exception as exc:                       # never take the headline line down
            blocks["error"] = repr(exc)
        line["configs"] = blocks
        for name in ("shor", "layered_n33"):
            if name in blocks and not blocks[name]["parity"]["ok"]:
                line["error"] = f"{name} parity check failed"

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    failed = rank == 0 and "error" in line
    ranks.close()
    if failed:
        raise SystemExit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--qubits", type=int, default=0, help="override n (default 30 + log2(gpus))")
    ap.add_argument("--no-fusion", action="store_true", help="gate-by-gate reference-order kernels")
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--prefetch", type=int, default=-1, help="L2 prefetch distance in tiles (-1: library default)")
    ap.add_argument("--workload", choices=["iqft", "layered", "shor"], default="iqft",
                    help="iqft: BASELINE configs[2] (default, the headline metric); layered: configs[3]; "
                         "shor: configs[0] / [1] wall time per find_period beside the reference's + an n = 30 quantum_computation")
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--pipe-shape", type=int, default=-1)
    ap.add_argument("--min-run-bits", type=int, default=0)
    ap.add_argument("--global-run-bits", type=int, default=0)
    ap.add_argument("--overlap-slices", type=int, default=-1)
    ap.add_argument("--global-sms", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity block")
    ap.add_argument("--no-north-star", action="store_true", help="skip the n = 33/34/35 block")
    ap.add_argument("--no-configs", action="store_true", help="skip the Shor cfg1/cfg2 and layered n = 33 blocks (N = 1)")
    ap.add_argument("--north-star-steps", type=int, default=5)
    ap.add_argument("--l2-pair", type=int, default=-1)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
