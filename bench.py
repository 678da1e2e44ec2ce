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
    committed `ncu --set full` capture of the same workload (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_qft_sweep_tma_t12_n30.csv")
    if n != 30 or not os.path.exists(path):
        return None, None
    try:
        import csv
        with open(path) as f:
            rows = list(csv.reader(f))
        hdr = rows[0]
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        per = [float(r[rd]) + float(r[wr]) for r in rows[2:] if len(r) > wr]
        return 1e9 * sum(per) / len(per), os.path.relpath(path, ROOT)
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


def pick_reference_n(budget_s, kind_is_reference=True):
    # measured cost of one reference inverse_QFT grows ~4.5x per qubit
    table = {8: 0.015, 9: 0.07, 10: 0.32, 11: 1.6, 12: 6.8, 13: 31.0, 14: 146.0}
    best = 8
    for n, t in table.items():
        if t <= budget_s:
            best = max(best, n)
    return best


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    n = args.qubits if args.qubits else pick_reference_n(150.0 / max(total, 1))
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
              f"the reference builds each gate as a 4^n-scan COO matrix, so n=30 is out of reach")
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


def cpu_baseline(budget_s=25.0):
    n = pick_reference_n(budget_s / 1.5)
    kind, dt = reference_iqft_seconds(n)
    gates = qft_gate_count(n)
    return {"value": gates / dt, "unit": "gates/s", "cores": 1, "kind": kind,
            "sample": f"one inverse_QFT over all n={n} qubits ({gates} gates, {dt:.2f} s) of the synthetic "
                      f"state, seed {SEED}; serial program, 1 thread of {os.cpu_count()} host cores"}


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
def run_ours(args):
    import quantumcomputer_b200 as q

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world

    dist = None
    comm_id = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        ids = [q.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm_id = ids[0]

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
    reg = q.Register(n, 0, device=local_rank, rank=rank, world_size=world, comm_id=comm_id)
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

    if circuit is not None:
        args.no_cpu_baseline = True                  # the CPU arm times the headline (iqft) workload
        if 16 * reg.local_states > 32 * 2 ** 30:
            args.no_e2e = True                       # no 128 GiB pinned host mirror of an n = 33 state

    def one_step():
        if circuit is None:
            reg.inverse_QFT()
        else:
            with reg.fused():
                apply_gates(reg, circuit)

    def barrier():
        reg.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # synthetic state, generated on the device, normalised
    reg.fill_synthetic(SEED)
    reg.scale(1.0 / math.sqrt(reg.norm2()))
    norm_in = reg.norm2()

    for _ in range(args.warmup):
        one_step()
    barrier()

    # ---- device-resident throughput ("value")
    reg.set_option(q.OPT_PROFILE, 1)
    reg.profile_reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    reg.timer_start()
    for _ in range(args.steps):
        one_step()
    ms = reg.timer_stop()
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(ms)
    prof = reg.profile()
    launches = reg.launch_count
    reg.set_option(q.OPT_PROFILE, 0)
    norm_out = reg.norm2()

    # ---- end to end through the C ABI with host buffers ("e2e")
    e2e = None
    if not args.no_e2e:
        local = reg.local_states
        pinned = q.PinnedBuffer(2 * local)
        reg.get_state(0, local, out=pinned.array)          # a normalised host-resident input
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        barrier()
        t_ms = 0.0
        for step_i in range(e2e_steps):
            reg.timer_start()
            reg.set_state_async(pinned.array)               # H2D from pinned memory
            one_step()
            # the step's result, as in find_period (qc_shor.c:923-928): the measured index (8 bytes D2H)
            result = reg.measure_state(((step_i * 2654435761 + 12345) % 2 ** 32) / 2.0 ** 32)
            t_ms += reg.timer_stop()
        t_ms = max_over_ranks(t_ms)
        e2e = {"value": gates * e2e_steps / (t_ms * 1e-3), "unit": "gates/s",
               "h2d_bytes_per_step": int(16 * local * world), "d2h_bytes_per_step": 8 * world,
               "steps": e2e_steps, "ms_per_step": t_ms / e2e_steps,
               "result": "measure_state index read back each step (qc_shor.c:928)", "last_result": result}
        pinned.close()

    if rank == 0:
        peak, peak_src = measured_peaks()
        # dominant kernel class = most device time
        dom = max((k for k in prof if prof[k][0] > 0), key=lambda k: prof[k][1])
        d_launches, d_ms, d_bytes = prof[dom]
        achieved = d_bytes / (d_ms * 1e-3) / 1e9 if d_ms > 0 else 0.0
        kernel_ms = sum(v[1] for v in prof.values())
        traffic, traffic_src = profiled_traffic(n) if (dom == "tile_sweep" and circuit is None and world == 1) else (None, None)
        if dom == "global_sweep":
            # multi-GPU: the sweep over the global qubits is bound by NVLink; the denominator is the
            # measured peer copy per direction per GPU stated in B200_PROFILING.md (nominal 900)
            nv_peak = 770.0
            roofline = {"bound": "nvlink", "kernel": dom, "achieved": achieved, "peak": nv_peak, "unit": "GB/s",
                        "frac": achieved / nv_peak,
                        "peak_source": "B200_PROFILING.md measured peer copy per direction per GPU (of measured)",
                        "traffic": None, "launches": d_launches, "avg_launch_ms": d_ms / d_launches,
                        "algorithmic_bytes_per_launch": d_bytes / d_launches,
                        "share_of_kernel_time": d_ms / kernel_ms if kernel_ms else None,
                        "bytes_model": "NVLink bytes per direction per GPU and launch: 2*(P-1)/P * 16*2^n_local "
                                       "(remote reads + remote writes of the rank's share of the tiles)",
                        "local_sweeps_GBps": (prof["tile_sweep"][2] / (prof["tile_sweep"][1] * 1e-3) / 1e9
                                              if prof["tile_sweep"][1] > 0 else None)}
        else:
            roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "peak_source": peak_src + " (of measured)",
                        "traffic": traffic, "traffic_source": traffic_src, "launches": d_launches,
                        "avg_launch_ms": d_ms / d_launches,
                        "algorithmic_bytes_per_launch": d_bytes / d_launches,
                        "share_of_kernel_time": d_ms / kernel_ms if kernel_ms else None,
                        "bytes_model": CLASS_BYTES_NOTE.get(dom, "")}
        per_class = {k: {"launches": v[0], "ms": round(v[1], 4),
                         "GBps": round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None}
                     for k, v in prof.items() if v[0]}
        line = {
            "metric": "qft_gates_per_sec" if circuit is None else "layered_circuit_gates_per_sec",
            "value": gates * args.steps / (ms * 1e-3), "unit": "gates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload,
                       "qubits": n, "gates_per_step": gates, "state_bytes_per_gpu": int(16 * reg.local_states),
                       "fusion": int(reg.get_option(q.OPT_FUSION)), "parallelism": f"top {p} qubits global",
                       "pipe_shape": int(reg.get_option(q.OPT_PIPE_SHAPE)),
                       "min_run_bits": int(reg.get_option(q.OPT_MIN_RUN_BITS)),
                       "l2": f"state ({16 * reg.local_states / 2 ** 30:.0f} GiB per GPU) is far larger than the "
                             f"126 MB L2; no flush needed",
                       "norm_before": norm_in, "norm_after": norm_out},
            # gates/s cannot scale with the GPU count when n grows with it (a gate on n+1 qubits is
            # twice the work): amplitude updates per second = gates * 2^n / time is the rate that can
            "work_rate": {"value": gates * float(1 << n) * args.steps / (ms * 1e-3), "unit": "amplitude-gate updates/s"},
            "roofline": roofline, "kernels": per_class,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    reg.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
    ap.add_argument("--workload", choices=["iqft", "layered"], default="iqft",
                    help="iqft: BASELINE configs[2] (default, the headline metric); layered: configs[3]")
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--pipe-shape", type=int, default=-1)
    ap.add_argument("--min-run-bits", type=int, default=0)
    ap.add_argument("--global-run-bits", type=int, default=0)
    ap.add_argument("--overlap-slices", type=int, default=-1)
    ap.add_argument("--global-sms", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
