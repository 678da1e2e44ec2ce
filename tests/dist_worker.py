"""Multi-GPU parity worker: run under torch.distributed.run, one rank per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_worker.py

Every sharded result is gathered to rank 0 and compared there with a
single-GPU register running the same calls (which the single-GPU tests pin to
the oracle).  Prints DIST_OK on success."""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantumcomputer_b200 as q  # noqa: E402

TOL = 1e-12


def gather_state(reg, rank, world):
    local = torch.from_numpy(reg.get_state().view(np.float64).copy()).cuda()
    parts = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
    dist.gather(local, parts, dst=0)
    if rank == 0:
        return torch.cat(parts).cpu().numpy().view(np.complex128)
    return None


def main():
    import faulthandler
    faulthandler.dump_traceback_later(100, exit=True)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    failures = []

    def fresh_id():
        # one NCCL unique id per communicator (per sharded register)
        ids = [q.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return ids[0]

    def check(name, got, want, exact=False):
        if rank != 0:
            return
        if exact:
            ok = bool(np.all(got.view(np.float64) == want.view(np.float64)))
            err = 0.0 if ok else float(np.linalg.norm(got - want) / np.linalg.norm(want))
        else:
            err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
            ok = err <= TOL
        print(f"[dist] {name}: {'ok' if ok else 'FAIL'} (err {err:.2e})", flush=True)
        if not ok:
            failures.append(name)

    for (L, M) in [(18, 0), (12, 5), (21, 0), (20, 5)]:
        n = L + M
        sh = q.Register(L, M, device=local_rank, rank=rank, world_size=world, comm_id=fresh_id())
        if rank == 0:
            print(f"[dist] n={n} M={M}: peer memory {'on' if sh.peer_memory else 'off'} "
                  f"(shard {16 * sh.local_states / 2 ** 20:.1f} MiB)", flush=True)
        single = q.Register(L, M, device=local_rank) if rank == 0 else None

        def both(fn):
            fn(sh)
            if single is not None:
                fn(single)

        def prep(r):
            r.fill_synthetic(77 + n)
        both(prep)
        s = sh.norm2()
        both(lambda r: r.scale(1.0 / math.sqrt(s)))
        s2 = sh.norm2()                     # collective: every rank calls it
        if rank == 0:
            assert abs(single.norm2() - s2) < 1e-13

        # gate by gate, including Hadamards on the global qubits (pairwise exchange)
        both(lambda r: r.set_option(q.OPT_FUSION, 0))
        def gates(r):
            r.hadamard_gate(n - 1)
            r.c_phase_shift_gate(n - 1, 2, 0.37)
            r.hadamard_gate(3)
            r.c_phase_shift_gate(n - 2, n - 1, -1.1)
            if world > 2:
                r.hadamard_gate(n - 2)
            if M:
                r.c_amodc_gate(21, 4, n - 1)
                r.c_amodc_gate(21, 5, M)
        both(gates)
        got = gather_state(sh, rank, world)
        check(f"n={n} M={M} gate-by-gate", got, single.get_state() if rank == 0 else None, exact=True)

        # arbitrary gates with global target / control qubits
        U1 = np.array([[0.6, 0.8j], [0.8j, 0.6]])
        U2 = np.array([[np.exp(0.3j), 0.0], [0.0, np.exp(-1.1j)]]) @ (np.array([[1, 1], [1, -1]]) / math.sqrt(2.0))
        def general(r):
            r.apply_gate(n - 1, U1)
            r.apply_controlled_gate(1, n - 1, U2)
            r.apply_controlled_gate(n - 1, 4, U1)
            if world > 2:
                r.apply_controlled_gate(n - 1, n - 2, U2)
        both(general)
        got = gather_state(sh, rank, world)
        check(f"n={n} M={M} arbitrary gates", got, single.get_state() if rank == 0 else None)

        # deferred gate stream across the global qubits: layered circuit + a few stray gates
        from quantumcomputer_b200.workloads import apply_gates, layered_circuit
        stream = layered_circuit(n, 2) + [("h", n - 1), ("cp", n - 1, 0, 0.4), ("h", n - 1), ("h", 2)]
        def fused_stream(r):
            with r.fused():
                apply_gates(r, stream)
        both(fused_stream)
        got = gather_state(sh, rank, world)
        check(f"n={n} M={M} fused gate stream", got, single.get_state() if rank == 0 else None)

        # sampling without collapse
        rs = [0.0, 0.2, 0.5, 0.93, 1.0]
        a = sh.sample_states(rs)
        if rank == 0:
            b = single.sample_states(rs)
            print(f"[dist] n={n} sample_states: {a} vs {b}", flush=True)
            if a != b:
                failures.append("sample_states")

        # fused inverse / forward QFT across the global qubits
        both(lambda r: r.set_option(q.OPT_FUSION, 1))
        both(lambda r: r.inverse_QFT())
        got = gather_state(sh, rank, world)
        check(f"n={n} M={M} fused inverse_QFT", got, single.get_state() if rank == 0 else None)
        both(lambda r: r.QFT())
        got = gather_state(sh, rank, world)
        check(f"n={n} M={M} fused QFT", got, single.get_state() if rank == 0 else None)

        # measurement: same index everywhere, bit-exact with the single-GPU scan
        for rr in (0.0, 0.31, 0.77, 0.999999, 1.5):
            if rank == 0:
                keep = single.get_state().copy()
            a = sh.measure_state(rr)
            if rank == 0:
                b = single.measure_state(rr)
                print(f"[dist] n={n} measure r={rr}: {a} vs {b}", flush=True)
                if a != b:
                    failures.append(f"measure {rr}")
                single.set_state(keep)
            # restore the sharded state from the single-GPU copy
            obj = [keep if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            nl = sh.local_states
            sh.set_state(obj[0][rank * nl:(rank + 1) * nl])

        if M:
            both(lambda r: r.reset_register())
            both(lambda r: r.quantum_computation(21, 2, q.POW_MODULAR))
            got = gather_state(sh, rank, world)
            check(f"n={n} M={M} fused quantum_computation (closed form from reset)", got, single.get_state() if rank == 0 else None)
            # ... and from a state that is not the reset state: sharded Walsh-Hadamard sweeps + modexp sweep + inverse QFT
            both(lambda r: r.quantum_computation(21, 4, q.POW_MODULAR))
            got = gather_state(sh, rank, world)
            check(f"n={n} M={M} fused quantum_computation (general path)", got, single.get_state() if rank == 0 else None)
            # a measurement, then the next trial's reset + computation (deferred collapse and reset on every shard)
            a = sh.measure_state(0.4242)
            if rank == 0:
                b = single.measure_state(0.4242)
                if a != b:
                    failures.append("measure after quantum_computation")
            both(lambda r: r.reset_register())
            both(lambda r: r.quantum_computation(21, 5, q.POW_MODULAR))
            got = gather_state(sh, rank, world)
            check(f"n={n} M={M} second trial after measure_state", got, single.get_state() if rank == 0 else None)
        sh.close()
        if single is not None:
            single.close()

    flag = torch.tensor([len(failures)], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()):
        if rank == 0:
            print("DIST_FAIL", failures, flush=True)
        sys.exit(1)
    if rank == 0:
        print("DIST_OK", flush=True)


if __name__ == "__main__":
    main()
