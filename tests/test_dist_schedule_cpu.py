"""CPU, gloo, world_size 2 and 4: the sharded inverse-QFT schedule of csrc/dist.cu
(exchange in -> one sweep over the global qubits with y_const -> exchange out ->
local inverse QFT on each shard) restated with numpy + torch.distributed, checked
against the oracle's gate-by-gate inverse_QFT of the whole register.  This pins
the index/twiddle bookkeeping of the multi-GPU path without a GPU."""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


def _bitrev(k, r):
    out = 0
    for b in range(r):
        if (k >> b) & 1:
            out |= 1 << (r - 1 - b)
    return out


def _top_sweep(cols, p, y, j):
    """cols: array [P, m] (slot = value of the global qubits); radix-2^p DIF over the
    slots, external twiddle exp(i pi y k / 2^j), scale 2^(-p/2)."""
    from qft_plan_prototype import dft_dif
    P = 1 << p
    x = dft_dif([cols[d] for d in range(P)], p)
    out = np.empty_like(cols)
    w = np.exp(1j * math.pi * y / float(1 << j))
    for d in range(P):
        out[d] = x[d] * w ** _bitrev(d, p) * (2.0 ** (-p / 2))
    return out


def _worker(rank, world, port, L, M, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import oracle
    n = L + M
    p = int(math.log2(world))
    n_local = n - p
    q = n_local - p
    lo = M
    full = oracle.Restatement(L, M)
    full.fill_synthetic(5)
    full.scale(1.0 / math.sqrt(full.norm2()))
    state = full.get_state().copy()
    shard = state[rank << n_local:(rank + 1) << n_local].copy()

    # exchange in: piece (r; s, m) -> rank s; afterwards cols[s'] = x(s'; r, :)
    B = 1 << q
    sendbuf = [torch.from_numpy(shard[s * B:(s + 1) * B].copy().view(np.float64)) for s in range(world)]
    recvbuf = [torch.empty_like(t) for t in sendbuf]
    reqs = []
    for s in range(world):
        if s == rank:
            recvbuf[s].copy_(sendbuf[s])
        else:
            reqs.append(dist.isend(sendbuf[s], s))
            reqs.append(dist.irecv(recvbuf[s], s))
    for rq in reqs:
        rq.wait()
    cols = np.stack([t.numpy().view(np.complex128) for t in recvbuf])
    # one sweep over the global qubits; y = register bits below them: m (physical) and r (held by the rank)
    m = np.arange(B, dtype=np.int64)
    y = (m >> lo) + (rank << (q - lo))
    cols = _top_sweep(cols, p, y.astype(np.float64), (n - 1) - lo)
    # exchange out
    sendbuf = [torch.from_numpy(cols[s].copy().view(np.float64)) for s in range(world)]
    reqs = []
    for s in range(world):
        if s == rank:
            recvbuf[s].copy_(sendbuf[s])
        else:
            reqs.append(dist.isend(sendbuf[s], s))
            reqs.append(dist.irecv(recvbuf[s], s))
    for rq in reqs:
        rq.wait()
    shard = np.concatenate([t.numpy().view(np.complex128) for t in recvbuf])
    # local stages: inverse QFT on qubits [lo, n_local) of the shard
    loc = oracle.Restatement(n_local - lo, lo)
    loc.set_state(shard)
    loc.inverse_QFT()
    shard = loc.get_state()

    gathered = [torch.empty(2 << n_local, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(shard.copy().view(np.float64)))
    if rank == 0:
        got = np.concatenate([g.numpy().view(np.complex128) for g in gathered])
        full.set_state(state)
        full.inverse_QFT()
        want = full.get_state()
        ret.put(float(np.linalg.norm(got - want) / np.linalg.norm(want)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,L,M", [(2, 8, 0), (2, 6, 3), (4, 9, 0), (4, 7, 2)])
def test_sharded_inverse_qft_schedule(oracle_built, world, L, M):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29600 + world * 10 + M
    procs = [ctx.Process(target=_worker, args=(r, world, port, L, M, ret)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    err = ret.get(timeout=5)
    assert err <= 1e-12, err
