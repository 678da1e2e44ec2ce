"""CPU, gloo, world_size 2 and 4: the two sharded inverse-QFT schedules restated with numpy +
torch.distributed and checked against the oracle's gate-by-gate inverse_QFT of the whole
register.  This pins the index / twiddle / tile-share bookkeeping of the multi-GPU paths
without a GPU.

  * peer-memory schedule (csrc/peer.cu + qcs_fused_sweeps_sharded): one tile sweep over the
    stages [split, n) on the stitched array, every rank taking its 1/world share of the tiles
    (reads and writes of the peers' rows), a barrier, then the stages [lo, split) on each shard;
  * exchange schedule (csrc/dist.cu, the fallback without peer memory): exchange in -> one
    sweep over the global qubits with y_const -> exchange out -> local stages."""
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


# ---------------------------------------------------------------------------
# peer-memory schedule
# ---------------------------------------------------------------------------
def _apply_stages(idx, v, lo, stage_lo, stage_hi):
    """The stages stage_hi-1 .. stage_lo of the reference circuit (qc_shor.c:682-689) on the
    amplitudes v at basis states idx (a set closed under flipping the stage bits): Hadamard on
    bit l, then exp(i pi (x mod 2^j) / 2^j) on the amplitudes whose bit l is 1, x = idx >> lo."""
    order = np.argsort(idx)
    idx, v = idx[order], v[order].copy()
    pos = {int(i): k for k, i in enumerate(idx)}
    for l in range(stage_hi - 1, stage_lo - 1, -1):
        bit = 1 << l
        zero = np.array([k for k, i in enumerate(idx) if not (int(i) & bit)])
        one = np.array([pos[int(idx[k]) | bit] for k in zero])
        a, b = v[zero].copy(), v[one].copy()
        v[zero] = (a + b) / math.sqrt(2.0)
        v[one] = (a - b) / math.sqrt(2.0)
        j = l - lo
        x = (idx[one] >> lo) & ((1 << j) - 1)
        v[one] *= np.exp(1j * math.pi * x / float(1 << j))
    return idx, v


def _peer_worker(rank, world, port, L, M, a_glob, g, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import oracle
    n = L + M
    p = int(math.log2(world))
    n_local = n - p
    lo = M
    split = n - g                       # global sweep: stages [split, n), tile = [0, a_glob) U [split, n)
    assert g >= p and split >= a_glob and split >= lo
    full = oracle.Restatement(L, M)
    full.fill_synthetic(7)
    full.scale(1.0 / math.sqrt(full.norm2()))
    state = full.get_state().copy()
    shard = state[rank << n_local:(rank + 1) << n_local].copy()

    # "peer memory": every rank can address the whole register
    parts = [torch.empty(2 << n_local, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(shard.copy().view(np.float64)))
    amp_all = np.concatenate([t.numpy().view(np.complex128) for t in parts])

    # this rank's share of the tiles of the global sweep
    n_tiles = 1 << (n - a_glob - g)
    share = n_tiles // world
    gap = split - a_glob
    my_idx, my_val = [], []
    e = np.arange(1 << (a_glob + g), dtype=np.int64)
    spread = (e & ((1 << a_glob) - 1)) | ((e >> a_glob) << split)
    for tix in range(rank * share, (rank + 1) * share):
        base = (tix & ((1 << gap) - 1)) << a_glob          # no index bits above the tile (g_hi = n)
        idx = base | spread
        i2, v2 = _apply_stages(idx, amp_all[idx], lo, split, n)
        my_idx.append(i2)
        my_val.append(v2)
    my_idx = np.concatenate(my_idx)
    my_val = np.concatenate(my_val)
    # the writes land in the owners' shards (remote stores); here: gather everybody's writes
    idx_parts = [torch.empty(my_idx.size, dtype=torch.int64) for _ in range(world)]
    val_parts = [torch.empty(2 * my_val.size, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(idx_parts, torch.from_numpy(my_idx))
    dist.all_gather(val_parts, torch.from_numpy(my_val.copy().view(np.float64)))
    written = np.zeros(1 << n, dtype=np.int64)
    for ip, vp in zip(idx_parts, val_parts):
        amp_all[ip.numpy()] = vp.numpy().view(np.complex128)
        written[ip.numpy()] += 1
    assert np.all(written == 1), "the ranks' tile shares must write every amplitude exactly once"
    shard = amp_all[rank << n_local:(rank + 1) << n_local].copy()

    # local stages [lo, split) on the shard: blocks of 2^split amplitudes are independent registers
    if split > lo:
        blk = oracle.Restatement(split - lo, lo)
        for b in range(1 << (n_local - split)):
            blk.set_state(shard[b << split:(b + 1) << split])
            blk.inverse_QFT()
            shard[b << split:(b + 1) << split] = blk.get_state()

    gathered = [torch.empty(2 << n_local, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(shard.copy().view(np.float64)))
    if rank == 0:
        got = np.concatenate([t.numpy().view(np.complex128) for t in gathered])
        full.set_state(state)
        full.inverse_QFT()
        want = full.get_state()
        ret.put(float(np.linalg.norm(got - want) / np.linalg.norm(want)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,L,M,a_glob,g", [(2, 9, 0, 2, 3), (2, 7, 3, 3, 2), (4, 10, 0, 2, 3), (4, 8, 2, 1, 2)])
def test_peer_memory_sharded_inverse_qft_schedule(oracle_built, world, L, M, a_glob, g):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29700 + world * 10 + M
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, L, M, a_glob, g, ret)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(180)
        assert pr.exitcode == 0
    err = ret.get(timeout=5)
    assert err <= 1e-12, err


# ---------------------------------------------------------------------------
# overlapped schedule: slices of the global sweep vs slices of the strided local sweeps
# ---------------------------------------------------------------------------
def _tile_number(t, pos, bits, val):
    """csrc/qft_common.cuh tile_number(): deposit the slice value into the tile number"""
    low = t & ((1 << pos) - 1)
    return ((t >> pos) << (pos + bits)) | (val << pos) | low


def _tile_indices(tix, a, g_lo, g_hi):
    """amplitudes of tile `tix` of a strided sweep with tile bits [0,a) U [g_lo,g_hi)"""
    gap = g_lo - a
    base = ((tix >> gap) << g_hi) | ((tix & ((1 << gap) - 1)) << a)
    e = np.arange(1 << (a + g_hi - g_lo), dtype=np.int64)
    return base | (e & ((1 << a) - 1)) | ((e >> a) << g_lo)


@pytest.mark.parametrize("n,p,a_glob,g,sb,locals_", [
    (16, 1, 4, 3, 2, [(2, 10, 13), (3, 7, 10)]),
    (17, 2, 5, 4, 1, [(3, 9, 13)]),
    (18, 3, 4, 3, 2, [(4, 11, 15), (2, 6, 11)]),
])
def test_overlap_slices_partition_and_dependencies(n, p, a_glob, g, sb, locals_):
    """qcs_fused_sweeps_sharded, overlapped schedule: for every slice j the global-sweep tiles of
    slice j (all ranks' shares) and the tiles of slice j of each strided local sweep (all ranks,
    local coordinates) cover the same amplitudes, and the slices partition the register."""
    P = 1 << p
    n_local = n - p
    split = n - g
    K = 1 << sb
    seen = np.zeros(1 << n, dtype=np.int64)
    for j in range(K):
        glob = np.zeros(1 << n, dtype=bool)
        n_tiles = 1 << (n - a_glob - g)
        share = (n_tiles >> p) >> sb
        for r in range(P):
            for t in range(r * share, (r + 1) * share):
                glob[_tile_indices(_tile_number(t, 0, sb, j), a_glob, split, n)] = True
        seen[glob] += 1
        for (a1, g_lo, g_hi) in locals_:
            assert a1 <= a_glob and g_lo >= a_glob + sb and g_hi <= split
            loc = np.zeros(1 << n, dtype=bool)
            n_tiles_local = (1 << (n_local - a1 - (g_hi - g_lo))) >> sb
            for r in range(P):
                for t in range(n_tiles_local):
                    idx = _tile_indices(_tile_number(t, a_glob - a1, sb, j), a1, g_lo, g_hi)
                    assert idx.max() < (1 << n_local)
                    loc[(r << n_local) | idx] = True
            assert np.array_equal(glob, loc), (j, a1, g_lo, g_hi)
    assert np.all(seen == 1)
