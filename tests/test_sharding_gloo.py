"""CPU tests of the N>1 plumbing with the gloo backend (world_size 2): field / event partitioning,
the table all-gather layout and the event-result gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mceik_b200 import sharding


def test_block_range_partitions_everything():
    for n in (0, 1, 7, 16, 100, 128):
        for w in (1, 2, 3, 8):
            r = [sharding.block_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert list(sharding.shard_fields(10, 4, 1)) == [3, 4, 5]


def test_shard_events_csr():
    obs_ptr = np.array([0, 3, 3, 8, 10, 15])
    lo, hi, local, p0, p1 = sharding.shard_events(obs_ptr, 2, 1)
    assert (lo, hi, p0, p1) == (3, 5, 8, 15) and list(local) == [0, 2, 7]
    lo, hi, local, p0, p1 = sharding.shard_events(obs_ptr, 2, 0)
    assert (lo, hi, p0, p1) == (0, 3, 0, 8) and list(local) == [0, 3, 3, 8]


def _worker(rank, world, port, nfields, nevents, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ld = 40
        full = torch.arange(nfields * ld, dtype=torch.float32).reshape(nfields, ld)
        mine = sharding.shard_fields(nfields, world, rank)
        out = sharding.all_gather_tables(full[mine[0]:mine[-1] + 1].clone() if len(mine) else full[:0].clone(), nfields)
        ok1 = bool(torch.equal(out, full))
        ev = torch.arange(nevents, dtype=torch.float64) * 1.5
        lo, hi = sharding.block_range(nevents, world, rank)
        got = sharding.gather_event_results(ev[lo:hi].clone(), nevents)
        ok2 = bool(torch.equal(got, ev))
        q.put((rank, ok1, ok2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nfields,nevents", [(8, 10), (7, 5)])
def test_gloo_world2_gathers(nfields, nevents):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nfields, nevents, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True, True), (1, True, True)]


def test_field_assignment_of_the_c_abi():
    """mceik_fsm_assign_fields (host logic, no GPU): every rank holds one slowness model when there are at least as
    many ranks as models, fields are dealt evenly, a cost estimate balances the load (longest first), rows are
    rank-major, and the answer is deterministic."""
    import ctypes as C
    from mceik_b200 import _lib, sharding
    lib = _lib.load()
    nf, world = 128, 8
    fmodel = np.repeat(np.array([0, 1], np.int32), 64)
    rk, row, slots = sharding.assign_fields(fmodel, world)
    assert slots == 16 and np.all(np.bincount(rk, minlength=world) == 16)
    assert set(rk[:64]) == {0, 1, 2, 3} and set(rk[64:]) == {4, 5, 6, 7}
    assert sorted(row) == list(range(nf)) and np.all(row // slots == rk)
    rng = np.random.default_rng(0)
    cost = rng.integers(6, 11, nf).astype(np.int32)
    rk2, row2, _ = sharding.assign_fields(fmodel, world, cost)
    load = np.array([cost[rk2 == r].sum() for r in range(world)])
    naive = np.array([cost[r * 16:(r + 1) * 16].sum() for r in range(world)])
    assert load.max() - load.min() <= 2 and load.max() <= naive.max()
    assert np.all(np.bincount(rk2, minlength=world) == 16)
    rk3, row3, _ = sharding.assign_fields(fmodel, world, cost)
    assert np.array_equal(rk2, rk3) and np.array_equal(row2, row3)
    # more models than ranks, ragged counts
    fm = np.array([0, 1, 2, 0, 1, 2, 2], np.int32)
    rk4, row4, sl4 = sharding.assign_fields(fm, 2)
    assert sl4 == 4 and np.bincount(rk4, minlength=2).max() <= 4 and len(set(row4)) == 7
    # a single rank
    rk5, row5, sl5 = sharding.assign_fields(fm, 1)
    assert sl5 == 7 and not rk5.any() and list(row5) == list(range(7))
    # fewer fields than ranks: some ranks get nothing, every rank still owns `slots` rows
    rk6, row6, sl6 = sharding.assign_fields(np.array([3, 3, 5], np.int32), 8)
    assert sl6 == 1 and len(set(rk6)) == 3 and np.array_equal(row6, rk6)
    rk7, row7, sl7 = sharding.assign_fields(np.zeros(0, np.int32), 4)
    assert sl7 == 0 and len(rk7) == 0 and len(row7) == 0
