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
