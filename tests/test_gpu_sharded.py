"""Two-rank GPU test of the multi-GPU C ABI (mceik_comm_init / mceik_fsm_solve_sharded_dev): every rank ends up with
every field's fp32 table, bit-equal to the oracle, and with all iteration counts -- through the in-place NCCL all-gather
(caller's buffer) and through one-sided puts into the library's replicated buffer (mceik_tables_alloc_replicated).  Needs two GPUs (skipped
otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`)."""
import numpy as np
import pytest

import cases
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _case():
    nx, ny, nz, h = 40, 32, 24, 500.0
    n = nx * ny * nz
    slow = np.stack([cases.checkerboard_slowness(nx, ny, nz, cell=8), cases.random_slowness(n, 9)])
    nf = 7
    fmodel = np.array([0, 1, 0, 1, 0, 0, 1], np.int32)
    xs, ys, zs = cases.interior_sources(nf, nx, ny, nz, h, seed=21)
    return nx, ny, nz, h, n, slow, nf, fmodel, xs, ys, zs


def _worker(rank, world, uid_q, res_q):
    import torch
    import mceik_b200
    from mceik_b200.eikonal import EikonalSolver
    nx, ny, nz, h, n, slow, nf, fmodel, xs, ys, zs = _case()
    torch.cuda.set_device(rank)
    ctx = mceik_b200.Context(rank)
    if rank == 0:
        uid = mceik_b200.Context.comm_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    ctx.comm_init(world, rank, uid)
    d_slow = torch.from_numpy(slow).cuda()
    slots = (nf + world - 1) // world
    d_all = torch.zeros((world * slots, n), dtype=torch.float32, device="cuda")
    sol = EikonalSolver(ctx, nx, ny, nz, h)
    torch.cuda.synchronize()
    out = []
    for cost in (None, np.array([3, 9, 4, 8, 5, 6, 7], np.int32)):  # plain deal, then longest-first by a cost estimate
        iters, ferr, row = sol.solve_sharded(d_slow, fmodel, np.zeros(nf), xs, ys, zs, d_all, cost=cost)
        ctx.synchronize()
        out.append((iters.copy(), ferr.copy(), row.copy(), d_all.cpu().numpy().copy()))
    # the library's replicated buffer: tables are put into the peers' copies over NVLink as the fields converge
    d_rep = ctx.tables_alloc_replicated(world * slots, n)  # (not cleared: a peer may already be putting into it)
    iters, ferr, row = sol.solve_sharded(d_slow, fmodel, np.zeros(nf), xs, ys, zs, d_rep, cost=None)
    ctx.synchronize()
    out.append((iters.copy(), ferr.copy(), row.copy(), d_rep.cpu().numpy().copy()))
    del d_rep
    ctx.tables_free_replicated()
    # fewer fields than ranks: rank 1 solves nothing and still ends up with the table, by either replication path
    one = []
    d_one = torch.zeros((world, n), dtype=torch.float32, device="cuda")
    d_rep1 = ctx.tables_alloc_replicated(world, n)
    for buf in (d_one, d_rep1):
        iters, ferr, row = sol.solve_sharded(d_slow, fmodel[:1], np.zeros(1), xs[:1], ys[:1], zs[:1], buf, cost=None)
        ctx.synchronize()
        one.append((iters.copy(), ferr.copy(), row.copy(), buf.cpu().numpy().copy()))
    del d_rep1, buf
    ctx.tables_free_replicated()
    ctx.comm_destroy()
    res_q.put((rank, out, one))


def test_sharded_solve_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    nx, ny, nz, h, n, slow, nf, fmodel, xs, ys, zs = _case()
    ref = [O.eikonal_serial(nx, ny, nz, h, slow[fmodel[f]], 0.0, xs[f], ys[f], zs[f]) for f in range(nf)]
    ctx = mp.get_context("spawn")
    uid_q, res_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, uid_q, res_q)) for r in range(2)]
    for p in procs:
        p.start()
    import queue
    import time
    got, t_end = [], time.time() + 240
    while len(got) < len(procs):  # a worker that died would leave its peer waiting in a collective: fail at once
        try:
            got.append(res_q.get(timeout=2))
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() > t_end:
                for p in procs:
                    p.kill()
                pytest.fail(f"sharded workers: exit codes {[p.exitcode for p in procs]}" if dead else "sharded workers timed out")
    res = {r: out for r, out, _ in got}
    res_one = {r: one for r, _, one in got}
    for p in procs:
        p.join(timeout=120)
    for rank in (0, 1):
        for iters, ferr, row, tabs in res[rank]:
            assert not ferr.any() and sorted(row) == sorted(set(row))
            for f in range(nf):
                u, ierr, it = ref[f]
                assert ierr == 0 and iters[f] == it, (rank, f)
                assert np.array_equal(tabs[row[f]], u.astype(np.float32)), (rank, f)
        for iters, ferr, row, tabs in res_one[rank]:
            u, ierr, it = ref[0]
            assert not ferr.any() and iters[0] == it and row[0] == 0
            assert np.array_equal(tabs[0], u.astype(np.float32)), rank
    # both ranks computed the same assignment
    assert all(np.array_equal(a[2], b[2]) for a, b in zip(res[0], res[1]))
