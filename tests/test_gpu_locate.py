"""GPU parity tests of the L2 grid search: CUDA (through the C ABI) vs the reference's own
locate.c (oracle/_ref) and the oracle restatements.  Located indices must be identical; t0 and
objective are compared bit for bit (tolerance 0) because both sides do the same IEEE operations
in the same order."""
import numpy as np
import pytest

import cases
import oracle_lib as O
import refcases

pytestmark = pytest.mark.gpu


def test_locate_c_main_known_answer_double(gpu_ctx):
    """locate.c main: index 107312, t0 = 4.0 (locate.c:118-182), and grids equal the reference's own code."""
    from mceik_b200 import locate as L
    c = refcases.locate_c_main_case()
    t0 = L.aligned_empty(c["ngrd"], np.float64)
    obj = L.aligned_empty(c["ngrd"], np.float64)
    rc = L.locate_l2_gridSearch__double64(c["ldgrd"], c["ngrd"], c["nobs"], 1, c["t0use"], c["mask"], c["tobs"], c["tcorr"],
                                          c["varobs"], c["test"], t0, obj)
    assert rc == 0
    iopt = L.locate_minLocDouble64(c["ngrd"], obj)
    assert iopt == c["true_index"] == 107312
    assert abs(t0[iopt] - 4.0) < 1e-9
    use_ref = O.ref() is not None
    rc2, t0r, objr = O.l2_gridsearch(c["ldgrd"], c["ngrd"], c["nobs"], 1, c["t0use"], c["mask"], c["tobs"], c["tcorr"],
                                     c["varobs"], c["test"], np.float64, use_ref=use_ref)
    assert rc2 == 0 and np.array_equal(t0, t0r) and np.array_equal(obj, objr)
    assert iopt == O.minloc(objr, use_ref=use_ref)


def test_locate_c_main_known_answer_float(gpu_ctx):
    from mceik_b200 import locate as L
    c = refcases.locate_c_main_case()
    n, ld, nobs = c["ngrd"], c["ldgrd"], c["nobs"]
    test4 = L.aligned_empty(nobs * ld, np.float32)
    test4[:] = c["test"].astype(np.float32)
    tobs4, var4, tc4 = c["tobs"].astype(np.float32), c["varobs"].astype(np.float32), np.zeros(nobs, np.float32)
    t0 = L.aligned_empty(n, np.float32)
    obj = L.aligned_empty(n, np.float32)
    assert L.locate_l2_gridSearch__float64(ld, n, nobs, 1, 4.0, c["mask"], tobs4, tc4, var4, test4, t0, obj) == 0
    assert L.locate_minLocFloat64(n, obj) == 107312
    use_ref = O.ref() is not None
    rc2, t0r, objr = O.l2_gridsearch(ld, n, nobs, 1, 4.0, c["mask"], tobs4, tc4, var4, test4, np.float32, use_ref=use_ref)
    assert rc2 == 0 and np.array_equal(t0, t0r) and np.array_equal(obj, objr)


def test_l2_masks_tcorr_fixed_t0_and_errors(gpu_ctx):
    """Masked picks, static corrections, iwantOT != 1, NULL tcorr and the argument errors of locate.c:948-974."""
    from mceik_b200 import locate as L
    rng = np.random.default_rng(3)
    nx, ny, nz, nobs = 23, 19, 11, 9
    n = nx * ny * nz
    ld = n + 64 - n % 64
    test = L.aligned_empty(nobs * ld, np.float64)
    for i in range(nobs):
        test[i * ld: i * ld + n] = refcases._table(nx, ny, nz, 500.0, 500.0, 500.0, *(rng.random(3) * 5000.0))
    tobs = rng.random(nobs) * 3 + 1
    tcorr = rng.normal(0, 0.05, nobs)
    var = rng.uniform(0.1, 0.9, nobs)
    mask = (rng.random(nobs) < 0.3).astype(np.int32)
    mask[0] = 0
    for want, tc in ((1, tcorr), (0, tcorr), (1, None)):
        t0 = L.aligned_empty(n, np.float64)
        obj = L.aligned_empty(n, np.float64)
        assert L.locate_l2_gridSearch__double64(ld, n, nobs, want, 2.5, mask, tobs, tc, var, test, t0, obj) == 0
        rc, t0r, objr = O.l2_gridsearch(ld, n, nobs, want, 2.5, mask, tobs, tc, var, test, np.float64, use_ref=O.ref() is not None)
        assert rc == 0 and np.array_equal(t0, t0r) and np.array_equal(obj, objr)
        assert L.locate_minLocDouble64(n, obj) == O.minloc(objr)
    t0 = L.aligned_empty(n, np.float64)
    obj = L.aligned_empty(n, np.float64)
    assert L.locate_l2_gridSearch__double64(ld + 1, n, nobs, 1, 0.0, mask, tobs, tcorr, var, test, t0, obj) == 1
    assert L.locate_l2_gridSearch__double64(64, n, nobs, 1, 0.0, mask, tobs, tcorr, var, test, t0, obj) == 1
    assert L.locate_l2_gridSearch__double64(ld, n, 0, 1, 0.0, mask, tobs, tcorr, var, test, t0, obj) == 1
    assert L.locate_l2_gridSearch__double64(ld, n, nobs, 1, 0.0, None, tobs, tcorr, var, test, t0, obj) == 1
    assert L.locate_l2_gridSearch__double64(ld, n, nobs, 1, 0.0, mask, tobs, tcorr, var, test[1:], t0, obj) == 1  # unaligned


def test_minloc_ties_and_nan(gpu_ctx):
    from mceik_b200 import locate as L
    x = np.array([3.0, 1.0, 2.0, 1.0, 1.0, 5.0] * 1000)
    assert L.locate_minLocDouble64(x.size, x) == O.minloc(x) == 1
    x[4001] = -7.0
    x[5000] = -7.0
    assert L.locate_minLocDouble64(x.size, x) == O.minloc(x) == 4001
    y = x.astype(np.float32)
    assert L.locate_minLocFloat64(y.size, y) == O.minloc(y) == 4001
    x[0] = np.nan  # sticky in the reference scan (locate.c:821-828)
    assert L.locate_minLocDouble64(x.size, x) == O.minloc(x) == 0
    z = np.full(100, np.inf)
    assert L.locate_minLocDouble64(z.size, z) == O.minloc(z) == 0


def test_gridsearch_f90_known_answer(gpu_ctx):
    """gridsearch.f90 program: 1-based index 21124, t0 = 4 by construction; fp64 and fp32 grids equal the oracle."""
    from mceik_b200 import locate as L
    c = refcases.gridsearch_f90_case()
    n, ld, nobs = c["ngrd"], c["ldgrd"], c["nobs"]
    pdf = L.aligned_empty(n, np.float64)
    assert L.locate3d_gridsearch__double64(ld, n, nobs, 1, c["mask"], c["tobs"], c["varobs"], c["test"], pdf) == 0
    ierr, ref, t0r = O.gridsearch_f90(ld, n, nobs, 1, c["mask"], c["tobs"], c["varobs"], c["test"], np.float64)
    assert ierr == 0 and np.array_equal(pdf, ref)
    iopt = L.locate_minLocDouble64(n, pdf)
    assert iopt + 1 == c["true_index_1based"] == 21124
    assert abs(t0r[iopt] - 4.0) < 1e-9
    test4 = L.aligned_empty(nobs * ld, np.float32)
    test4[:] = c["test"].astype(np.float32)
    tobs4, var4 = c["tobs"].astype(np.float32), c["varobs"].astype(np.float32)
    pdf4 = L.aligned_empty(n, np.float32)
    assert L.locate3d_gridsearch__float64(ld, n, nobs, 1, c["mask"], tobs4, var4, test4, pdf4) == 0
    ierr, ref4, _ = O.gridsearch_f90(ld, n, nobs, 1, c["mask"], tobs4, var4, test4, np.float32)
    assert ierr == 0 and np.array_equal(pdf4, ref4)
    # non-unit variances + a masked pick + the error returns of gridsearch.f90:404-424
    var = np.linspace(0.2, 1.5, nobs)
    mask = c["mask"].copy()
    mask[3] = 1
    assert L.locate3d_gridsearch__double64(ld, n, nobs, 1, mask, c["tobs"], var, c["test"], pdf) == 0
    ierr, ref, _ = O.gridsearch_f90(ld, n, nobs, 1, mask, c["tobs"], var, c["test"], np.float64)
    assert np.array_equal(pdf, ref)
    assert L.locate3d_gridsearch__double64(ld, n, nobs, 0, mask, c["tobs"], var, c["test"], pdf) == 0
    ierr, ref, _ = O.gridsearch_f90(ld, n, nobs, 0, mask, c["tobs"], var, c["test"], np.float64)
    assert np.array_equal(pdf, ref)
    assert L.locate3d_gridsearch__double64(ld - 1, n, nobs, 1, mask, c["tobs"], var, c["test"], pdf) == 1
    assert L.locate3d_gridsearch__double64(ld, n, nobs, 1, np.ones(nobs, np.int32), c["tobs"], var, c["test"], pdf) == 1
    assert L.locate3d_gridsearch__double64(ld, n, nobs, 1, mask, c["tobs"], np.zeros(nobs), c["test"], pdf) == 1


def _c1_case(nevents=100, n=50, nstat=20):
    """BASELINE config 1: homogeneous analytic tables, 20 stations on the top face, 100 events."""
    h, vp = 1000.0, 2000.0
    rng = np.random.default_rng(2016)
    sx = rng.integers(0, n, nstat) * h
    sy = rng.integers(0, n, nstat) * h
    sz = np.full(nstat, (n - 1) * h)
    tables = cases.homog_tables(n, n, n, h, sx, sy, sz, vp)
    cat = cases.synthetic_catalog(tables, nevents, seed=2017, mask_frac=0.0, variances=(0.25,))
    return n, h, tables, cat


@pytest.mark.parametrize("job", [2, 1])
def test_c1_catalog_matches_oracle(gpu_ctx, job):
    """locate3d_gridsearch drop-in on config 1: located nodes identical, hypo/t0 bit-equal to the oracle."""
    from mceik_b200 import locate as L
    n, h, tables, cat = _c1_case()
    ngrd = n ** 3
    X, Y, Z = cases.node_coords(n, n, n, h, h, h)
    assert L.locate3d_initialize() == 0
    L.locate3d_set_tables(tables, ngrd)
    L.locate3d_set_grid(X, Y, Z)
    hypo = np.zeros(4 * cat["nevents"])
    ierr = L.locate3d_gridsearch(1, job, cat["nobs"], cat["nevents"], cat["luseObs"], cat["statPtr"], cat["pickType"],
                                 cat["statCor"], cat["tori"], cat["varobs"], cat["tobs"], None, hypo)
    assert ierr == 0
    rc, hypo_ref, iopt_ref, obj_ref = O.locate3d_catalog(job, ngrd, ngrd, tables, cat["nobs"], cat["nevents"], cat["luseObs"],
                                                         cat["statPtr"], cat["pickType"], cat["statCor"], cat["tori"],
                                                         cat["varobs"], cat["tobs"], X, Y, Z)
    assert rc == 0
    assert np.array_equal(hypo, hypo_ref)
    assert np.array_equal(iopt_ref, cat["true_node"])  # noise-free picks sit on a node
    assert L.locate3d_gridsearch(1, 3, cat["nobs"], cat["nevents"], cat["luseObs"], cat["statPtr"], cat["pickType"],
                                 cat["statCor"], cat["tori"], cat["varobs"], cat["tobs"], None, hypo) == 1
    L.locate3d_finalize()


def test_batched_ragged_catalog(gpu_ctx):
    """Locator on a ragged CSR catalogue: masked picks, an event without picks, per-event table
    order that differs between events, noisy picks (optimum off the true node)."""
    from mceik_b200.locate import Locator
    n, h, tables, _ = _c1_case(nevents=1, n=24, nstat=7)
    ngrd = n ** 3
    ntab = tables.shape[0]
    rng = np.random.default_rng(99)
    ne = 37
    obs_ptr, tid, tc, var, tori = [0], [], [], [], rng.uniform(0, 5, ne)
    for e in range(ne):
        k = 0 if e == 5 else int(rng.integers(1, ntab + 1))
        ids = rng.permutation(ntab)[:k]
        node = int(rng.integers(0, ngrd))
        for t in ids:
            used = rng.random() > 0.15
            tid.append(int(t) if used else -1)
            tc.append(float(tables[t, node]) + tori[e] + rng.normal(0, 0.02))
            var.append(float(rng.choice([0.1, 0.25, 0.5])))
        obs_ptr.append(len(tid))
    obs_ptr, tid, tc, var = np.array(obs_ptr, np.int32), np.array(tid, np.int32), np.array(tc), np.array(var)
    loc = Locator(gpu_ctx)
    loc.set_tables_host(tables, ngrd)
    for job in (2, 1):
        iopt, t0, obj = loc.locate_host(job, obs_ptr, tid, tc, var, tori)
        for e in range(ne):
            b, en = obs_ptr[e], obs_ptr[e + 1]
            k = en - b
            use = (tid[b:en] >= 0).astype(np.int32)
            if use.sum() == 0:
                assert iopt[e] == -1
                continue
            stat = np.where(use == 1, tid[b:en] // 2 + 1, 1).astype(np.int32)
            ph = np.where(use == 1, tid[b:en] % 2 + 1, 1).astype(np.int32)
            rc, hy, io, ob = O.locate3d_catalog(job, ngrd, ngrd, tables, k, 1, use, stat, ph, np.zeros(k), tori[e:e + 1],
                                                var[b:en], tc[b:en], np.zeros(ngrd), np.zeros(ngrd), np.zeros(ngrd))
            assert rc == 0 and io[0] == iopt[e], f"event {e}"
            assert ob[0] == obj[e] and hy[3] == t0[e]


@pytest.mark.parametrize("ragged", [False, True])
def test_two_picks_of_one_table_in_an_event(gpu_ctx, ragged):
    """The same station and phase picked twice in one event (two analysts): both picks enter the sums, in catalogue
    order.  Uniform blocks (every event lists the same tables, duplicate included: fast kernel) and ragged, sorted
    ones (a repeated table is not strictly increasing, so the block stays on the general kernel).  Equal to the
    oracle bit for bit."""
    from mceik_b200.locate import Locator
    n, h, tables, _ = _c1_case(nevents=1, n=24, nstat=6)
    ngrd = n ** 3
    ntab = tables.shape[0]
    rng = np.random.default_rng(7 + ragged)
    ne = 19
    base = sorted(rng.permutation(ntab)[:7].tolist())
    base = base[:3] + [base[2]] + base[3:]          # table base[2] twice in a row
    obs_ptr, tid, tc, var, tori = [0], [], [], [], rng.uniform(0, 5, ne)
    for e in range(ne):
        ids = list(base)
        if ragged and e % 3 == 1:
            ids = ids[1:]                           # another subset, still sorted with the duplicate
        node = int(rng.integers(0, ngrd))
        for t in ids:
            tid.append(int(t))
            tc.append(float(tables[t, node]) + tori[e] + rng.normal(0, 0.03))
            var.append(float(rng.choice([0.1, 0.25, 0.5])))
        obs_ptr.append(len(tid))
    obs_ptr, tid, tc, var = np.array(obs_ptr, np.int32), np.array(tid, np.int32), np.array(tc), np.array(var)
    loc = Locator(gpu_ctx)
    loc.set_tables_host(tables, ngrd)
    for job in (2, 1):
        iopt, t0, obj = loc.locate_host(job, obs_ptr, tid, tc, var, tori)
        for e in range(ne):
            b, en = obs_ptr[e], obs_ptr[e + 1]
            k = en - b
            stat, ph = (tid[b:en] // 2 + 1).astype(np.int32), (tid[b:en] % 2 + 1).astype(np.int32)
            rc, hy, io, ob = O.locate3d_catalog(job, ngrd, ngrd, tables, k, 1, np.ones(k, np.int32), stat, ph, np.zeros(k),
                                                tori[e:e + 1], var[b:en], tc[b:en], np.zeros(ngrd), np.zeros(ngrd), np.zeros(ngrd))
            assert rc == 0 and io[0] == iopt[e], f"event {e}"
            assert ob[0] == obj[e] and hy[3] == t0[e]


def test_ragged_sorted_catalog_is_aligned_onto_the_fast_kernel(gpu_ctx):
    """Events with different subsets of the tables, picks in increasing table order (a station-ordered catalogue):
    the host entry re-lays each block of 8 events out over the union of its tables so the branch-free kernel runs it;
    results are the same bits as the general kernel (alignment switched off) and as the oracle, event by event."""
    import os
    from mceik_b200.locate import Locator
    n, h, tables, _ = _c1_case(nevents=1, n=24, nstat=9)
    ngrd = n ** 3
    ntab = tables.shape[0]
    rng = np.random.default_rng(123)
    ne = 45
    obs_ptr, tid, tc, var, tori = [0], [], [], [], rng.uniform(0, 5, ne)
    for e in range(ne):
        k = 0 if e == 11 else int(rng.integers(1, ntab + 1))
        ids = np.sort(rng.permutation(ntab)[:k])
        if e == 20:
            ids = ids[::-1]                      # one block that is not in table order -> stays on the general kernel
        node = int(rng.integers(0, ngrd))
        for t in ids:
            used = rng.random() > 0.15
            tid.append(int(t) if used else -1)
            tc.append(float(tables[t, node]) + tori[e] + rng.normal(0, 0.02))
            var.append(float(rng.choice([0.1, 0.25, 0.5])))
        obs_ptr.append(len(tid))
    obs_ptr, tid, tc, var = np.array(obs_ptr, np.int32), np.array(tid, np.int32), np.array(tc), np.array(var)
    loc = Locator(gpu_ctx)
    loc.set_tables_host(tables, ngrd)
    for job in (2, 1):
        iopt, t0, obj = loc.locate_host(job, obs_ptr, tid, tc, var, tori)
        gpu_ctx.set_tuning("LOCATE_NO_ALIGN", 1)
        try:
            iopt_g, t0_g, obj_g = loc.locate_host(job, obs_ptr, tid, tc, var, tori)
        finally:
            gpu_ctx.set_tuning("LOCATE_NO_ALIGN", 0)
        assert np.array_equal(iopt, iopt_g) and np.array_equal(t0, t0_g) and np.array_equal(obj, obj_g)
        for e in range(ne):
            b, en = obs_ptr[e], obs_ptr[e + 1]
            k = en - b
            use = (tid[b:en] >= 0).astype(np.int32)
            if use.sum() == 0:
                assert iopt[e] == -1
                continue
            stat = np.where(use == 1, tid[b:en] // 2 + 1, 1).astype(np.int32)
            ph = np.where(use == 1, tid[b:en] % 2 + 1, 1).astype(np.int32)
            rc, hy, io, ob = O.locate3d_catalog(job, ngrd, ngrd, tables, k, 1, use, stat, ph, np.zeros(k), tori[e:e + 1],
                                                var[b:en], tc[b:en], np.zeros(ngrd), np.zeros(ngrd), np.zeros(ngrd))
            assert rc == 0 and io[0] == iopt[e] and ob[0] == obj[e] and hy[3] == t0[e], f"event {e}"


def _check_events_against_oracle(job, ngrd, tables, obs_ptr, tid, tc, var, tori, iopt, t0, obj):
    for e in range(len(obs_ptr) - 1):
        b, en = obs_ptr[e], obs_ptr[e + 1]
        k = en - b
        use = (tid[b:en] >= 0).astype(np.int32)
        if use.sum() == 0:
            assert iopt[e] == -1, f"event {e}"
            continue
        stat = np.where(use == 1, tid[b:en] // 2 + 1, 1).astype(np.int32)
        ph = np.where(use == 1, tid[b:en] % 2 + 1, 1).astype(np.int32)
        rc, hy, io, ob = O.locate3d_catalog(job, ngrd, ngrd, tables, k, 1, use, stat, ph, np.zeros(k), tori[e:e + 1],
                                            var[b:en], tc[b:en], np.zeros(ngrd), np.zeros(ngrd), np.zeros(ngrd))
        assert rc == 0 and io[0] == iopt[e] and ob[0] == obj[e] and hy[3] == t0[e], f"event {e}"


@pytest.mark.parametrize("all_empty", [False, True])
def test_event_blocks_without_picks(gpu_ctx, all_empty):
    """A trailing block of 8 whose only event has a zero-length pick list (9 events), and a catalogue in which every
    event is empty: flagged iopt = -1, the other events equal the oracle, no fault in the fast kernel's preload."""
    from mceik_b200.locate import Locator
    n, h, tables, _ = _c1_case(nevents=1, n=20, nstat=4)
    ngrd = n ** 3
    ntab = tables.shape[0]
    rng = np.random.default_rng(77)
    ne = 9
    obs_ptr, tid, tc, var, tori = [0], [], [], [], rng.uniform(0, 5, ne)
    for e in range(ne):
        if not all_empty and e < ne - 1:
            node = int(rng.integers(0, ngrd))
            for t in range(ntab):
                tid.append(t)
                tc.append(float(tables[t, node]) + tori[e])
                var.append(0.25)
        obs_ptr.append(len(tid))
    obs_ptr = np.array(obs_ptr, np.int32)
    tid, tc, var = np.array(tid, np.int32), np.array(tc, np.float64), np.array(var, np.float64)
    loc = Locator(gpu_ctx)
    loc.set_tables_host(tables, ngrd)
    iopt, t0, obj = loc.locate_host(2, obs_ptr, tid, tc, var, tori)
    assert iopt[-1] == -1
    if all_empty:
        assert np.all(iopt == -1)
    _check_events_against_oracle(2, ngrd, tables, obs_ptr, tid, tc, var, tori, iopt, t0, obj)
    gpu_ctx.synchronize()


def test_more_picks_than_the_fast_kernel_holds(gpu_ctx):
    """420 picks per event (the same tables in every event's slot j, so the blocks are uniform, but wider than the
    392 slots of the fast kernel's shared memory): routed to the general kernel, equal to the oracle."""
    from mceik_b200.locate import Locator
    n, h, tables, _ = _c1_case(nevents=1, n=16, nstat=3)
    ngrd = n ** 3
    ntab = tables.shape[0]
    rng = np.random.default_rng(78)
    ne, npk = 10, 420
    obs_ptr, tid, tc, var, tori = [0], [], [], [], rng.uniform(0, 5, ne)
    for e in range(ne):
        node = int(rng.integers(0, ngrd))
        for j in range(npk):
            t = j % ntab
            tid.append(t if rng.random() > 0.1 else -1)
            tc.append(float(tables[t, node]) + tori[e] + rng.normal(0, 0.05))
            var.append(float(rng.choice([0.1, 0.25, 0.5])))
        obs_ptr.append(len(tid))
    obs_ptr = np.array(obs_ptr, np.int32)
    tid, tc, var = np.array(tid, np.int32), np.array(tc, np.float64), np.array(var, np.float64)
    loc = Locator(gpu_ctx)
    loc.set_tables_host(tables, ngrd)
    for job in (2, 1):
        iopt, t0, obj = loc.locate_host(job, obs_ptr, tid, tc, var, tori)
        _check_events_against_oracle(job, ngrd, tables, obs_ptr, tid, tc, var, tori, iopt, t0, obj)


def test_catalog_struct_entry(gpu_ctx):
    """mceik_locate_catalog with the mceik_struct.h layouts (CSR obsPtr, 1-based statPtr, P/S corrections)."""
    from mceik_b200.locate import Locator
    n, h, tables, cat = _c1_case(nevents=12, n=20, nstat=5)
    ngrd = n ** 3
    X, Y, Z = cases.node_coords(n, n, n, h, h, h)
    nobs = cat["nobs"]
    pcorr = np.linspace(-0.1, 0.1, 5)
    scorr = np.linspace(0.2, -0.2, 5)
    tobs = cat["tobs"].reshape(12, nobs).copy()
    tobs[:, 0::2] += pcorr[None, :]
    tobs[:, 1::2] += scorr[None, :]
    catalog = dict(nevents=12, tori=cat["tori"], tobs=tobs.ravel(), varObs=cat["varobs"], luseObs=cat["luseObs"],
                   pickType=cat["pickType"], statPtr=cat["statPtr"], obsPtr=np.arange(13, dtype=np.int32) * nobs)
    stations = dict(nstat=5, pcorr=pcorr, scorr=scorr)
    loc = Locator(gpu_ctx)
    loc.set_tables_host(tables, ngrd)
    loc.set_grid(X, Y, Z)
    hypo, iopt, obj = loc.locate_catalog(catalog, stations, 2)
    assert np.array_equal(iopt, cat["true_node"])
    assert np.allclose(hypo[:, 3], cat["tori"], atol=1e-5)
    assert np.array_equal(hypo[:, 0], X[iopt].astype(np.float64))
    # ... and a ragged, noisy catalogue (CSR obsPtr: every event its own pick list, unused picks, both jobs) against
    # the oracle event by event: located node, objective and hypo = (x, y, z, t0) bit for bit
    rng = np.random.default_rng(41)
    ne, ntab = 19, tables.shape[0]
    obs_ptr, stat, ptype, use, tob, var = [0], [], [], [], [], []
    tori = rng.uniform(0, 5, ne)
    for e in range(ne):
        k = int(rng.integers(1, ntab + 1))
        node = int(rng.integers(0, ngrd))
        for t in rng.permutation(ntab)[:k]:
            stat.append(t // 2 + 1)
            ptype.append(t % 2 + 1)
            use.append(int(rng.random() > 0.2))
            corr = pcorr[t // 2] if t % 2 == 0 else scorr[t // 2]
            tob.append(float(tables[t, node]) + tori[e] + corr + rng.normal(0, 0.03))
            var.append(float(rng.choice([0.1, 0.25, 0.5])))
        obs_ptr.append(len(stat))
    obs_ptr = np.array(obs_ptr, np.int32)
    stat, ptype, use = np.array(stat, np.int32), np.array(ptype, np.int32), np.array(use, np.int32)
    tob, var = np.array(tob), np.array(var)
    catalog = dict(nevents=ne, tori=tori, tobs=tob, varObs=var, luseObs=use, pickType=ptype, statPtr=stat, obsPtr=obs_ptr)
    for job in (2, 1):
        hypo, iopt, obj = loc.locate_catalog(catalog, stations, job)
        for e in range(ne):
            b, en = obs_ptr[e], obs_ptr[e + 1]
            if use[b:en].sum() == 0:
                assert iopt[e] == -1
                continue
            corr = np.where(ptype[b:en] == 1, pcorr[stat[b:en] - 1], scorr[stat[b:en] - 1])
            rc, hy, io, ob = O.locate3d_catalog(job, ngrd, ngrd, tables, en - b, 1, use[b:en], stat[b:en], ptype[b:en], corr,
                                                tori[e:e + 1], var[b:en], tob[b:en], X, Y, Z)
            assert rc == 0 and io[0] == iopt[e] and ob[0] == obj[e], f"job {job} event {e}"
            assert np.array_equal(hy, hypo[e]), f"job {job} event {e}"


def test_large_catalog_noise_free_round_trip(gpu_ctx):
    """Size-independent property at a large size (128^3 grid, 64 device-generated tables, 512 events, 10 %
    masked picks): picks synthesised from the tables at a node are located on exactly that node with
    t0 = origin time; sharding the events into two halves gives the same answers (no data-path collective)."""
    import ctypes as C
    import torch
    from mceik_b200 import _lib
    from mceik_b200.locate import Locator
    from mceik_b200 import sharding
    n, h, ns = 128, 1000.0, 32
    N, ntab, ne = n ** 3, 64, 512
    rng = np.random.default_rng(5)
    X, Y, Z = np.repeat(rng.uniform(0, (n - 1) * h, ns), 2), np.repeat(rng.uniform(0, (n - 1) * h, ns), 2), np.full(ntab, (n - 1) * h)
    V = np.tile(np.array([5000.0, 5000.0 / np.sqrt(3.0)]), ns)
    d_tab = torch.empty((ntab, N), dtype=torch.float32, device="cuda")
    p = lambda x: x.ctypes.data_as(_lib.c_dbl_p)
    assert _lib.load().mceik_homogeneous_tables_dev(gpu_ctx.handle, n, n, n, 0.0, 0.0, 0.0, h, h, h, ntab, p(X), p(Y), p(Z), p(V),
                                                    C.c_void_p(d_tab.data_ptr()), N) == 0
    true_node = rng.integers(0, N, ne)
    tori = rng.uniform(0, 10, ne)
    tobs = (d_tab[:, torch.from_numpy(true_node).cuda()].T.double().cpu().numpy() + tori[:, None]).ravel()
    use = rng.uniform(size=ne * ntab) >= 0.1
    tid = np.where(use, np.tile(np.arange(ntab), ne), -1).astype(np.int32)
    var = rng.choice(np.array([0.1, 0.25, 0.5]), ne * ntab)
    obs_ptr = (np.arange(ne + 1) * ntab).astype(np.int32)
    loc = Locator(gpu_ctx)
    loc.set_tables_device(d_tab, N)
    iopt, t0, obj = loc.locate_host(2, obs_ptr, tid, tobs, var)
    assert np.array_equal(iopt, true_node)
    assert np.allclose(t0, tori, rtol=0, atol=1e-5) and np.all(obj < 1e-9)
    parts = []
    for r in range(2):
        lo, hi, lptr, p0, p1 = sharding.shard_events(obs_ptr, 2, r)
        parts.append(loc.locate_host(2, lptr, tid[p0:p1], tobs[p0:p1], var[p0:p1]))
    assert np.array_equal(np.concatenate([q[0] for q in parts]), iopt)
    assert np.array_equal(np.concatenate([q[1] for q in parts]), t0)
    assert np.array_equal(np.concatenate([q[2] for q in parts]), obj)


def test_event_posterior_volume(gpu_ctx):
    """Full logPDF / t0 volumes of one event equal the oracle bit for bit; the fp32 copy
    is the rounded fp64 volume; its argmax is the event located by the batched search."""
    from mceik_b200.locate import Locator
    n, h, tables, cat = _c1_case(nevents=3, n=30, nstat=6)
    ngrd = n ** 3
    loc = Locator(gpu_ctx)
    loc.set_tables_host(tables, ngrd)
    nobs = cat["nobs"]
    for e in range(3):
        sl = slice(e * nobs, (e + 1) * nobs)
        tid = np.where(cat["luseObs"][sl] == 1, np.arange(nobs), -1).astype(np.int32)
        tid[e] = -1
        tc = cat["tobs"][sl] + np.random.default_rng(e).normal(0, 0.01, nobs)
        for job in (2, 1):
            pdf, pdf4, t0 = loc.event_logpdf(job, tid, tc, cat["varobs"][sl], tori=cat["tori"][e], want_f32=True)
            rc, pdf_ref, t0_ref = O.event_logpdf(job, ngrd, ngrd, tables, tid, tc, cat["varobs"][sl], cat["tori"][e])
            assert rc == 0 and np.array_equal(pdf, pdf_ref) and np.array_equal(t0, t0_ref)
            assert np.array_equal(pdf4, pdf_ref.astype(np.float32))
            iopt, t0o, objo = loc.locate_host(job, np.array([0, nobs], np.int32), tid, tc, cat["varobs"][sl], cat["tori"][e:e + 1])
            assert iopt[0] == int(np.argmax(pdf)) and objo[0] == -pdf[iopt[0]] and t0o[0] == t0[iopt[0]]


def test_pdf_helpers(gpu_ctx):
    """LOCATE_OPTNODE = first index of the maximum (bit-exact); LOCATE_NORMALIZE_PDF = pdf * (1/sum) with the
    sum within 1e-12 relative of the sequential Fortran-order sum (locate.f90:43-117); zero sum -> ierr 1."""
    from mceik_b200.locate import Locator
    loc = Locator(gpu_ctx)
    rng = np.random.default_rng(11)
    for n in (1, 7, 1000, 300_001):
        pdf = np.exp(-rng.uniform(0.0, 30.0, n))
        if n > 10:
            pdf[[n // 3, n // 2]] = pdf.max() * 2.0   # tie: the first one wins
        assert loc.optnode(pdf) == int(np.argmax(pdf))
        seq = float(np.cumsum(pdf)[-1])               # sequential sum, the order of the reference's SUM at -O2
        work = pdf.copy()
        ierr, xsum = loc.normalize_pdf(work)
        assert ierr == 0 and abs(xsum - seq) <= 1e-12 * abs(seq)
        assert np.array_equal(work, pdf * (1.0 / xsum))   # DSCAL by the reciprocal (locate.f90:61-62)
        assert abs(work.sum() - 1.0) < 1e-12
    zero = np.zeros(64)
    ierr, xsum = loc.normalize_pdf(zero)
    assert ierr == 1 and xsum == 0.0 and not zero.any()


def test_l1_gridsearch_matches_oracle(gpu_ctx):
    """L1 flavour (locate.c:1205-1335; parity unpinned by the reference, see include/mceik_b200.h): t0 and misfit
    grids equal the oracle bit for bit; on the noise-free locate.c main case the weighted median is the origin
    time and the misfit vanishes at the true node."""
    from mceik_b200 import locate as L
    c = refcases.locate_c_main_case()
    n, ld, nobs = c["ngrd"], c["ldgrd"], c["nobs"]
    t0, obj = np.zeros(n), np.zeros(n)
    assert L.locate_l1_gridSearch__double64(ld, n, nobs, 1, 0.0, c["mask"], c["tobs"], c["varobs"], c["test"], t0, obj) == 0
    rc, t0r, objr = O.l1_gridsearch(ld, n, nobs, 1, 0.0, c["mask"], c["tobs"], c["varobs"], c["test"])
    assert rc == 0 and np.array_equal(t0, t0r) and np.array_equal(obj, objr)
    iopt = L.locate_minLocDouble64(n, obj)
    assert iopt == c["true_index"] and abs(t0[iopt] - 4.0) < 1e-9 and obj[iopt] < 1e-9
    # noisy picks, unequal variances, masked picks, ties between residuals; then the fixed-t0 branch
    rng = np.random.default_rng(3)
    nobs, n = 23, 4097
    ld = 4104
    test = rng.uniform(0.5, 9.0, (nobs, ld)).round(2)            # two decimals: many equal residuals
    tobs = (test[:, 1234] + 3.0 + rng.normal(0, 0.05, nobs)).round(2)
    var = rng.choice([0.1, 0.25, 0.5, 1.0], nobs)
    mask = (rng.uniform(size=nobs) < 0.2).astype(np.int32)
    for want, t0use in ((1, 0.0), (0, 2.5)):
        t0, obj = np.zeros(n), np.zeros(n)
        assert L.locate_l1_gridSearch__double64(ld, n, nobs, want, t0use, mask, tobs, var, test.ravel(), t0, obj) == 0
        rc, t0r, objr = O.l1_gridsearch(ld, n, nobs, want, t0use, mask, tobs, var, test.ravel())
        assert rc == 0 and np.array_equal(t0, t0r) and np.array_equal(obj, objr)
    assert L.locate_l1_gridSearch__double64(n - 1, n, nobs, 1, 0.0, mask, tobs, var, test.ravel(), t0, obj) == 1
