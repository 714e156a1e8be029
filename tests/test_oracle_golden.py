"""CPU tests: the oracle against the reference's own known answers, against outputs of the
reference's own locate.c (live oracle/_ref when present, committed fixtures otherwise), and the
by-construction facts that pin the FSM restatement (the reference has no FSM golden vector)."""
import os

import numpy as np
import pytest

import cases
import oracle_lib as O
import refcases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_locate_c_main_known_answer():
    """locate.c:118-182 -> 'double estimate: 107312 4.000000', 'float estimate: 107312'."""
    c = refcases.locate_c_main_case()
    rc, t0, obj = O.l2_gridsearch(c["ldgrd"], c["ngrd"], c["nobs"], 1, 4.0, c["mask"], c["tobs"], c["tcorr"], c["varobs"], c["test"])
    assert rc == 0
    iopt = O.minloc(obj)
    assert iopt == c["true_index"] == 107312 and abs(t0[iopt] - 4.0) < 1e-9
    if O.ref() is not None:  # the reference's own object code gives the same bits
        rc, t0r, objr = O.l2_gridsearch(c["ldgrd"], c["ngrd"], c["nobs"], 1, 4.0, c["mask"], c["tobs"], c["tcorr"], c["varobs"],
                                        c["test"], use_ref=True)
        assert np.array_equal(t0, t0r) and np.array_equal(obj, objr) and O.minloc(objr, use_ref=True) == 107312
    t4 = O.aligned(c["nobs"] * c["ldgrd"], np.float32)
    t4[:] = c["test"].astype(np.float32)
    rc, t0f, objf = O.l2_gridsearch(c["ldgrd"], c["ngrd"], c["nobs"], 1, 4.0, c["mask"], c["tobs"].astype(np.float32),
                                    np.zeros(c["nobs"], np.float32), c["varobs"].astype(np.float32), t4, np.float32)
    assert rc == 0 and O.minloc(objf) == 107312


def test_oracle_equals_reference_fixture():
    """Fixture = outputs of the unmodified reference locate.c (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "locate_ref_small.npz"))
    n, ld, nobs = int(g["nx"]) * int(g["ny"]) * int(g["nz"]), int(g["ld"]), int(g["nobs"])
    test = O.aligned(nobs * ld, np.float64)
    test[:] = g["test"]
    rc, t0, obj = O.l2_gridsearch(ld, n, nobs, 1, 0.0, g["mask"], g["tobs"], g["tcorr"], g["var"], test)
    assert rc == 0 and np.array_equal(t0, g["t0_f64"]) and np.array_equal(obj, g["obj_f64"])
    assert O.minloc(obj) == int(g["iopt_f64"])
    rc, t0, obj = O.l2_gridsearch(ld, n, nobs, 0, 1.75, g["mask"], g["tobs"], None, g["var"], test)
    assert rc == 0 and np.array_equal(t0, g["t0_fix"]) and np.array_equal(obj, g["obj_fix"])
    t4 = O.aligned(nobs * ld, np.float32)
    t4[:] = g["test"].astype(np.float32)
    rc, t0, obj = O.l2_gridsearch(ld, n, nobs, 1, 0.0, g["mask"], g["tobs"].astype(np.float32), g["tcorr"].astype(np.float32),
                                  g["var"].astype(np.float32), t4, np.float32)
    assert rc == 0 and np.array_equal(t0, g["t0_f32"]) and np.array_equal(obj, g["obj_f32"])
    assert O.minloc(obj) == int(g["iopt_f32"])


def test_l2_argument_errors():
    """locate.c:948-974: ldgrd bytes % 64, ldgrd < ngrd, nobs < 1, NULLs, misaligned arrays -> 1."""
    n, ld, nobs = 100, 128, 3
    test = O.aligned(nobs * ld, np.float64)
    a = (np.zeros(nobs, np.int32), np.ones(nobs), np.zeros(nobs), np.ones(nobs))
    assert O.l2_gridsearch(ld, n, nobs, 1, 0.0, *a, test)[0] == 0
    assert O.l2_gridsearch(ld + 1, n, nobs, 1, 0.0, *a, test)[0] == 1
    assert O.l2_gridsearch(64, n, nobs, 1, 0.0, *a, test)[0] == 1
    assert O.l2_gridsearch(ld, n, 0, 1, 0.0, *a, test)[0] == 1
    assert O.l2_gridsearch(ld, n, nobs, 1, 0.0, None, a[1], a[2], a[3], test)[0] == 1
    assert O.l2_gridsearch(ld, n, nobs, 1, 0.0, *a, test[1:])[0] == 1
    if O.ref() is not None:
        assert O.l2_gridsearch(ld + 1, n, nobs, 1, 0.0, *a, test, use_ref=True)[0] == 1
        assert O.l2_gridsearch(ld, n, nobs, 1, 0.0, *a, test[1:], use_ref=True)[0] == 1


def test_minloc_first_strict_minimum():
    x = np.array([3.0, 1.0, 2.0, 1.0, 1.0])
    assert O.minloc(x) == 1
    x[0] = np.nan
    assert O.minloc(x) == 0
    if O.ref() is not None:
        assert O.minloc(x, use_ref=True) == 0
        assert O.minloc(np.array([2.0, 2.0, 0.5, 0.5]), use_ref=True) == 2


def test_gridsearch_f90_known_answer_and_flavour_equivalence():
    """gridsearch.f90:38-87 -> 1-based index 21124, t0 = 4; equals the C flavour when all var = 1."""
    c = refcases.gridsearch_f90_case()
    ierr, pdf, t0 = O.gridsearch_f90(c["ldgrd"], c["ngrd"], c["nobs"], 1, c["mask"], c["tobs"], c["varobs"], c["test"])
    assert ierr == 0
    iopt = O.minloc(pdf)
    assert iopt + 1 == c["true_index_1based"] == 21124 and abs(t0[iopt] - 4.0) < 1e-9
    rc, t0c, objc = O.l2_gridsearch(c["ldgrd"], c["ngrd"], c["nobs"], 1, 0.0, c["mask"], c["tobs"], None, c["varobs"], c["test"])
    assert O.minloc(objc) == iopt
    np.testing.assert_allclose(t0, t0c, rtol=0, atol=1e-12)
    np.testing.assert_allclose(pdf, objc, rtol=1e-12, atol=1e-18)
    # error returns of gridsearch.f90:404-424
    assert O.gridsearch_f90(c["ldgrd"] - 1, c["ngrd"], c["nobs"], 1, c["mask"], c["tobs"], c["varobs"], c["test"])[0] == 1
    assert O.gridsearch_f90(c["ldgrd"], c["ngrd"], c["nobs"], 1, np.ones(c["nobs"], np.int32), c["tobs"], c["varobs"], c["test"])[0] == 1
    assert O.gridsearch_f90(c["ldgrd"], c["ngrd"], c["nobs"], 1, c["mask"], c["tobs"], np.zeros(c["nobs"]), c["test"])[0] == 1


def test_catalog_flavour_consistent_with_c_flavour():
    """The catalogue search (fp32 tables promoted to fp64) equals the C flavour run on the promoted table."""
    rng = np.random.default_rng(8)
    n = 12
    ngrd = n ** 3
    tables = cases.homog_tables(n, n, n, 500.0, rng.uniform(0, 5000, 3), rng.uniform(0, 5000, 3), np.full(3, 5500.0), 3000.0)
    cat = cases.synthetic_catalog(tables, 5, seed=1)
    X, Y, Z = cases.node_coords(n, n, n, 500.0, 500.0, 500.0)
    rc, hypo, iopt, obj = O.locate3d_catalog(2, ngrd, ngrd, tables, 6, 5, cat["luseObs"], cat["statPtr"], cat["pickType"],
                                             cat["statCor"], cat["tori"], cat["varobs"], cat["tobs"], X, Y, Z)
    assert rc == 0
    ld = ngrd + 64 - ngrd % 64
    test = O.aligned(6 * ld, np.float64)
    for t in range(6):
        test[t * ld:t * ld + ngrd] = tables[t].astype(np.float64)
    for e in range(5):
        sl = slice(6 * e, 6 * e + 6)
        mask = (cat["luseObs"][sl] == 0).astype(np.int32)
        rc, t0, ob = O.l2_gridsearch(ld, ngrd, 6, 1, 0.0, mask, cat["tobs"][sl], None, cat["varobs"][sl], test)
        k = O.minloc(ob)
        assert k == iopt[e] and ob[k] == obj[e] and t0[k] == hypo[4 * e + 3]
        assert hypo[4 * e] == float(X[k])
    assert O.locate3d_catalog(3, ngrd, ngrd, tables, 6, 5, cat["luseObs"], cat["statPtr"], cat["pickType"], cat["statCor"],
                              cat["tori"], cat["varobs"], cat["tobs"], X, Y, Z)[0] == 1


# ------------------------------------------------------------------------------------------ FSM
def test_fsm_xfsm3d_case_facts():
    """fsm3d.f90:2085-2100 (70x80x90, h=100, v=5000, centre source, maxit=5): the solve converges in 2
    iterations; u >= analytic (upwind scheme over-estimates), first-order error ~0.038 s, min = 0."""
    nx, ny, nz, h = 70, 80, 90, 100.0
    slow = np.full(nx * ny * nz, 1.0 / 5000.0)
    xs, ys, zs = h * nx / 2, h * ny / 2, h * nz / 2
    u, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow, 0.0, xs, ys, zs, tol=1e-7, maxit=5)
    assert ierr == 0 and it == 2
    x, y, z = np.arange(nx) * h, np.arange(ny) * h, np.arange(nz) * h
    d = np.sqrt((x[None, None, :] - xs) ** 2 + (y[None, :, None] - ys) ** 2 + (z[:, None, None] - zs) ** 2).ravel()
    assert u.min() == 0.0 and abs(u.max() - 1.4308203212738235) < 1e-12
    assert (u - d / 5000.0).min() > -1e-12 and np.abs(u - d / 5000.0).max() < 0.04
    # the 27 stencil nodes hold ts + d*slow exactly (fsm3d.f90:823-829)
    near = np.argsort(d)[:27]
    assert np.array_equal(u[near], 0.0 + d[near] * slow[near])


def test_fsm_golden_regression_and_properties():
    g = np.load(os.path.join(GOLD, "fsm_oracle_small.npz"))
    nx, ny, nz, h = int(g["nx"]), int(g["ny"]), int(g["nz"]), float(g["h"])
    args = (nx, ny, nz, h, g["slow"], [0.0, 0.1], [h * 7.3, h * 15.0], [h * 9.9, h * 3.0], [h * 6.2, h * 11.5])
    u, ierr, it = O.eikonal_serial(*args, tol=1e-7, maxit=20)
    assert ierr == 0 and it == int(g["iters"]) and np.array_equal(u, g["u"])
    prev = None
    for k in (1, 2, 3):  # travel times only decrease from one iteration to the next
        uk, _, itk = O.eikonal_serial(*args, tol=1e-7, maxit=k)
        assert itk == min(k, it)
        if prev is not None:
            assert np.all(uk <= prev)
        prev = uk
    assert np.all(u < 1e300) and np.all(u >= 0.0)


@pytest.mark.parametrize("case", ["random_two_sources", "checkerboard_source_on_a_node", "offset_origin_maxit",
                                  "source_near_the_far_corner", "source_on_node_1_fails"])
def test_fsm_oracle_equals_a_second_reading_of_the_reference(case):
    """No Fortran compiler exists here or on the GPU boxes, so the C oracle cannot be pinned against the compiled
    reference.  tests/fsm_restatement.py restates the same fsm3d.f90 lines a second time (plain Python floats: IEEE
    doubles, no FMA, correctly rounded sqrt); the two readings must agree bit for bit on fields, error codes and
    iteration counts."""
    import cases
    import fsm_restatement as R
    tol, maxit, x0, y0, z0 = 1e-6, 20, 0.0, 0.0, 0.0
    if case == "random_two_sources":
        nx, ny, nz, h = 13, 11, 9, 100.0
        slow = cases.random_slowness(nx * ny * nz, seed=17)
        ts, xs, ys, zs = [0.0, 0.3], [h * 4.37, h * 10.2], [h * 6.61, h * 2.0], [h * 3.45, h * 6.5]
    elif case == "checkerboard_source_on_a_node":  # the three-node stencil of EIKONAL_INIT_GRID's ELSE branch
        nx, ny, nz, h = 12, 10, 14, 250.0
        slow = cases.checkerboard_slowness(nx, ny, nz, cell=4)
        ts, xs, ys, zs = [0.0], [h * 5.0], [h * 4.0], [h * 7.0]
    elif case == "offset_origin_maxit":            # non-zero grid origin, stopped by maxit before convergence
        nx, ny, nz, h = 10, 12, 8, 50.0
        slow = cases.random_slowness(nx * ny * nz, seed=3)
        x0, y0, z0, maxit = -120.0, 35.5, 1000.0, 2
        ts, xs, ys, zs = [1.5], [x0 + h * 2.7], [y0 + h * 9.1], [z0 + h * 4.4]
    elif case == "source_near_the_far_corner":     # stencil against the upper grid faces, clamped neighbours
        nx, ny, nz, h = 9, 9, 9, 100.0
        slow = cases.random_slowness(nx * ny * nz, seed=5)
        ts, xs, ys, zs = [0.0], [h * 7.6], [h * 7.2], [h * 7.9]
    else:                                          # node 1 needs node 0: SETBCS fails, nothing is solved
        nx, ny, nz, h = 8, 8, 8, 100.0
        slow = cases.random_slowness(nx * ny * nz, seed=9)
        ts, xs, ys, zs = [0.0], [0.0], [h * 3.3], [h * 4.1]
    ref, ierr_ref, it_ref = R.serial_driver(nx, ny, nz, h, list(map(float, slow)), ts, xs, ys, zs, tol=tol, maxit=maxit,
                                            x0=x0, y0=y0, z0=z0)
    u, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow, ts, xs, ys, zs, tol=tol, maxit=maxit, x0=x0, y0=y0, z0=z0)
    if case == "source_on_node_1_fails":
        assert ierr_ref == 1 and ierr == 1
        return
    assert ierr == ierr_ref == 0 and it == it_ref
    assert np.array_equal(u, np.array(ref)), f"{np.count_nonzero(u != np.array(ref))} nodes differ"
    if case == "offset_origin_maxit":
        assert it == 2
    else:
        assert 2 <= it < 20


def test_fsm_source_stencil_quirks():
    """EIKONAL_INIT_GRID (fsm3d.f90:716-755): a source on node 1 or outside the grid fails; on node nx-1
    it keeps 2 nodes along that axis; EIKONAL_SOURCE_INDEX rounds to the nearest node."""
    import ctypes as C
    L = O.lib()
    loc = (C.c_int * 3)()
    f = lambda n, xs: L.oracle_source_index(C.c_int(n), C.c_double(0.0), C.c_double(10.0), C.c_double(xs))
    assert f(10, -3.0) == 1 and f(10, 0.0) == 1 and f(10, 4.9) == 1 and f(10, 5.0) == 2 and f(10, 90.0) == 10 and f(10, 200.0) == 10
    g = lambda n, isx, xs: (L.oracle_init_grid(C.c_int(n), C.c_int(isx), C.c_double(0.0), C.c_double(10.0), C.c_double(xs), loc), list(loc))
    assert g(10, 3, 23.0) == (0, [3, 4, -1])
    assert g(10, 3, 17.0) == (0, [2, 3, -1])
    assert g(10, 3, 20.0) == (0, [2, 3, 4])
    assert g(10, 9, 80.0) == (0, [8, 9, -1])       # isx = nx-1: upper neighbour dropped (:742)
    assert g(10, 1, 0.0)[0] == 1                   # node 0 requested (:736)
    assert g(10, 10, 95.0)[0] == 1                 # node nx+1 requested
    n = 8
    slow = np.full(n ** 3, 1e-3)
    assert O.eikonal_serial(n, n, n, 10.0, slow, 0.0, 0.0, 35.0, 35.0)[1] == 1
    assert O.eikonal_serial(n, n, n, 10.0, slow, 0.0, 35.0, 35.0, 35.0)[1] == 0


def test_hamiltonian_branches():
    """SOLVE_HAMILTONIAN3D (fsm3d.f90:648-693): p = 1, 2, 3 and the u_nan passthrough."""
    H = 1.7976931348623157e308
    assert O.hamiltonian3d(H, H, H, 1.0) == (H, 0)
    assert O.hamiltonian3d(1.0, H, H, 0.5) == (1.5, 0)
    assert O.hamiltonian3d(H, 5.0, 1.0, 0.5) == (1.5, 0)
    v, e = O.hamiltonian3d(1.0, 1.0, H, 1.0)
    assert e == 0 and abs(v - (1.0 + np.sqrt(0.5))) < 1e-15
    v, e = O.hamiltonian3d(1.0, 1.0, 1.0, 1.0)
    assert e == 0 and abs(v - (1.0 + 1.0 / np.sqrt(3.0))) < 1e-15
    v, e = O.hamiltonian3d(2.0, 1.0, 3.0, 0.25)
    assert (v, e) == (1.25, 0)


def test_homogeneous_tables_match_numpy():
    t = O.homogeneous_traveltimes(9, 8, 7, 0.0, 0.0, 0.0, 100.0, 100.0, 100.0, 250.0, 330.0, 600.0, 2000.0)
    ref = cases.homog_tables(9, 8, 7, 100.0, [250.0], [330.0], [600.0], 2000.0)[0]
    assert np.array_equal(t.astype(np.float32), ref)


def _wmed_numpy(x, w):
    """Independent statement of the weighted median of include/mceik_b200.h."""
    order = np.lexsort((np.arange(x.size), x))
    cum = np.cumsum(w[order])
    half = 0.5 * cum[-1]
    k = int(np.argmax(cum >= half))
    if cum[k] == half and k + 1 < x.size:
        return 0.5 * (x[order[k]] + x[order[k + 1]])
    return x[order[k]]


def test_weighted_median_definition():
    """The reference only declares its weighted median (locate.c:73); its demo inputs (locate.c:228-237) give
    0.2 with weights = values and the ordinary median 0.1 with unit weights.  Oracle == numpy statement ==
    the host helper the library exports under the reference's name."""
    from mceik_b200 import locate as L
    xs = np.array([0.1, 0.35, 0.05, 0.1, 0.15, 0.05, 0.2])
    assert O.weighted_median(xs, xs) == 0.2
    assert O.weighted_median(xs, np.ones(7)) == 0.1
    assert O.weighted_median(np.array([1.0, 4.0, 2.0, 3.0]), np.ones(4)) == 2.5      # even count: mean of the middle pair
    assert O.weighted_median(np.array([5.0]), np.array([2.0])) == 5.0
    rng = np.random.default_rng(0)
    for n in (2, 3, 8, 33, 100):
        for _ in range(20):
            x = rng.normal(size=n).round(1)
            w = rng.choice([0.25, 0.5, 1.0, 2.0], n)
            ref = _wmed_numpy(x, w)
            assert O.weighted_median(x, w) == ref
            perm = np.arange(n, dtype=np.int32)
            med, lsort, ierr = L.weightedMedian__double(x, w, perm)
            assert med == ref and ierr == 0 and lsort == bool(np.any(np.diff(x) < 0))
            med2, lsort2, _ = L.weightedMedian__double(x, w, perm)     # the ordering is kept: no second sort
            assert med2 == ref and not lsort2


def test_l1_oracle_noise_free_case():
    """L1 oracle on the locate.c main inputs: weighted-median t0 = 4 and zero misfit at the true node."""
    c = refcases.locate_c_main_case()
    rc, t0, obj = O.l1_gridsearch(c["ldgrd"], c["ngrd"], c["nobs"], 1, 0.0, c["mask"], c["tobs"], c["varobs"], c["test"])
    assert rc == 0 and int(np.argmin(obj)) == c["true_index"] and abs(t0[c["true_index"]] - 4.0) < 1e-9
