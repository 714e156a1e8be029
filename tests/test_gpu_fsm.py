"""GPU parity tests of the eikonal path: CUDA (through the C ABI) vs the CPU oracle, bit for bit.

Bit-exactness holds because updates inside one hyperplane are order independent
and neither side fuses multiply-adds.  Tolerance in these tests: 0 ulp (np.array_equal).
"""
import numpy as np
import pytest

import cases
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _oracle_fields(nx, ny, nz, h, slow, fmodel, xs, ys, zs, ts, tol, maxit):
    out, its = [], []
    for f in range(len(fmodel)):
        u, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow[fmodel[f]], ts[f], xs[f], ys[f], zs[f], tol=tol, maxit=maxit)
        assert ierr == 0
        out.append(u)
        its.append(it)
    return np.stack(out), np.array(its)


@pytest.mark.parametrize("algo", [0, 1, 2])
def test_xfsm3d_case_matches_oracle(gpu_ctx, algo):
    """The reference's own driver case (fsm3d.f90:2085-2100): 70x80x90, h=100, v=5000, centre source."""
    from mceik_b200.eikonal import EikonalSolver
    nx, ny, nz, h = 70, 80, 90, 100.0
    slow = np.full(nx * ny * nz, 1.0 / 5000.0)
    xs, ys, zs = h * nx / 2, h * ny / 2, h * nz / 2
    ref, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow, 0.0, xs, ys, zs, tol=1e-7, maxit=5)
    sol = EikonalSolver(gpu_ctx, nx, ny, nz, h, tol=1e-7, maxit=5, algo=algo)
    u, _, iters, ferr = sol.solve_host(slow[None], [0], [0.0], [xs], [ys], [zs])
    assert ferr[0] == 0 and iters[0] == it == 2
    assert np.array_equal(u[0], ref)
    assert sol.node_updates == nx * ny * nz * 8 * it


@pytest.mark.parametrize("shape", [(37, 50, 21), (16, 16, 16), (5, 70, 3), (33, 17, 48),
                                   # more than one brick layer in z (bricks span 256 planes), incl. a 1-plane layer
                                   (16, 24, 300), (8, 8, 257), (12, 9, 260)])
def test_random_model_two_sources_bit_exact(gpu_ctx, shape):
    """Heterogeneous random slowness, partial tiles, two sources seeding one field."""
    from mceik_b200.eikonal import EikonalSolver
    nx, ny, nz = shape
    h = 100.0
    slow = cases.random_slowness(nx * ny * nz, seed=7)
    ts = np.array([0.0, 0.3])
    xs = np.array([h * nx * 0.37, h * nx * 0.8])
    ys = np.array([h * ny * 0.61, h * 2.0])
    zs = np.array([h * nz * 0.45, h * (nz - 2.5)])
    ref, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow, ts, xs, ys, zs, tol=1e-6, maxit=20)
    assert ierr == 0
    for algo in (0, 1, 2):
        sol = EikonalSolver(gpu_ctx, nx, ny, nz, h, tol=1e-6, maxit=20, algo=algo)
        u, _, iters, ferr = sol.solve_host(slow[None], [0], ts, xs, ys, zs, src_ptr=[0, 2])
        assert ferr[0] == 0 and iters[0] == it
        assert np.array_equal(u[0], ref), f"algo {algo}: {np.count_nonzero(u[0] != ref)} nodes differ"


@pytest.mark.parametrize("algo", [0, 2])
def test_batched_fields_two_models(gpu_ctx, algo):
    """7 fields over 2 slowness models (groups of 4, 1 and 2 slots), each equal to its own serial solve;
    fields converge after different iteration counts."""
    from mceik_b200.eikonal import EikonalSolver
    nx, ny, nz, h = 48, 40, 56, 250.0
    n = nx * ny * nz
    slow = np.stack([cases.checkerboard_slowness(nx, ny, nz, cell=8), cases.random_slowness(n, 3)])
    fmodel = np.array([0, 1, 0, 0, 1, 0, 0], dtype=np.int32)
    xs, ys, zs = cases.interior_sources(7, nx, ny, nz, h, seed=11)
    ts = np.linspace(0.0, 1.0, 7)
    ref, its = _oracle_fields(nx, ny, nz, h, slow, fmodel, xs, ys, zs, ts, 1e-6, 20)
    sol = EikonalSolver(gpu_ctx, nx, ny, nz, h, tol=1e-6, maxit=20, algo=algo)
    u, tab, iters, ferr = sol.solve_host(slow, fmodel, ts, xs, ys, zs, want_tables=True)
    assert not ferr.any()
    assert np.array_equal(iters, its)
    assert np.array_equal(u, ref)
    assert np.array_equal(tab, ref.astype(np.float32))  # SNGL(u), fsm3d.f90:1870-1872
    assert sol.node_updates == int(n * 8 * its.sum())


@pytest.mark.parametrize("publisher", ["0", "1"])
def test_both_publication_flavours_of_the_brick_kernel(gpu_ctx, publisher):
    """18 fields (more than the 16 up to which the publisher-warp flavour is the default) with either flavour
    forced: bit-equal to the oracle, so both launch paths of sweep_bricks16_kernel are parity-tested."""
    import os
    from mceik_b200.eikonal import EikonalSolver
    nx, ny, nz, h = 40, 24, 300, 200.0
    n = nx * ny * nz
    slow = np.stack([cases.checkerboard_slowness(nx, ny, nz, cell=8), cases.random_slowness(n, 21)])
    nf = 18
    fmodel = (np.arange(nf) % 2).astype(np.int32)
    xs, ys, zs = cases.interior_sources(nf, nx, ny, nz, h, seed=4)
    ts = np.zeros(nf)
    ref, its = _oracle_fields(nx, ny, nz, h, slow, fmodel[:6], xs[:6], ys[:6], zs[:6], ts[:6], 1e-6, 20)
    gpu_ctx.set_tuning("PUBLISHER", int(publisher))
    try:
        sol = EikonalSolver(gpu_ctx, nx, ny, nz, h, tol=1e-6, maxit=20)
        u, _, iters, ferr = sol.solve_host(slow, fmodel, ts, xs, ys, zs)
    finally:
        gpu_ctx.set_tuning("PUBLISHER", -1)
    assert not ferr.any()
    assert np.array_equal(iters[:6], its) and np.array_equal(u[:6], ref)
    # the remaining fields against the other flavour (default for this field count)
    sol2 = EikonalSolver(gpu_ctx, nx, ny, nz, h, tol=1e-6, maxit=20)
    u2, _, iters2, _ = sol2.solve_host(slow, fmodel, ts, xs, ys, zs)
    assert np.array_equal(u, u2) and np.array_equal(iters, iters2)


def test_pinned_output_is_copied_back_as_fields_converge(gpu_ctx):
    """Page-locked output: fields leave the device as soon as they converge (different iterations per field),
    fields stopped by maxit or refused by the boundary conditions at the end; same bits as the pageable path."""
    import torch
    from mceik_b200.eikonal import EikonalSolver
    nx, ny, nz, h = 48, 40, 56, 250.0
    n = nx * ny * nz
    slow = np.stack([cases.checkerboard_slowness(nx, ny, nz, cell=8), cases.random_slowness(n, 3)])
    fmodel = np.array([0, 1, 0, 0, 1, 0, 0, 1], dtype=np.int32)
    xs, ys, zs = cases.interior_sources(8, nx, ny, nz, h, seed=11)
    xs[5] = -10.0                                             # field 5: source outside the grid -> ierr 1
    ts = np.linspace(0.0, 1.0, 8)
    for maxit in (20, 3):
        sol = EikonalSolver(gpu_ctx, nx, ny, nz, h, tol=1e-6, maxit=maxit)
        ref, _, iters_ref, ferr_ref = sol.solve_host(slow, fmodel, ts, xs, ys, zs)
        pinned = torch.zeros((8, n), dtype=torch.float64).pin_memory()
        u, _, iters, ferr = sol.solve_host(slow, fmodel, ts, xs, ys, zs, out_u=pinned.numpy())
        assert list(ferr) == list(ferr_ref) and ferr[5] == 1 and np.array_equal(iters, iters_ref)
        ok = ferr == 0
        assert np.array_equal(u[ok], ref[ok])
        if maxit == 20:
            assert len(set(iters[ok])) > 1                    # the early copies really happen at different times
            oracle, its = _oracle_fields(nx, ny, nz, h, slow, fmodel[:2], xs[:2], ys[:2], zs[:2], ts[:2], 1e-6, 20)
            assert np.array_equal(u[:2], oracle) and np.array_equal(iters[:2], its)


@pytest.mark.parametrize("algo", [0, 2])
def test_c2_layered_128_bit_exact(gpu_ctx, algo):
    """BASELINE config 2: 128^3, 1-D layered model, one station."""
    from mceik_b200.eikonal import EikonalSolver
    nx = ny = nz = 128
    h = 1000.0
    slow = cases.layered_slowness(nx, ny, nz)
    xs, ys, zs = cases.interior_sources(1, nx, ny, nz, h, seed=1)
    ref, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow, 0.0, xs[0], ys[0], zs[0])
    sol = EikonalSolver(gpu_ctx, nx, ny, nz, h, algo=algo)
    u, _, iters, ferr = sol.solve_host(slow[None], [0], [0.0], xs, ys, zs)
    assert iters[0] == it and np.array_equal(u[0], ref)


@pytest.mark.parametrize("algo", [0, 2])
def test_maxit_cap_and_edge_sources(gpu_ctx, algo):
    """maxit smaller than needed is not an error; a source on node 1 / outside the grid gives ierr=1
    for that field only (EIKONAL_INIT_GRID quirk, fsm3d.f90:736-751); a source on node nx-1 keeps two nodes."""
    from mceik_b200.eikonal import EikonalSolver
    nx, ny, nz, h = 24, 20, 18, 50.0
    n = nx * ny * nz
    slow = cases.random_slowness(n, 5)
    xs = np.array([h * 11.3, 0.0, h * (nx - 2), -5.0])
    ys = np.array([h * 7.7, h * 5.5, h * 9.0, h * 3.0])
    zs = np.array([h * 8.1, h * 5.5, h * 9.0, h * 3.0])
    sol = EikonalSolver(gpu_ctx, nx, ny, nz, h, tol=1e-9, maxit=1, algo=algo)
    u, _, iters, ferr = sol.solve_host(slow[None], [0, 0, 0, 0], np.zeros(4), xs, ys, zs)
    assert list(ferr) == [0, 1, 0, 1]
    for f in (0, 2):
        ref, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow, 0.0, xs[f], ys[f], zs[f], tol=1e-9, maxit=1)
        assert ierr == 0 and it == 1 and iters[f] == 1
        assert np.array_equal(u[f], ref)
    _, ierr, _ = O.eikonal_serial(nx, ny, nz, h, slow, 0.0, xs[1], ys[1], zs[1])
    assert ierr == 1


@pytest.mark.parametrize("shape", [(8, 8, 8), (16, 8, 24), (24, 9, 7), (40, 16, 12), (9, 10, 11), (8, 3, 5), (32, 8, 260),
                                   (3, 3, 3)])
def test_seeded_random_sources_anywhere(gpu_ctx, shape):
    """Six fields per grid with 1-3 sources each, dropped anywhere from half a cell outside the grid to half a cell
    outside the far faces, random start times, two slowness models: fields whose stencil leaves the grid fail with
    ierr = 1 exactly where the oracle's SETBCS fails (fsm3d.f90:716-755), every other field and its iteration count
    equal the oracle bit for bit.  Grids with and without nx % 8 == 0, thinner than a brick, taller than one."""
    from mceik_b200.eikonal import EikonalSolver
    nx, ny, nz = shape
    h = 125.0
    n = nx * ny * nz
    rng = np.random.default_rng(1000 * nx + 10 * ny + nz)
    slow = np.stack([cases.random_slowness(n, seed=nx + ny), cases.checkerboard_slowness(nx, ny, nz, cell=3)])
    nf = 6
    fmodel = rng.integers(0, 2, nf).astype(np.int32)
    nsrc = rng.integers(1, 4, nf)
    src_ptr = np.concatenate([[0], np.cumsum(nsrc)]).astype(np.int32)
    tot = int(src_ptr[-1])
    xs = rng.uniform(-0.5 * h, (nx - 0.5) * h, tot)
    ys = rng.uniform(-0.5 * h, (ny - 0.5) * h, tot)
    zs = rng.uniform(-0.5 * h, (nz - 0.5) * h, tot)
    for q in range(0, tot, 3):  # every third source well inside, so that some fields certainly solve
        xs[q], ys[q], zs[q] = (rng.uniform(1.2 * h, (d - 2.2) * h) if d > 4 else 1.5 * h for d in (nx, ny, nz))
    ts = rng.uniform(0.0, 1.0, tot)
    sol = EikonalSolver(gpu_ctx, nx, ny, nz, h, tol=1e-6, maxit=20)
    u, _, iters, ferr = sol.solve_host(slow, fmodel, ts, xs, ys, zs, src_ptr=src_ptr)
    solved = 0
    for f in range(nf):
        a, b = src_ptr[f], src_ptr[f + 1]
        ref, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow[fmodel[f]], ts[a:b], xs[a:b], ys[a:b], zs[a:b])
        assert (ierr != 0) == (ferr[f] != 0), f"field {f}: oracle ierr {ierr}, library {ferr[f]}"
        if ierr == 0:
            solved += 1
            assert it == iters[f], f"field {f}"
            assert np.array_equal(u[f], ref), f"field {f}: {np.count_nonzero(u[f] != ref)} nodes differ"
    assert solved >= 1 or min(shape) <= 3


def test_argument_errors_and_empty_batch(gpu_ctx):
    """Bad sizes are refused before anything is allocated; an empty batch is a no-op."""
    from mceik_b200 import _lib
    from mceik_b200.eikonal import EikonalSolver
    nx, ny, nz, h = 16, 8, 8, 10.0
    slow = np.full((1, nx * ny * nz), 1e-3)
    sol = EikonalSolver(gpu_ctx, nx, ny, nz, h)
    u, _, iters, ferr = sol.solve_host(slow, np.zeros(0, np.int32), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0))
    assert u.shape[0] == 0 and iters.size == 0 and sol.node_updates == 0
    with pytest.raises(_lib.MceikError, match="field_model"):
        sol.solve_host(slow, [1], [0.0], [55.0], [33.0], [41.0])
    big = EikonalSolver(gpu_ctx, 2048, 2048, 512, h)   # 2^31 nodes: refused, nothing is dereferenced
    import ctypes as C
    one_i, one_d = (C.c_int * 2)(0, 1), (C.c_double * 1)(55.0)
    rc = big.lib.mceik_fsm_solve_batched_dev(gpu_ctx.handle, C.byref(big.grid), 1, C.c_void_p(256), 1, one_i, one_i, one_d,
                                             one_d, one_d, one_d, None, None, 0, one_i, one_i)
    assert rc != 0 and "2^31" in _lib.last_error()


def test_serial_driver_lifecycle(gpu_ctx):
    """Drop-in eikonal3d_serial_driver: job 1/2/3, double init and solve-before-init give ierr=1
    (fsm3d.f90:1993-2013); eikonal3d_initialize/solve/finalize give the same field."""
    from mceik_b200 import eikonal as E
    nx, ny, nz, h = 30, 26, 22, 200.0
    n = nx * ny * nz
    slow = cases.random_slowness(n, 9)
    xs, ys, zs = cases.interior_sources(1, nx, ny, nz, h, seed=2)
    u = np.zeros(n)
    args = (0, 20, 1, nx, ny, nz, 1e-6, h, 0.0, 0.0, 0.0, [0.25], xs, ys, zs, slow, u)
    assert E.eikonal3d_serial_driver(2, *args) == 1
    assert E.eikonal3d_serial_driver(1, *args) == 0
    assert E.eikonal3d_serial_driver(1, *args) == 1
    assert E.eikonal3d_serial_driver(2, *args) == 0
    assert E.eikonal3d_serial_driver(3, *args) == 0
    ref, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow, 0.25, xs[0], ys[0], zs[0])
    assert np.array_equal(u, ref)
    u2 = np.zeros(n)
    assert E.eikonal3d_solve(0, 1, n, [0.25], xs, ys, zs, slow, u2) == 1  # before initialize
    assert E.eikonal3d_initialize(0, 0, nx, ny, nz, 2, 2, 2, 1, 20, 0.0, 0.0, 0.0, h, 1e-6) == 0
    assert E.eikonal3d_solve(0, 1, n, [0.25], xs, ys, zs, slow, u2) == 0
    assert E.eikonal3d_finalize(0) == 0
    assert E.eikonal3d_finalize(0) == 1
    assert np.array_equal(u2, ref)


def test_homogeneous_tables(gpu_ctx):
    """computeHomogeneousTraveltimes drop-in vs the oracle restatement of homog.c:594-621."""
    from mceik_b200 import eikonal as E
    nx, ny, nz = 32, 29, 26
    t = E.compute_homogeneous_traveltimes(nx, ny, nz, 0.0, 0.0, 0.0, 1000.0, 1000.0, 1000.0, 7000.0, 12000.0, 25000.0, 2000.0)
    ref = O.homogeneous_traveltimes(nx, ny, nz, 0.0, 0.0, 0.0, 1000.0, 1000.0, 1000.0, 7000.0, 12000.0, 25000.0, 2000.0)
    assert np.array_equal(t, ref)


@pytest.mark.parametrize("natural", [0, 1])
def test_blocked_and_natural_layouts(gpu_ctx, natural):
    """sweep_bricks16_kernel on its blocked layout (default: 512-byte brick planes + x-face copies) and on the
    caller's [z][y][x] layout: 9 fields over 2 models, several sources in one field, a grid with partial bricks in y
    and more than one brick along z.  Bit-equal to the oracle."""
    from mceik_b200.eikonal import EikonalSolver
    nx, ny, nz, h = 40, 28, 300, 200.0
    n = nx * ny * nz
    slow = np.stack([cases.checkerboard_slowness(nx, ny, nz, cell=8), cases.random_slowness(n, 5)])
    nf = 9
    fmodel = np.array([0, 1, 0, 0, 1, 0, 1, 0, 0], np.int32)
    xs, ys, zs = cases.interior_sources(nf + 1, nx, ny, nz, h, seed=12)
    src_ptr = np.array([0, 2] + list(range(3, nf + 2)), np.int32)  # field 0 has two sources
    ts = np.linspace(0.0, 0.5, nf + 1)
    gpu_ctx.set_tuning("NATURAL", natural)
    try:
        sol = EikonalSolver(gpu_ctx, nx, ny, nz, h, tol=1e-6, maxit=20)
        u, _, iters, ferr = sol.solve_host(slow, fmodel, ts, xs, ys, zs, src_ptr=src_ptr)
    finally:
        gpu_ctx.set_tuning("NATURAL", 0)
    assert not ferr.any()
    for f in range(nf):
        a, b = src_ptr[f], src_ptr[f + 1]
        ref, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow[fmodel[f]], ts[a:b], xs[a:b], ys[a:b], zs[a:b])
        assert ierr == 0 and it == iters[f], f"field {f}"
        assert np.array_equal(u[f], ref), f"field {f}: {np.count_nonzero(u[f] != ref)} nodes differ"


def test_many_stacked_sources_take_the_generic_brick_kernel(gpu_ctx):
    """A field with 24 sources stacked along z inside one brick column has boundary-condition nodes on more than 16
    planes of that brick -- more than the list sweep_bricks16_kernel keeps in shared memory -- so the host routes
    the solve to the generic brick kernel; a second field with one source rides along.  Bit-equal to the oracle."""
    from mceik_b200.eikonal import EikonalSolver
    nx, ny, nz, h = 32, 24, 64, 100.0
    n = nx * ny * nz
    slow = cases.checkerboard_slowness(nx, ny, nz, cell=8)
    ns = 24
    xs = np.concatenate([np.full(ns, 1234.0), [2020.0]])
    ys = np.concatenate([np.full(ns, 1111.0), [777.0]])
    zs = np.concatenate([150.0 + 230.0 * np.arange(ns), [3333.0]])
    ts = np.concatenate([0.01 * np.arange(ns), [0.0]])
    src_ptr = np.array([0, ns, ns + 1], np.int32)
    sol = EikonalSolver(gpu_ctx, nx, ny, nz, h)
    u, _, iters, ferr = sol.solve_host(slow[None], [0, 0], ts, xs, ys, zs, src_ptr=src_ptr)
    assert not ferr.any()
    for f in range(2):
        a, b = src_ptr[f], src_ptr[f + 1]
        ref, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow, ts[a:b], xs[a:b], ys[a:b], zs[a:b])
        assert ierr == 0 and it == iters[f]
        assert np.array_equal(u[f], ref), f"field {f}"


def test_solver_building_blocks_selftest(gpu_ctx):
    """sqrt_fast == __dsqrt_rn and the straight-line solver == the reference-ordered solver, bit for
    bit, on 4e8 pseudo-random inputs each (incl. exact squares, ties and u_nan neighbours)."""
    import ctypes as C
    from mceik_b200 import _lib
    b1, b2 = C.c_longlong(-1), C.c_longlong(-1)
    for seed in (1, 2):
        rc = _lib.load().mceik_selftest_solver(gpu_ctx.handle, seed, 200_000_000, C.byref(b1), C.byref(b2))
        assert rc == 0 and b1.value == 0 and b2.value == 0, (b1.value, b2.value)


def test_full_size_256_all_kernels_agree(gpu_ctx):
    """BASELINE config 3 grid (256^3 checkerboard), 3 fields over the P and S models: the four sweep
    kernels (16-byte bricks, generic bricks, tiles, per-hyperplane cross-check) give bit-identical
    fields and iteration counts, equal to the oracle's for a P and an S field; size-independent properties of the solution hold (stencil nodes keep
    ts + d*slow, times are finite, non-negative and bounded by the slowest straight ray)."""
    import os
    import torch
    from mceik_b200.eikonal import EikonalSolver
    n, h = 256, 1000.0
    N = n ** 3
    slow = np.stack([cases.checkerboard_slowness(n, n, n, cell=32), cases.checkerboard_slowness(n, n, n, cell=32, vs=True)])
    xs, ys, zs = cases.interior_sources(3, n, n, n, h, seed=3)
    fmodel = np.array([0, 1, 0], np.int32)
    d_slow = torch.from_numpy(slow).cuda()
    outs = []
    for algo, env in ((2, {}), (2, {"NO16": 1}), (0, {}), (1, {})):
        for k, v in env.items():
            gpu_ctx.set_tuning(k, v)
        try:
            d_u = torch.empty((3, N), dtype=torch.float64, device="cuda")
            sol = EikonalSolver(gpu_ctx, n, n, n, h, algo=algo)
            iters, ferr = sol.solve_device(d_slow, fmodel, np.zeros(3), xs, ys, zs, d_u=d_u)
            gpu_ctx.synchronize()
            outs.append((d_u, iters.copy()))
            assert not ferr.any()
        finally:
            for k in env:
                gpu_ctx.set_tuning(k, 0)
    for d_u, iters in outs[1:]:
        assert np.array_equal(iters, outs[0][1])
        assert torch.equal(d_u, outs[0][0])
    u = outs[0][0]
    # ... and the oracle at this size: one P and one S field, bit for bit, with the oracle's iteration counts
    O.set_threads(0)
    for f in (0, 1):
        ref, ierr, it = O.eikonal_serial(n, n, n, h, slow[fmodel[f]], 0.0, xs[f], ys[f], zs[f], tol=1e-6, maxit=20)
        assert ierr == 0 and it == outs[0][1][f]
        assert np.array_equal(u[f].cpu().numpy(), ref), f"field {f}: differs from the oracle at 256^3"
    assert bool(torch.isfinite(u).all()) and float(u.min()) >= 0.0
    diag = np.sqrt(3.0) * (n - 1) * h
    assert float(u[0].max()) <= 1.5 * diag * slow[0].max() and float(u[1].max()) <= 1.5 * diag * slow[1].max()
    for f in range(3):
        ix, iy, iz = int(xs[f] // h), int(ys[f] // h), int(zs[f] // h)
        node = (iz * n + iy) * n + ix
        d = np.sqrt((xs[f] - ix * h) ** 2 + (ys[f] - iy * h) ** 2 + (zs[f] - iz * h) ** 2)
        assert float(u[f, node]) == 0.0 + d * slow[fmodel[f]][node]
