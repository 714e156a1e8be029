"""Host-side mirror of the reference's eikonal entry points (fsm3d.f90) over the C ABI.

Drop-in style functions take numpy arrays and return ``ierr`` like the Fortran ``BIND(C)``
routines; :class:`EikonalSolver` is the batched form (many stations / velocity models per call,
host numpy arrays or device-resident torch tensors).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FsmGrid, c_dbl_p, c_flt_p, c_int_p

ALGO_TILES, ALGO_LEVELS, ALGO_BRICKS = 0, 1, 2


def _i(v):
    return C.byref(C.c_int(int(v)))


def _d(v):
    return C.byref(C.c_double(float(v)))


def _f64(a):
    return np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


def eikonal3d_serial_driver(job, iverb, maxit, nsrc, nx, ny, nz, tol, h, x0, y0, z0, ts, xs, ys, zs, slow, u):
    """``eikonal3d_serial_driver`` (fsm3d.f90:1968-2052): job 1 init, 2 solve (fills ``u`` in
    place), 3 finalize.  Returns ierr."""
    lib = _lib.load()
    ts, xs, ys, zs = (_f64(a) for a in (ts, xs, ys, zs))
    slow = _f64(slow).ravel()
    assert u.dtype == np.float64 and u.flags.c_contiguous
    ierr = C.c_int(0)
    lib.eikonal3d_serial_driver(_i(job), _i(iverb), _i(maxit), _i(nsrc), _i(nx), _i(ny), _i(nz), _d(tol), _d(h),
                                _d(x0), _d(y0), _d(z0), _ptr(ts, c_dbl_p), _ptr(xs, c_dbl_p), _ptr(ys, c_dbl_p),
                                _ptr(zs, c_dbl_p), _ptr(slow, c_dbl_p), _ptr(u, c_dbl_p), C.byref(ierr))
    return ierr.value


def eikonal3d_initialize(comm, iverb, nx, ny, nz, ndivx, ndivy, ndivz, noverlap, maxit, x0, y0, z0, h, tol):
    """``eikonal3d_initialize`` (fsm3d.f90:1583-1674).  Returns ierr."""
    ierr = C.c_int(0)
    _lib.load().eikonal3d_initialize(_i(comm), _i(iverb), _i(nx), _i(ny), _i(nz), _i(ndivx), _i(ndivy), _i(ndivz),
                                     _i(noverlap), _i(maxit), _d(x0), _d(y0), _d(z0), _d(h), _d(tol), C.byref(ierr))
    return ierr.value


def eikonal3d_solve(comm, nsrc, n, ts, xs, ys, zs, slow, u):
    """``eikonal3d_solve`` (fsm3d.f90:1754-1840).  Returns ierr."""
    ts, xs, ys, zs = (_f64(a) for a in (ts, xs, ys, zs))
    slow = _f64(slow).ravel()
    ierr = C.c_int(0)
    _lib.load().eikonal3d_solve(_i(comm), _i(nsrc), _i(n), _ptr(ts, c_dbl_p), _ptr(xs, c_dbl_p), _ptr(ys, c_dbl_p),
                                _ptr(zs, c_dbl_p), _ptr(slow, c_dbl_p), _ptr(u, c_dbl_p), C.byref(ierr))
    return ierr.value


def eikonal3d_finalize(comm=0):
    ierr = C.c_int(0)
    _lib.load().eikonal3d_finalize(_i(comm), C.byref(ierr))
    return ierr.value


def compute_homogeneous_traveltimes(nx, ny, nz, x0, y0, z0, dx, dy, dz, xs, ys, zs, vel):
    """``computeHomogeneousTraveltimes`` (homog.c:594-621) -> fp64 array [nx*ny*nz]."""
    t = np.empty(nx * ny * nz, dtype=np.float64)
    rc = _lib.load().computeHomogeneousTraveltimes(nx, ny, nz, x0, y0, z0, dx, dy, dz, xs, ys, zs, vel, _ptr(t, c_dbl_p))
    if rc != 0:
        raise _lib.MceikError(f"computeHomogeneousTraveltimes failed: {_lib.last_error()}")
    return t


class EikonalSolver:
    """Batched fast-sweeping solver: ``nfields`` independent (sources, slowness model) problems on
    one grid per call.  Geometry / solver parameters follow ``solverParametersType``
    (module.F90:117-131)."""

    def __init__(self, ctx, nx, ny, nz, h, x0=0.0, y0=0.0, z0=0.0, tol=1e-6, maxit=20, algo=ALGO_BRICKS):
        self.ctx = ctx
        self.lib = _lib.load()
        self.grid = FsmGrid(int(nx), int(ny), int(nz), float(h), float(x0), float(y0), float(z0), float(tol), int(maxit))
        self.n = int(nx) * int(ny) * int(nz)
        self.algo = algo
        self.last_iters = None
        self.last_ierr = None

    def _sources(self, nfields, ts, xs, ys, zs, src_ptr):
        ts, xs, ys, zs = (_f64(a) for a in (ts, xs, ys, zs))
        if src_ptr is None:  # one source per field
            src_ptr = np.arange(nfields + 1, dtype=np.int32)
        src_ptr = np.ascontiguousarray(src_ptr, dtype=np.int32)
        assert src_ptr.size == nfields + 1 and ts.size == xs.size == ys.size == zs.size >= src_ptr[-1]
        return ts, xs, ys, zs, src_ptr

    @property
    def node_updates(self):
        """N * 8 * iterations summed over the fields of the last solve."""
        return int(self.lib.mceik_fsm_last_node_updates(self.ctx.handle))

    @property
    def sweep_stats(self):
        """(device ms spent in the sweep kernel, number of sweep-kernel launches) of the last solve."""
        ms, n = C.c_double(0.0), C.c_int(0)
        self.lib.mceik_fsm_last_sweep_stats(self.ctx.handle, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def solve_host(self, slow, field_model, ts, xs, ys, zs, src_ptr=None, want_u=True, want_tables=False, ldtab=None,
                   out_u=None):
        """Host (numpy) buffers in and out.  slow: [nmodels, N] fp64.  Returns (u, tables, iters, ierr).
        out_u: optional caller-owned [nfields, N] fp64 array for the fields; when it is page-locked, every field is
        copied back as soon as it has converged, overlapped with the iterations of the others."""
        slow = np.ascontiguousarray(slow, dtype=np.float64).reshape(-1, self.n)
        field_model = np.ascontiguousarray(field_model, dtype=np.int32)
        nf = field_model.size
        ts, xs, ys, zs, src_ptr = self._sources(nf, ts, xs, ys, zs, src_ptr)
        ldtab = self.n if ldtab is None else int(ldtab)
        u = np.empty((nf, self.n), dtype=np.float64) if want_u else None
        if out_u is not None:
            assert out_u.dtype == np.float64 and out_u.flags.c_contiguous and out_u.size == nf * self.n
            u = out_u
        tab = np.zeros((nf, ldtab), dtype=np.float32) if want_tables else None
        iters = np.zeros(nf, dtype=np.int32)
        ferr = np.zeros(nf, dtype=np.int32)
        self.lib.mceik_fsm_set_algo(self.ctx.handle, self.algo)
        rc = self.lib.mceik_fsm_solve_batched_host(
            self.ctx.handle, C.byref(self.grid), slow.shape[0], _ptr(slow, c_dbl_p), nf, _ptr(field_model, c_int_p),
            _ptr(src_ptr, c_int_p), _ptr(ts, c_dbl_p), _ptr(xs, c_dbl_p), _ptr(ys, c_dbl_p), _ptr(zs, c_dbl_p),
            _ptr(u, c_dbl_p), _ptr(tab, c_flt_p), ldtab, _ptr(iters, c_int_p), _ptr(ferr, c_int_p))
        _lib.check(rc, "mceik_fsm_solve_batched_host")
        self.last_iters, self.last_ierr = iters, ferr
        return u, tab, iters, ferr

    def solve_device(self, d_slow, field_model, ts, xs, ys, zs, d_u=None, d_tables=None, src_ptr=None):
        """Device-resident torch tensors: d_slow [nmodels, N] fp64, d_u [nfields, N] fp64 (optional),
        d_tables [nfields, ldtab] fp32 (optional).  Runs on the context stream.  Returns (iters, ierr)."""
        field_model = np.ascontiguousarray(field_model, dtype=np.int32)
        nf = field_model.size
        ts, xs, ys, zs, src_ptr = self._sources(nf, ts, xs, ys, zs, src_ptr)
        assert d_slow.is_cuda and d_slow.is_contiguous() and d_slow.dtype.itemsize == 8 and d_slow.numel() % self.n == 0
        nmodels = d_slow.numel() // self.n
        ldtab = 0
        if d_u is not None:
            assert d_u.is_cuda and d_u.is_contiguous() and d_u.numel() == nf * self.n and d_u.dtype.itemsize == 8
        if d_tables is not None:
            assert d_tables.is_cuda and d_tables.is_contiguous() and d_tables.dtype.itemsize == 4
            ldtab = d_tables.numel() // nf
        iters = np.zeros(nf, dtype=np.int32)
        ferr = np.zeros(nf, dtype=np.int32)
        self.lib.mceik_fsm_set_algo(self.ctx.handle, self.algo)
        rc = self.lib.mceik_fsm_solve_batched_dev(
            self.ctx.handle, C.byref(self.grid), nmodels, C.c_void_p(d_slow.data_ptr()), nf, _ptr(field_model, c_int_p),
            _ptr(src_ptr, c_int_p), _ptr(ts, c_dbl_p), _ptr(xs, c_dbl_p), _ptr(ys, c_dbl_p), _ptr(zs, c_dbl_p),
            C.c_void_p(d_u.data_ptr()) if d_u is not None else None,
            C.c_void_p(d_tables.data_ptr()) if d_tables is not None else None, ldtab,
            _ptr(iters, c_int_p), _ptr(ferr, c_int_p))
        _lib.check(rc, "mceik_fsm_solve_batched_dev")
        self.last_iters, self.last_ierr = iters, ferr
        return iters, ferr

    def solve_sharded(self, d_slow, field_model, ts, xs, ys, zs, d_tables_all, cost=None, src_ptr=None):
        """``mceik_fsm_solve_sharded_dev`` (collective over the context's communicator): all arguments describe ALL
        fields and are the same on every rank; d_tables_all [world * slots, ldtab] fp32 receives every field's table
        on every rank.  Returns (iters, ierr, table_row) for all fields."""
        field_model = np.ascontiguousarray(field_model, dtype=np.int32)
        nf = field_model.size
        ts, xs, ys, zs, src_ptr = self._sources(nf, ts, xs, ys, zs, src_ptr)
        assert d_slow.is_cuda and d_slow.is_contiguous() and d_slow.dtype.itemsize == 8 and d_slow.numel() % self.n == 0
        assert d_tables_all.is_cuda and d_tables_all.is_contiguous() and d_tables_all.dtype.itemsize == 4
        ldtab = d_tables_all.shape[-1]
        iters, ferr, row = (np.zeros(nf, dtype=np.int32) for _ in range(3))
        c = None if cost is None else np.ascontiguousarray(cost, dtype=np.int32)
        self.lib.mceik_fsm_set_algo(self.ctx.handle, self.algo)
        rc = self.lib.mceik_fsm_solve_sharded_dev(
            self.ctx.handle, C.byref(self.grid), d_slow.numel() // self.n, C.c_void_p(d_slow.data_ptr()), nf,
            _ptr(field_model, c_int_p), _ptr(src_ptr, c_int_p), _ptr(ts, c_dbl_p), _ptr(xs, c_dbl_p), _ptr(ys, c_dbl_p),
            _ptr(zs, c_dbl_p), _ptr(c, c_int_p), C.c_void_p(d_tables_all.data_ptr()), ldtab, _ptr(iters, c_int_p),
            _ptr(ferr, c_int_p), _ptr(row, c_int_p))
        _lib.check(rc, "mceik_fsm_solve_sharded_dev")
        self.last_iters, self.last_ierr = iters, ferr
        return iters, ferr, row
