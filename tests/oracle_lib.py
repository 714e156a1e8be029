"""ctypes bindings to the CPU checkers under oracle/ (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package mceik_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
_REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_locate.so")

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_flt_p = C.POINTER(C.c_float)


def build(force=False):
    """(Re)build liboracle.so and, when /root/reference exists, oracle/_ref/."""
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("fsm3d_oracle.c", "locate_oracle.c", "Makefile")]
    stale = force or not os.path.exists(_ORACLE_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_ORACLE_SO) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "all"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/locate.c") and (force or not os.path.exists(_REF_SO)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def aligned(n, dtype, align=64):
    """numpy array of n elements whose data pointer is `align`-byte aligned."""
    dt = np.dtype(dtype)
    raw = np.zeros(n * dt.itemsize + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * dt.itemsize].view(dt)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_ORACLE_SO)
        _lib.oracle_last_iterations.restype = C.c_int
    return _lib


def ref():
    """The reference's own locate.c (oracle/_ref); None when it was never built."""
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(_REF_SO):
            return None
        _ref = C.CDLL(_REF_SO)
    return _ref


# ---------------------------------------------------------------- eikonal
def eikonal_serial(nx, ny, nz, h, slow, ts, xs, ys, zs, tol=1e-6, maxit=20, x0=0.0, y0=0.0, z0=0.0):
    """oracle restatement of eikonal3d_serial_driver jobs 1,2,3 -> (u, ierr, iterations)."""
    L = lib()
    slow = np.ascontiguousarray(slow, dtype=np.float64).ravel()
    ts, xs, ys, zs = (np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64) for a in (ts, xs, ys, zs))
    u = np.empty(nx * ny * nz, dtype=np.float64)
    ierr = C.c_int(0)
    ci = lambda v: C.byref(C.c_int(v))
    cd = lambda v: C.byref(C.c_double(v))

    def call(job):
        L.oracle_eikonal3d_serial_driver(ci(job), ci(0), ci(maxit), ci(len(ts)), ci(nx), ci(ny), ci(nz),
                                         cd(tol), cd(h), cd(x0), cd(y0), cd(z0),
                                         _p(ts, c_dbl_p), _p(xs, c_dbl_p), _p(ys, c_dbl_p), _p(zs, c_dbl_p),
                                         _p(slow, c_dbl_p), _p(u, c_dbl_p), C.byref(ierr))
        return ierr.value

    if call(1) != 0:
        raise RuntimeError("oracle driver init failed")
    e2 = call(2)
    iters = L.oracle_last_iterations()
    call(3)
    return u, e2, iters


def eikonal_serial_timed(nx, ny, nz, h, slow, ts, xs, ys, zs, tol=1e-6, maxit=20):
    """As eikonal_serial, timing job 2 (boundary conditions + sweeps) only -- the level structure of job 1 is built
    once per grid by the reference's callers (fsm3d.f90:1985-1998).  Returns (u, ierr, iterations, seconds)."""
    import time
    L = lib()
    slow = np.ascontiguousarray(slow, dtype=np.float64).ravel()
    ts, xs, ys, zs = (np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64) for a in (ts, xs, ys, zs))
    u = np.empty(nx * ny * nz, dtype=np.float64)
    ierr = C.c_int(0)
    ci = lambda v: C.byref(C.c_int(v))
    cd = lambda v: C.byref(C.c_double(v))

    def call(job):
        L.oracle_eikonal3d_serial_driver(ci(job), ci(0), ci(maxit), ci(len(ts)), ci(nx), ci(ny), ci(nz),
                                         cd(tol), cd(h), cd(0.0), cd(0.0), cd(0.0),
                                         _p(ts, c_dbl_p), _p(xs, c_dbl_p), _p(ys, c_dbl_p), _p(zs, c_dbl_p),
                                         _p(slow, c_dbl_p), _p(u, c_dbl_p), C.byref(ierr))
        return ierr.value

    if call(1) != 0:
        raise RuntimeError("oracle driver init failed")
    t = time.perf_counter()
    e2 = call(2)
    dt = time.perf_counter() - t
    iters = L.oracle_last_iterations()
    call(3)
    return u, e2, iters, dt


def hamiltonian3d(a, b, c, f):
    L = lib()
    L.oracle_hamiltonian3d.restype = C.c_double
    ierr = C.c_int(0)
    v = L.oracle_hamiltonian3d(C.c_double(a), C.c_double(b), C.c_double(c), C.c_double(f), C.byref(ierr))
    return v, ierr.value


def homogeneous_traveltimes(nx, ny, nz, x0, y0, z0, dx, dy, dz, xs, ys, zs, vel):
    t = np.empty(nx * ny * nz, dtype=np.float64)
    lib().oracle_homogeneous_traveltimes(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_double(x0), C.c_double(y0),
                                         C.c_double(z0), C.c_double(dx), C.c_double(dy), C.c_double(dz),
                                         C.c_double(xs), C.c_double(ys), C.c_double(zs), C.c_double(vel),
                                         _p(t, c_dbl_p))
    return t


# ---------------------------------------------------------------- grid search
def _l2(fn, ldgrd, ngrd, nobs, iwantOT, t0use, mask, tobs, tcorr, varobs, test, dtype):
    ct = C.c_double if dtype == np.float64 else C.c_float
    pt = C.POINTER(ct)
    t0 = aligned(max(ngrd, 1), dtype)
    obj = aligned(max(ngrd, 1), dtype)
    fn.restype = C.c_int
    rc = fn(C.c_int(ldgrd), C.c_int(ngrd), C.c_int(nobs), C.c_int(iwantOT), ct(t0use),
            _p(mask, c_int_p), _p(tobs, pt), _p(tcorr, pt), _p(varobs, pt), _p(test, pt), _p(t0, pt), _p(obj, pt))
    return rc, t0, obj


def l2_gridsearch(ldgrd, ngrd, nobs, iwantOT, t0use, mask, tobs, tcorr, varobs, test, dtype=np.float64, use_ref=False):
    """locate_l2_gridSearch__{double64,float64}: oracle restatement, or the reference's own code."""
    if use_ref:
        fn = getattr(ref(), "locate_l2_gridSearch__double64" if dtype == np.float64 else "locate_l2_gridSearch__float64")
    else:
        fn = getattr(lib(), "oracle_l2_gridsearch_f64" if dtype == np.float64 else "oracle_l2_gridsearch_f32")
    return _l2(fn, ldgrd, ngrd, nobs, iwantOT, t0use, mask, tobs, tcorr, varobs, test, dtype)


def minloc(x, use_ref=False):
    x = np.ascontiguousarray(x)
    if x.dtype == np.float64:
        fn = ref().locate_minLocDouble64 if use_ref else lib().oracle_minloc_f64
        return fn(C.c_int(x.size), _p(x, c_dbl_p))
    fn = ref().locate_minLocFloat64 if use_ref else lib().oracle_minloc_f32
    return fn(C.c_int(x.size), _p(x, c_flt_p))


def gridsearch_f90(ldgrd, ngrd, nobs, iwantOT, mask, tobs, varobs, test, dtype=np.float64):
    """locate3d_gridsearch__{double64,float64} (gridsearch.f90 flavour) -> (ierr, logPDF, t0)."""
    pt = c_dbl_p if dtype == np.float64 else c_flt_p
    fn = lib().oracle_gridsearch_f90_f64 if dtype == np.float64 else lib().oracle_gridsearch_f90_f32
    out = aligned(max(ngrd, 1), dtype)
    t0 = aligned(max(ngrd, 1), dtype)
    ierr = C.c_int(0)
    fn(C.c_int(ldgrd), C.c_int(ngrd), C.c_int(nobs), C.c_int(iwantOT), _p(mask, c_int_p), _p(tobs, pt),
       _p(varobs, pt), _p(test, pt), _p(out, pt), _p(t0, pt), C.byref(ierr))
    return ierr.value, out, t0


def locate3d_catalog(job, ngrd, ldgrd, tables, nobs, nevents, luseObs, statPtr, pickType, statCor, tori,
                     varobs, tobs, xlocs, ylocs, zlocs):
    """Catalogue contract of locate.f90:322-519 with the canonical arithmetic -> (rc, hypo, iopt, objmin)."""
    L = lib()
    L.oracle_locate3d_catalog.restype = C.c_int
    tables = np.ascontiguousarray(tables, dtype=np.float32)
    ntables = tables.size // ldgrd
    hypo = np.zeros(4 * nevents, dtype=np.float64)
    iopt = np.zeros(nevents, dtype=np.int32)
    objmin = np.zeros(nevents, dtype=np.float64)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    a = [i32(luseObs), i32(statPtr), i32(pickType), f64(statCor), f64(tori), f64(varobs), f64(tobs),
         f32(xlocs), f32(ylocs), f32(zlocs)]
    rc = L.oracle_locate3d_catalog(C.c_int(job), C.c_int(ngrd), C.c_size_t(ldgrd), C.c_int(ntables),
                                   _p(tables, c_flt_p), C.c_int(nobs), C.c_int(nevents),
                                   _p(a[0], c_int_p), _p(a[1], c_int_p), _p(a[2], c_int_p), _p(a[3], c_dbl_p),
                                   _p(a[4], c_dbl_p), _p(a[5], c_dbl_p), _p(a[6], c_dbl_p), _p(a[7], c_flt_p),
                                   _p(a[8], c_flt_p), _p(a[9], c_flt_p), _p(hypo, c_dbl_p), _p(iopt, c_int_p),
                                   _p(objmin, c_dbl_p))
    return rc, hypo, iopt, objmin


def event_logpdf(job, ngrd, ldgrd, tables, table_id, tobs_cor, varobs, tori=0.0):
    """Per-event posterior volume (catalogue flavour) -> (rc, logPDF, t0 grid)."""
    L = lib()
    L.oracle_event_logpdf.restype = C.c_int
    tables = np.ascontiguousarray(tables, dtype=np.float32)
    tid = np.ascontiguousarray(table_id, dtype=np.int32)
    tc = np.ascontiguousarray(tobs_cor, dtype=np.float64)
    var = np.ascontiguousarray(varobs, dtype=np.float64)
    pdf, t0 = np.empty(ngrd), np.empty(ngrd)
    rc = L.oracle_event_logpdf(C.c_int(job), C.c_int(ngrd), C.c_size_t(ldgrd), _p(tables, c_flt_p), C.c_int(tid.size),
                               _p(tid, c_int_p), _p(tc, c_dbl_p), _p(var, c_dbl_p), C.c_double(tori), _p(pdf, c_dbl_p),
                               _p(t0, c_dbl_p))
    return rc, pdf, t0


def set_threads(n=0):
    """OpenMP threads of the FSM oracle (n <= 0: query only) -> threads the next solve uses."""
    L = lib()
    L.oracle_set_threads.restype = C.c_int
    return int(L.oracle_set_threads(C.c_int(int(n))))


def weighted_median(x, w):
    L = lib()
    L.oracle_weighted_median.restype = C.c_double
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    return float(L.oracle_weighted_median(C.c_int(x.size), _p(x, c_dbl_p), _p(w, c_dbl_p)))


def l1_gridsearch(ldgrd, ngrd, nobs, iwantOT, t0use, mask, tobs, varobs, test):
    """L1 flavour (locate.c:1205-1335, parity unpinned) -> (rc, t0, objfn)."""
    L = lib()
    L.oracle_l1_gridsearch_f64.restype = C.c_int
    mask = np.ascontiguousarray(mask, dtype=np.int32)
    tobs, varobs, test = (np.ascontiguousarray(a, dtype=np.float64) for a in (tobs, varobs, test))
    t0, obj = np.zeros(ngrd), np.zeros(ngrd)
    rc = L.oracle_l1_gridsearch_f64(C.c_int(ldgrd), C.c_int(ngrd), C.c_int(nobs), C.c_int(iwantOT), C.c_double(t0use),
                                    _p(mask, c_int_p), _p(tobs, c_dbl_p), _p(varobs, c_dbl_p), _p(test, c_dbl_p),
                                    _p(t0, c_dbl_p), _p(obj, c_dbl_p))
    return rc, t0, obj
