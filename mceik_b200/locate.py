"""Host-side mirror of the reference's L2 grid-search entry points over the C ABI.

* ``locate_l2_gridSearch__double64/float64``, ``locate_minLocDouble64/Float64`` -- locate.c:811-1203
* ``locate3d_gridsearch__double64/float64`` -- gridsearch.f90:382-540
* ``locate3d_initialize/gridsearch/finalize`` -- locate.f90:322-689 (catalogue contract)
* :class:`Locator` -- batched events against tables resident in HBM.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CatalogStruct, StationsStruct, c_dbl_p, c_flt_p, c_int_p


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _i(v):
    return C.byref(C.c_int(int(v)))


def aligned_empty(n, dtype, align=64):
    """numpy array whose data pointer is 64-byte aligned (the reference's alignment contract,
    locate.c:967-974)."""
    dt = np.dtype(dtype)
    raw = np.zeros(n * dt.itemsize + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * dt.itemsize].view(dt)


def locate_l2_gridSearch__double64(ldgrd, ngrd, nobs, iwantOT, t0use, mask, tobs, tcorr, varobs, test, t0, objfn):
    """locate.c:923-1047.  Fills t0/objfn in place, returns the C return code (0 ok, 1 error)."""
    return _lib.load().locate_l2_gridSearch__double64(
        int(ldgrd), int(ngrd), int(nobs), int(iwantOT), float(t0use), _ptr(mask, c_int_p), _ptr(tobs, c_dbl_p),
        _ptr(tcorr, c_dbl_p), _ptr(varobs, c_dbl_p), _ptr(test, c_dbl_p), _ptr(t0, c_dbl_p), _ptr(objfn, c_dbl_p))


def locate_l2_gridSearch__float64(ldgrd, ngrd, nobs, iwantOT, t0use, mask, tobs, tcorr, varobs, test, t0, objfn):
    """locate.c:1079-1203."""
    return _lib.load().locate_l2_gridSearch__float64(
        int(ldgrd), int(ngrd), int(nobs), int(iwantOT), float(t0use), _ptr(mask, c_int_p), _ptr(tobs, c_flt_p),
        _ptr(tcorr, c_flt_p), _ptr(varobs, c_flt_p), _ptr(test, c_flt_p), _ptr(t0, c_flt_p), _ptr(objfn, c_flt_p))


def locate_l1_gridSearch__double64(ldgrd, ngrd, nobs, iwantOT, t0use, mask, tobs, varobs, test, t0, objfn):
    """L1 flavour (locate.c:1205-1335): weighted-median origin time, weighted L1 misfit.  Returns the C return code."""
    return _lib.load().locate_l1_gridSearch__double64(
        int(ldgrd), int(ngrd), int(nobs), int(iwantOT), float(t0use), _ptr(mask, c_int_p), _ptr(tobs, c_dbl_p),
        _ptr(varobs, c_dbl_p), _ptr(test, c_dbl_p), _ptr(t0, c_dbl_p), _ptr(objfn, c_dbl_p))


def weightedMedian__double(x, w, perm=None):
    """Host helper with the prototype of locate.c:73 -> (median, lsort, ierr); perm (int32 array) is updated in place."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    lsort, ierr = C.c_bool(False), C.c_int(0)
    med = _lib.load().weightedMedian__double(x.size, _ptr(x, c_dbl_p), _ptr(w, c_dbl_p), _ptr(perm, c_int_p),
                                             C.byref(lsort), C.byref(ierr))
    return med, bool(lsort.value), ierr.value


def locate_minLocDouble64(n, x):
    return _lib.load().locate_minLocDouble64(int(n), _ptr(x, c_dbl_p))


def locate_minLocFloat64(n, x):
    return _lib.load().locate_minLocFloat64(int(n), _ptr(x, c_flt_p))


def locate3d_gridsearch__double64(ldgrd, ngrd, nobs, iwantOT, mask, tobs, varobs, test, logPDF):
    """gridsearch.f90:382-459 (by-reference Fortran binding).  Returns ierr."""
    ierr = C.c_int(0)
    _lib.load().locate3d_gridsearch__double64(_i(ldgrd), _i(ngrd), _i(nobs), _i(iwantOT), _ptr(mask, c_int_p),
                                              _ptr(tobs, c_dbl_p), _ptr(varobs, c_dbl_p), _ptr(test, c_dbl_p),
                                              _ptr(logPDF, c_dbl_p), C.byref(ierr))
    return ierr.value


def locate3d_gridsearch__float64(ldgrd, ngrd, nobs, iwantOT, mask, tobs, varobs, test, logPDF):
    """gridsearch.f90:463-540."""
    ierr = C.c_int(0)
    _lib.load().locate3d_gridsearch__float64(_i(ldgrd), _i(ngrd), _i(nobs), _i(iwantOT), _ptr(mask, c_int_p),
                                             _ptr(tobs, c_flt_p), _ptr(varobs, c_flt_p), _ptr(test, c_flt_p),
                                             _ptr(logPDF, c_flt_p), C.byref(ierr))
    return ierr.value


def locate3d_initialize(comm=0, iverb=0, tttFileID=0, locFileID=0, ndivx=1, ndivy=1, ndivz=1):
    """locate.f90:562-677 (HDF5 ids are recorded only; see :func:`locate3d_set_tables`)."""
    ierr = C.c_int(0)
    _lib.load().locate3d_initialize(_i(comm), _i(iverb), C.byref(C.c_long(tttFileID)), C.byref(C.c_long(locFileID)),
                                    _i(ndivx), _i(ndivy), _i(ndivz), C.byref(ierr))
    return ierr.value


def locate3d_set_tables(tables, ngrd, ldgrd=None):
    """Hand the fp32 tables [ntables, ldgrd] (what the reference reads from
    /TravelTimeTables/Model_m/Station_s/{P,S}TravelTimes) to the drop-in locator."""
    tables = np.ascontiguousarray(tables, dtype=np.float32)
    ldgrd = tables.shape[-1] if ldgrd is None else int(ldgrd)
    rc = _lib.load().mceik_locate3d_set_tables(tables.size // ldgrd, int(ngrd), ldgrd, _ptr(tables, c_flt_p))
    return _lib.check(rc, "mceik_locate3d_set_tables")


def locate3d_set_grid(xlocs, ylocs, zlocs):
    x, y, z = (np.ascontiguousarray(a, dtype=np.float32) for a in (xlocs, ylocs, zlocs))
    return _lib.check(_lib.load().mceik_locate3d_set_grid(x.size, _ptr(x, c_flt_p), _ptr(y, c_flt_p), _ptr(z, c_flt_p)),
                      "mceik_locate3d_set_grid")


def locate3d_gridsearch(model, job, nobs, nevents, luseObs, statPtr, pickType, statCor, tori, varobs, tobs, test, hypo):
    """locate.f90:322-519.  Rectangular [nevents x nobs] inputs; fills hypo[4*nevents]; returns ierr."""
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    a = [i32(luseObs), i32(statPtr), i32(pickType), f64(statCor), f64(tori), f64(varobs), f64(tobs)]
    ierr = C.c_int(0)
    _lib.load().locate3d_gridsearch(_i(model), _i(job), _i(nobs), _i(nevents), _ptr(a[0], c_int_p), _ptr(a[1], c_int_p),
                                    _ptr(a[2], c_int_p), _ptr(a[3], c_dbl_p), _ptr(a[4], c_dbl_p), _ptr(a[5], c_dbl_p),
                                    _ptr(a[6], c_dbl_p), _ptr(test, c_dbl_p), _ptr(hypo, c_dbl_p), C.byref(ierr))
    return ierr.value


def locate3d_finalize():
    _lib.load().locate3d_finalize()


class Locator:
    """Batched L2 grid search against travel-time tables resident in HBM."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.lib = _lib.load()
        self.ngrd = 0
        self._keep = None

    def set_tables_host(self, tables, ngrd, ldgrd=None):
        tables = np.ascontiguousarray(tables, dtype=np.float32)
        ldgrd = tables.shape[-1] if ldgrd is None else int(ldgrd)
        _lib.check(self.lib.mceik_locate_set_tables_host(self.ctx.handle, tables.size // ldgrd, int(ngrd), ldgrd,
                                                         _ptr(tables, c_flt_p)), "mceik_locate_set_tables_host")
        self.ngrd = int(ngrd)

    def set_tables_device(self, d_tables, ngrd):
        """d_tables: torch float32 CUDA tensor [ntables, ldgrd]; kept alive by this object."""
        assert d_tables.is_cuda and d_tables.is_contiguous() and d_tables.dim() == 2 and d_tables.dtype.itemsize == 4
        _lib.check(self.lib.mceik_locate_set_tables_dev(self.ctx.handle, d_tables.shape[0], int(ngrd), d_tables.shape[1],
                                                        C.c_void_p(d_tables.data_ptr())), "mceik_locate_set_tables_dev")
        self._keep = d_tables
        self.ngrd = int(ngrd)

    def set_grid(self, xlocs, ylocs, zlocs):
        x, y, z = (np.ascontiguousarray(a, dtype=np.float32) for a in (xlocs, ylocs, zlocs))
        _lib.check(self.lib.mceik_locate_set_grid(self.ctx.handle, x.size, _ptr(x, c_flt_p), _ptr(y, c_flt_p),
                                                  _ptr(z, c_flt_p)), "mceik_locate_set_grid")

    def locate_host(self, job, obs_ptr, table_id, tobs_cor, varobs, tori=None):
        """CSR picks on the host -> (iopt int32, t0opt, objopt) numpy arrays."""
        obs_ptr = np.ascontiguousarray(obs_ptr, dtype=np.int32)
        ne = obs_ptr.size - 1
        table_id = np.ascontiguousarray(table_id, dtype=np.int32)
        tobs_cor = np.ascontiguousarray(tobs_cor, dtype=np.float64)
        varobs = np.ascontiguousarray(varobs, dtype=np.float64)
        tori = None if tori is None else np.ascontiguousarray(tori, dtype=np.float64)
        iopt = np.zeros(max(ne, 1), dtype=np.int32)
        t0 = np.zeros(max(ne, 1), dtype=np.float64)
        obj = np.zeros(max(ne, 1), dtype=np.float64)
        rc = self.lib.mceik_locate_batched_host(self.ctx.handle, int(job), ne, _ptr(obs_ptr, c_int_p),
                                                _ptr(table_id, c_int_p), _ptr(tobs_cor, c_dbl_p), _ptr(varobs, c_dbl_p),
                                                _ptr(tori, c_dbl_p), _ptr(iopt, c_int_p), _ptr(t0, c_dbl_p),
                                                _ptr(obj, c_dbl_p))
        if rc != 0:
            raise _lib.MceikError(f"mceik_locate_batched_host rc={rc}: {_lib.last_error()}")
        return iopt[:ne], t0[:ne], obj[:ne]

    def event_logpdf(self, job, table_id, tobs_cor, varobs, tori=0.0, want_t0=True, want_f32=False):
        """Posterior volume of one event: (logPDF fp64 [ngrd], logPDF fp32 or None, t0 grid or None);
        logPDF = -objective (locate.f90:458-461)."""
        table_id = np.ascontiguousarray(table_id, dtype=np.int32)
        tobs_cor = np.ascontiguousarray(tobs_cor, dtype=np.float64)
        varobs = np.ascontiguousarray(varobs, dtype=np.float64)
        pdf = np.empty(self.ngrd, dtype=np.float64)
        pdf4 = np.empty(self.ngrd, dtype=np.float32) if want_f32 else None
        t0 = np.empty(self.ngrd, dtype=np.float64) if want_t0 else None
        rc = self.lib.mceik_locate_event_logpdf_host(self.ctx.handle, int(job), table_id.size, _ptr(table_id, c_int_p),
                                                     _ptr(tobs_cor, c_dbl_p), _ptr(varobs, c_dbl_p), float(tori),
                                                     _ptr(pdf, c_dbl_p), _ptr(pdf4, c_flt_p), _ptr(t0, c_dbl_p))
        if rc != 0:
            raise _lib.MceikError(f"mceik_locate_event_logpdf_host rc={rc}: {_lib.last_error()}")
        return pdf, pdf4, t0

    def optnode(self, pdf):
        """LOCATE_OPTNODE (locate.f90:75-117): 0-based first index of the maximum."""
        pdf = np.ascontiguousarray(pdf, dtype=np.float64)
        node = C.c_int(0)
        rc = self.lib.mceik_locate_optnode_host(self.ctx.handle, pdf.size, _ptr(pdf, c_dbl_p), C.byref(node))
        if rc != 0:
            raise _lib.MceikError(f"mceik_locate_optnode_host rc={rc}: {_lib.last_error()}")
        return node.value

    def normalize_pdf(self, pdf):
        """LOCATE_NORMALIZE_PDF (locate.f90:43-64): in place pdf /= sum(pdf); returns (ierr, sum)."""
        assert pdf.dtype == np.float64 and pdf.flags.c_contiguous
        xsum = C.c_double(0.0)
        rc = self.lib.mceik_locate_normalize_pdf_host(self.ctx.handle, pdf.size, _ptr(pdf, c_dbl_p), C.byref(xsum))
        if rc not in (0, 1):
            raise _lib.MceikError(f"mceik_locate_normalize_pdf_host rc={rc}: {_lib.last_error()}")
        return rc, xsum.value

    def locate_device(self, job, nevents, max_picks, d_obs_ptr, d_table_id, d_tobs_cor, d_varobs, d_tori, d_iopt,
                      d_t0opt, d_objopt):
        """All arguments are CUDA torch tensors (int32 / float64); asynchronous on the context stream."""
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        rc = self.lib.mceik_locate_batched_dev(self.ctx.handle, int(job), int(nevents), int(d_table_id.numel()),
                                               int(max_picks), p(d_obs_ptr), p(d_table_id), p(d_tobs_cor), p(d_varobs),
                                               p(d_tori), p(d_iopt), p(d_t0opt), p(d_objopt))
        if rc != 0:
            raise _lib.MceikError(f"mceik_locate_batched_dev rc={rc}: {_lib.last_error()}")

    def locate_catalog(self, catalog, stations, job):
        """catalog / stations: dicts of numpy arrays with the field names of mceik_struct.h.
        Returns (hypo [nevents,4], iopt, obj)."""
        keep = []

        def arr(d, k, dt):
            a = np.ascontiguousarray(d[k], dtype=dt)
            keep.append(a)
            return a.ctypes.data_as(c_dbl_p if dt == np.float64 else c_int_p)

        ne = int(catalog["nevents"])
        cs = CatalogStruct()
        for k in ("xsrc", "ysrc", "zsrc", "tori", "tobs", "test", "varObs"):
            if k in catalog:
                setattr(cs, k, arr(catalog, k, np.float64))
        for k in ("luseObs", "pickType", "statPtr", "obsPtr"):
            setattr(cs, k, arr(catalog, k, np.int32))
        cs.nevents = ne
        ss = StationsStruct()
        ss.nstat = int(stations["nstat"])
        ss.lcartesian = int(stations.get("lcartesian", 1))
        for k in ("xrec", "yrec", "zrec", "pcorr", "scorr"):
            if k in stations:
                setattr(ss, k, arr(stations, k, np.float64))
        hypo = np.zeros(4 * max(ne, 1), dtype=np.float64)
        iopt = np.zeros(max(ne, 1), dtype=np.int32)
        obj = np.zeros(max(ne, 1), dtype=np.float64)
        rc = self.lib.mceik_locate_catalog(self.ctx.handle, C.byref(cs), C.byref(ss), int(job), _ptr(hypo, c_dbl_p),
                                           _ptr(iopt, c_int_p), _ptr(obj, c_dbl_p))
        if rc != 0:
            raise _lib.MceikError(f"mceik_locate_catalog rc={rc}: {_lib.last_error()}")
        return hypo[:4 * ne].reshape(ne, 4), iopt[:ne], obj[:ne]


def catalog_misfit_device(ctx, d_tables, ngrd, nmodels, ntab, nevents, d_node, d_tobs, d_var, d_use, d_out):
    """``mceik_catalog_misfit_dev`` (BASELINE config 5): misfit of ``nmodels`` proposals -- model m owns rows
    [m*ntab, (m+1)*ntab) of the fp32 tables -- against a catalogue of ``nevents`` events at fixed nodes.  CUDA torch
    tensors: d_node int32 [ne], d_tobs / d_var float64 [ne, ntab], d_use int32 [ne, ntab], d_out float64 [nmodels]."""
    lib = _lib.load()
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.mceik_catalog_misfit_dev(ctx.handle, p(d_tables), int(d_tables.shape[-1]), int(ngrd), int(nmodels), int(ntab),
                                      int(nevents), p(d_node), p(d_tobs), p(d_var), p(d_use), p(d_out))
    if rc != 0:
        raise _lib.MceikError(f"mceik_catalog_misfit_dev rc={rc}: {_lib.last_error()}")
