"""ctypes binding of libmceik_b200.so (the C ABI declared in include/mceik_b200.h).

The shared library is the product; this module only loads it and declares signatures.  There is
no Python or CPU fallback: a missing library or a missing B200 raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCEIK_B200_LIB") or os.path.join(_HERE, "lib", "libmceik_b200.so")

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_flt_p = C.POINTER(C.c_float)
c_long_p = C.POINTER(C.c_long)


class MceikError(RuntimeError):
    pass


class FsmGrid(C.Structure):
    """mceik_fsm_grid"""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("h", C.c_double),
                ("x0", C.c_double), ("y0", C.c_double), ("z0", C.c_double),
                ("tol", C.c_double), ("maxit", C.c_int)]


class CatalogStruct(C.Structure):
    """struct mceik_catalog_struct (reference include/mceik_struct.h:10-32)"""
    _fields_ = [("xsrc", c_dbl_p), ("ysrc", c_dbl_p), ("zsrc", c_dbl_p), ("tori", c_dbl_p),
                ("tobs", c_dbl_p), ("test", c_dbl_p), ("varObs", c_dbl_p), ("luseObs", c_int_p),
                ("pickType", c_int_p), ("statPtr", c_int_p), ("obsPtr", c_int_p), ("nevents", C.c_int)]


class StationsStruct(C.Structure):
    """struct mceik_stations_struct (reference include/mceik_struct.h:34-49)"""
    _fields_ = [("netw", C.POINTER(C.c_char_p)), ("stnm", C.POINTER(C.c_char_p)),
                ("chan", C.POINTER(C.c_char_p)), ("loc", C.POINTER(C.c_char_p)),
                ("xrec", c_dbl_p), ("yrec", c_dbl_p), ("zrec", c_dbl_p), ("pcorr", c_dbl_p),
                ("scorr", c_dbl_p), ("lhasP", c_int_p), ("lhasS", c_int_p), ("nstat", C.c_int),
                ("lcartesian", C.c_int)]


# symbol -> (restype, argtypes); every symbol include/mceik_b200.h declares is listed here and
# tests/test_abi.py checks the two stay in step.
SIGNATURES = {
    # drop-in
    "eikonal3d_serial_driver": (None, [c_int_p] * 7 + [c_dbl_p] * 5 + [c_dbl_p] * 4 + [c_dbl_p, c_dbl_p, c_int_p]),
    "eikonal3d_initialize": (None, [c_int_p] * 10 + [c_dbl_p] * 5 + [c_int_p]),
    "eikonal3d_solve": (None, [c_int_p] * 3 + [c_dbl_p] * 6 + [c_int_p]),
    "eikonal3d_finalize": (None, [c_int_p, c_int_p]),
    "locate_l2_gridSearch__double64": (C.c_int, [C.c_int] * 4 + [C.c_double, c_int_p] + [c_dbl_p] * 6),
    "locate_l2_gridSearch__float64": (C.c_int, [C.c_int] * 4 + [C.c_float, c_int_p] + [c_flt_p] * 6),
    "locate_l1_gridSearch__double64": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, c_int_p, c_dbl_p, c_dbl_p,
                                                 c_dbl_p, c_dbl_p, c_dbl_p]),
    "weightedMedian__double": (C.c_double, [C.c_int, c_dbl_p, c_dbl_p, c_int_p, C.POINTER(C.c_bool), c_int_p]),
    "locate_minLocDouble64": (C.c_int, [C.c_int, c_dbl_p]),
    "locate_minLocFloat64": (C.c_int, [C.c_int, c_flt_p]),
    "locate3d_gridsearch__double64": (None, [c_int_p] * 5 + [c_dbl_p] * 4 + [c_int_p]),
    "locate3d_gridsearch__float64": (None, [c_int_p] * 5 + [c_flt_p] * 4 + [c_int_p]),
    "locate3d_initialize": (None, [c_int_p, c_int_p, c_long_p, c_long_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "locate3d_gridsearch": (None, [c_int_p] * 7 + [c_dbl_p] * 6 + [c_int_p]),
    "locate3d_finalize": (None, []),
    "computeHomogeneousTraveltimes": (C.c_int, [C.c_int] * 3 + [C.c_double] * 10 + [c_dbl_p]),
    # batched extensions
    "mceik_last_error": (C.c_char_p, []),
    "mceik_kernel_launch_count": (C.c_longlong, []),
    "mceik_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mceik_ctx_destroy": (None, [C.c_void_p]),
    "mceik_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "mceik_fsm_solve_batched_host": (C.c_int, [C.c_void_p, C.POINTER(FsmGrid), C.c_int, c_dbl_p, C.c_int, c_int_p, c_int_p,
                                               c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_flt_p, C.c_size_t, c_int_p, c_int_p]),
    "mceik_fsm_solve_batched_dev": (C.c_int, [C.c_void_p, C.POINTER(FsmGrid), C.c_int, C.c_void_p, C.c_int, c_int_p, c_int_p,
                                              c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_void_p, C.c_void_p, C.c_size_t, c_int_p, c_int_p]),
    "mceik_catalog_misfit_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mceik_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mceik_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mceik_comm_destroy": (C.c_int, [C.c_void_p]),
    "mceik_fsm_assign_fields": (C.c_int, [C.c_int, c_int_p, c_int_p, C.c_int, c_int_p, c_int_p, c_int_p]),
    "mceik_fsm_solve_sharded_dev": (C.c_int, [C.c_void_p, C.POINTER(FsmGrid), C.c_int, C.c_void_p, C.c_int, c_int_p, c_int_p,
                                              c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_int_p, C.c_void_p, C.c_size_t, c_int_p, c_int_p, c_int_p]),
    "mceik_tables_allgather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "mceik_tables_alloc_replicated": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p)]),
    "mceik_tables_free_replicated": (C.c_int, [C.c_void_p]),
    "mceik_fsm_set_algo": (C.c_int, [C.c_void_p, C.c_int]),
    "mceik_fsm_set_tuning": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "mceik_fsm_last_node_updates": (C.c_longlong, [C.c_void_p]),
    "mceik_fsm_last_sweep_stats": (C.c_int, [C.c_void_p, c_dbl_p, c_int_p]),
    "mceik_selftest_solver": (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_longlong, C.POINTER(C.c_longlong),
                                        C.POINTER(C.c_longlong)]),
    "mceik_homogeneous_tables_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_double] * 6 +
                                     [C.c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_void_p, C.c_size_t]),
    "mceik_locate_set_tables_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, c_flt_p]),
    "mceik_locate_set_tables_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p]),
    "mceik_locate_set_grid": (C.c_int, [C.c_void_p, C.c_int, c_flt_p, c_flt_p, c_flt_p]),
    "mceik_locate3d_set_tables": (C.c_int, [C.c_int, C.c_int, C.c_size_t, c_flt_p]),
    "mceik_locate3d_set_grid": (C.c_int, [C.c_int, c_flt_p, c_flt_p, c_flt_p]),
    "mceik_locate_batched_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p,
                                            c_int_p, c_dbl_p, c_dbl_p]),
    "mceik_locate_batched_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8),
    "mceik_locate_event_logpdf_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int_p, c_dbl_p, c_dbl_p, C.c_double,
                                                 c_dbl_p, c_flt_p, c_dbl_p]),
    "mceik_locate_optnode_host": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, c_int_p]),
    "mceik_locate_normalize_pdf_host": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, c_dbl_p]),
    "mceik_locate_catalog": (C.c_int, [C.c_void_p, C.POINTER(CatalogStruct), C.POINTER(StationsStruct), C.c_int,
                                       c_dbl_p, c_int_p, c_dbl_p]),
}

_lib = None


def load():
    """Load libmceik_b200.so (built by __graft_entry__.build() / make -C mceik_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MceikError(
            f"{LIB_PATH} is missing: build it with `make -C mceik_b200/csrc` (or __graft_entry__.build()). "
            "mceik_b200 has no Python/CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library diverged
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().mceik_last_error().decode()


def check(rc, what):
    """Raise on a negative (argument/CUDA) return code; pass 0 and positive status through."""
    if rc < 0:
        raise MceikError(f"{what} failed (rc={rc}): {last_error()}")
    return rc


def kernel_launch_count():
    return int(load().mceik_kernel_launch_count())
