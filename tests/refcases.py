"""The reference's own known-answer cases, rebuilt input for input.

* ``locate_c_main_case``: main() of locate.c:108-182 -- 145x145x45 grid (1 km), 20 receivers
  from srand(4042)/rand() of the C library, source node (12,15,5) -> flat index 107312, t0 = 4.
* ``gridsearch_f90_case``: the program in gridsearch.f90:29-87 -- 79x71x15, source node
  (31,55,4) 1-based -> 1-based index 21124, t0 = 4, unit variances.  Its receivers come from an
  unseeded random_number, so any receiver set must give the same answer; a fixed seed is used.
"""
import ctypes as C

import numpy as np

import oracle_lib as O

RAND_MAX = 2147483647


def _table(nx, ny, nz, dx, dy, dz, xr, yr, zr, slow=1.0 / 5000.0):
    X = (np.arange(nx) * dx)[None, None, :]
    Y = (np.arange(ny) * dy)[None, :, None]
    Z = (np.arange(nz) * dz)[:, None, None]
    ex, ey, ez = X - xr, Y - yr, Z - zr
    return (np.sqrt(ex * ex + ey * ey + ez * ez) * slow).ravel()


def locate_c_main_case():
    libc = C.CDLL("libc.so.6")
    libc.srand(4042)
    nobs, nx, ny, nz = 20, 145, 145, 45
    dx = dy = dz = 1.0e3
    ngrd = nx * ny * nz
    ldgrd = ngrd + 64 - ngrd % 64
    xsrc, ysrc, zsrc = 12 * dx, 15 * dy, 5 * dz
    test = O.aligned(nobs * ldgrd, np.float64)
    tobs = np.zeros(nobs)
    var = np.zeros(nobs)
    for io in range(nobs):
        xr = libc.rand() / RAND_MAX
        yr = libc.rand() / RAND_MAX
        zr = libc.rand() / RAND_MAX
        xr, yr, zr = xr * (nx - 1) * dx, yr * (ny - 1) * dy, zr * (nz - 1) * dz
        ex, ey, ez = xr - xsrc, yr - ysrc, zr - zsrc
        tobs[io] = np.sqrt(ex * ex + ey * ey + ez * ez) * (1.0 / 5000.0)
        var[io] = libc.rand() / RAND_MAX
        test[io * ldgrd: io * ldgrd + ngrd] = _table(nx, ny, nz, dx, dy, dz, xr, yr, zr)
    tobs = tobs + 4.0
    return dict(nobs=nobs, ngrd=ngrd, ldgrd=ldgrd, test=test, tobs=tobs, varobs=var, mask=np.zeros(nobs, np.int32),
                tcorr=np.zeros(nobs), t0use=4.0, true_index=5 * nx * ny + 15 * nx + 12)


def gridsearch_f90_case(seed=1992):
    rng = np.random.default_rng(seed)
    nobs, nx, ny, nz = 14, 79, 71, 15
    dx = dy = dz = 1.0e3
    ngrd = nx * ny * nz
    ldgrd = ngrd + 64 - ngrd % 64
    ixs, iys, izs = 31, 55, 4  # 1-based
    xsrc, ysrc, zsrc = (ixs - 1) * dx, (iys - 1) * dy, (izs - 1) * dz
    test = O.aligned(nobs * ldgrd, np.float64)
    tobs = np.zeros(nobs)
    for i in range(nobs):
        xr, yr, zr = rng.random() * (nx - 1) * dx, rng.random() * (ny - 1) * dy, rng.random() * (nz - 1) * dz
        test[i * ldgrd: i * ldgrd + ngrd] = _table(nx, ny, nz, dx, dy, dz, xr, yr, zr)
        ex, ey, ez = xr - xsrc, yr - ysrc, zr - zsrc
        tobs[i] = np.sqrt(ex * ex + ey * ey + ez * ez) * (1.0 / 5000.0)
    tobs = tobs + 4.0
    return dict(nobs=nobs, ngrd=ngrd, ldgrd=ldgrd, test=test, tobs=tobs, varobs=np.ones(nobs), mask=np.zeros(nobs, np.int32),
                true_index_1based=(izs - 1) * nx * ny + (iys - 1) * nx + ixs)
