"""A second, independent reading of the reference's serial eikonal path, in plain Python floats.  Test infrastructure.

`oracle/fsm3d_oracle.c` is the CPU oracle of this repository; no Fortran compiler exists in the build image or on the
GPU boxes, so it cannot be pinned against the compiled reference.  This file restates the same reference lines a second
time, in another language and without looking at the C text, so that `tests/test_oracle_golden.py` can check the two
readings against each other bit for bit (Python floats are IEEE doubles, every operation is rounded separately, and
`math.sqrt` is correctly rounded -- the arithmetic of the reference build, which has no FMA: Makefile.inc:4-13).

Followed lines (all in /root/reference/fsm3d.f90 unless noted):
  EIKONAL_SOURCE_INDEX :697-711, EIKONAL_INIT_GRID :716-755, EIKONAL3D_SETBCS :762-840,
  EIKONAL3D_FSM :28-99, EVAL_UPDATE3D :419-456, UPDATE3D :460-479, GET_U{X,Y,Z}MIN3D :483-541,
  SORT3 :557-608, SOLVE_HAMILTONIAN2D :618-632, SOLVE_HAMILTONIAN3D :644-693, the driver's job 2 :2007-2041,
  constants module.F90:2-10, u_nan module.F90:419.
Small grids only (pure Python loops).
"""
import math
import sys

U_NAN = sys.float_info.max          # HUGE(one)
THIRD = 1.0 / 3.0                   # one/three
TWO_THIRD = 2.0 / 3.0               # two/three
HALF = 1.0 / 2.0


def source_index(n, x0, dx, xs):
    if xs <= x0:
        return 1
    if xs >= x0 + float(n - 1) * dx:
        return n
    return int((xs - x0) / dx + HALF) + 1


def init_grid(n, isx, x0, dx, xs):
    """-> (ixloc[3] with -1 = unused, ierr)"""
    loc = [-1, -1, -1]
    est = x0 + float(isx - 1) * dx
    if est > xs:
        npinit = 2
        loc[0], loc[1] = isx - 1, isx
    elif est < xs:
        npinit = 2
        loc[0], loc[1] = isx, isx + 1
    else:
        npinit = 0
        if isx > 0:
            loc[npinit] = isx - 1
            npinit += 1
        loc[npinit] = isx
        npinit += 1
        if isx < n - 1:
            loc[npinit] = isx + 1
            npinit += 1
    ierr = 0
    for i in range(npinit):
        if loc[i] < 1 or loc[i] > n:
            ierr = 1
    return loc, ierr


def setbcs(nx, ny, nz, dx, dy, dz, x0, y0, z0, ts, xs, ys, zs, slow):
    """-> (u, lisbc, ierr); 1-based node (ix, iy, iz) lives at (iz-1)*nx*ny + (iy-1)*nx + ix - 1."""
    n = nx * ny * nz
    u = [U_NAN] * n
    lisbc = [False] * n
    nsrc = len(ts)
    isx = [source_index(nx, x0, dx, xs[s]) for s in range(nsrc)]
    isy = [source_index(ny, y0, dy, ys[s]) for s in range(nsrc)]
    isz = [source_index(nz, z0, dz, zs[s]) for s in range(nsrc)]
    for s in range(nsrc):
        ixloc, ierr = init_grid(nx, isx[s], x0, dx, xs[s])
        if ierr:
            return u, lisbc, ierr
        iyloc, ierr = init_grid(ny, isy[s], y0, dy, ys[s])
        if ierr:
            return u, lisbc, ierr
        izloc, ierr = init_grid(nz, isz[s], z0, dz, zs[s])
        if ierr:
            return u, lisbc, ierr
        for ix in ixloc:
            if ix == -1:
                continue
            for iy in iyloc:
                if iy == -1:
                    continue
                for iz in izloc:
                    if iz == -1:
                        continue
                    ijk = (iz - 1) * nx * ny + (iy - 1) * nx + ix - 1
                    x = x0 + float(ix - 1) * dx
                    y = y0 + float(iy - 1) * dy
                    z = z0 + float(iz - 1) * dz
                    ex, ey, ez = xs[s] - x, ys[s] - y, zs[s] - z
                    d = math.sqrt(ex * ex + ey * ey + ez * ez)
                    t = ts[s] + d * slow[ijk]
                    if abs(d) < 1.0e-10:
                        u[ijk] = t
                    else:
                        u[ijk] = min(u[ijk], t)
                    lisbc[ijk] = True
    return u, lisbc, 0


def sort3(a, b, c):
    lab, lac, lbc = not (a > b), not (a > c), not (b > c)
    if lab and lac:
        return (a, b, c) if lbc else (a, c, b)
    if (not lab) and lbc:
        return (b, a, c) if lac else (b, c, a)
    return (c, a, b) if lab else (c, b, a)


def hamiltonian2d(a, b, f):
    amb = a - b
    if abs(amb) < f:
        arg = 2.0 * f * f - amb * amb
        return HALF * (a + b + math.sqrt(arg))
    return min(a, b) + f


def hamiltonian3d(a, b, c, f):
    a1, a2, a3 = sort3(a, b, c)
    if a1 == U_NAN:
        return U_NAN
    x = a1 + f
    if x > a2:
        x = hamiltonian2d(a1, a2, f)
        if x > a3:
            qb = -TWO_THIRD * (a1 + a2 + a3)
            qc = (a1 * a1 + a2 * a2 + a3 * a3 - f * f) * THIRD
            disc = qb * qb - 4.0 * qc
            x = HALF * (-qb + math.sqrt(disc)) if disc >= 0.0 else float("nan")
            if x < U_NAN:
                return x
            return U_NAN                     # NaN or overflow: the function result keeps its initial u_nan
        return x
    return x


def fsm(nx, ny, nz, h, tol, maxit, slow, u, lisbc):
    """EIKONAL3D_FSM on u in place -> iterations run.  Levels ix + iy + iz = const in increasing order; the order of
    the nodes inside a level does not matter (no node of a level is a neighbour of another)."""
    assert nx >= 2 and ny >= 2 and nz >= 2
    nxy = nx * ny
    n = nxy * nz
    sweeps = [(False, False, False), (True, False, False), (False, True, False), (True, True, False),
              (False, False, True), (True, False, True), (False, True, True), (True, True, True)]
    levels = [[] for _ in range(nx + ny + nz + 1)]
    for iz in range(1, nz + 1):
        for iy in range(1, ny + 1):
            for ix in range(1, nx + 1):
                levels[ix + iy + iz].append((ix, iy, iz))
    u0 = list(u)
    it = 0
    for k in range(1, maxit + 1):
        it = k
        for rx, ry, rz in sweeps:
            for lev in levels:
                for ix, iy, iz in lev:
                    if rx:
                        ix = nx + 1 - ix
                    if ry:
                        iy = ny + 1 - iy
                    if rz:
                        iz = nz + 1 - iz
                    ijk = (iz - 1) * nxy + (iy - 1) * nx + ix - 1
                    if lisbc[ijk]:
                        continue
                    f = slow[ijk] * h
                    if 1 < ix < nx:
                        ux = min(u[ijk - 1], u[ijk + 1])
                    elif ix == 1:
                        ux = min(u[ijk], u[ijk + 1])
                    else:
                        ux = min(u[ijk - 1], u[ijk])
                    if 1 < iy < ny:
                        uy = min(u[ijk - nx], u[ijk + nx])
                    elif iy == 1:
                        uy = min(u[ijk], u[ijk + nx])
                    else:
                        uy = min(u[ijk - nx], u[ijk])
                    if 1 < iz < nz:
                        uz = min(u[ijk - nxy], u[ijk + nxy])
                    elif iz == 1:
                        uz = min(u[ijk], u[ijk + nxy])
                    else:
                        uz = min(u[ijk - nxy], u[ijk])
                    u[ijk] = min(u[ijk], hamiltonian3d(ux, uy, uz, f))
        lconv = 0
        for i in range(n):
            if abs(u0[i] - u[i]) < tol:
                lconv += 1
            u0[i] = u[i]
        if lconv == n:
            break
    return it


def serial_driver(nx, ny, nz, h, slow, ts, xs, ys, zs, tol=1e-6, maxit=20, x0=0.0, y0=0.0, z0=0.0):
    """Job 2 of eikonal3d_serial_driver -> (u, ierr, iterations)."""
    u, lisbc, ierr = setbcs(nx, ny, nz, h, h, h, x0, y0, z0, ts, xs, ys, zs, slow)
    if ierr:
        return u, ierr, 0
    it = fsm(nx, ny, nz, h, tol, maxit, slow, u, lisbc)
    return u, 0, it
