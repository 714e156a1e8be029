"""Synthetic inputs shared by the tests, smoke() and bench.py."""
import numpy as np


def layered_slowness(nx, ny, nz, layer=16):
    """C2: 1-D layered P model, v(iz) = 6500 - 500*(iz div layer) m/s (fastest at the base)."""
    v = 6500.0 - 500.0 * (np.arange(nz) // layer)
    v = np.maximum(v, 1500.0)
    return np.repeat(1.0 / v, nx * ny)


def checkerboard_slowness(nx, ny, nz, cell=32, v0=5000.0, pert=0.1, vs=False):
    """C3: vp = v0*(1 + pert*s), s = (-1)^(ix div cell + iy div cell + iz div cell); vs = vp/sqrt(3)."""
    ix = np.arange(nx) // cell
    iy = np.arange(ny) // cell
    iz = np.arange(nz) // cell
    s = 1.0 - 2.0 * ((ix[None, None, :] + iy[None, :, None] + iz[:, None, None]) % 2)
    v = v0 * (1.0 + pert * s)
    if vs:
        v = v / np.sqrt(3.0)
    return (1.0 / v).ravel()


def random_slowness(n, seed, vmin=3000.0, vmax=5500.0):
    rng = np.random.default_rng(seed)
    return 1.0 / rng.uniform(vmin, vmax, n)


def interior_sources(n, nx, ny, nz, h, seed):
    """Random off-node interior positions x in [h, (n-2)h) (avoids the edge quirks of fsm3d.f90:716-755)."""
    rng = np.random.default_rng(seed)
    return (rng.uniform(h, (nx - 2) * h, n), rng.uniform(h, (ny - 2) * h, n), rng.uniform(h, (nz - 2) * h, n))


def node_coords(nx, ny, nz, dx, dy, dz, x0=0.0, y0=0.0, z0=0.0):
    """fp32 node coordinates in flat order (what /Model/{x,y,z}locs holds)."""
    x = (x0 + np.arange(nx) * dx).astype(np.float32)
    y = (y0 + np.arange(ny) * dy).astype(np.float32)
    z = (z0 + np.arange(nz) * dz).astype(np.float32)
    X = np.broadcast_to(x[None, None, :], (nz, ny, nx)).ravel()
    Y = np.broadcast_to(y[None, :, None], (nz, ny, nx)).ravel()
    Z = np.broadcast_to(z[:, None, None], (nz, ny, nx)).ravel()
    return X.copy(), Y.copy(), Z.copy()


def homog_tables(nx, ny, nz, h, sx, sy, sz, vp):
    """fp32 analytic tables [2*nstat, N]: table 2*s = P, 2*s+1 = S (vs = vp/sqrt(3)), computed like
    homog.c:605-619 in fp64 then rounded to fp32 (numpy evaluates the same IEEE operations)."""
    x = np.arange(nx) * h
    y = np.arange(ny) * h
    z = np.arange(nz) * h
    vs = vp / np.sqrt(3.0)
    out = np.empty((2 * len(sx), nx * ny * nz), dtype=np.float32)
    for s in range(len(sx)):
        ex = (sx[s] - x)[None, None, :]
        ey = (sy[s] - y)[None, :, None]
        ez = (sz[s] - z)[:, None, None]
        d = np.sqrt(ex * ex + ey * ey + ez * ez).ravel()
        out[2 * s] = (d * (1.0 / vp)).astype(np.float32)
        out[2 * s + 1] = (d * (1.0 / vs)).astype(np.float32)
    return out


def synthetic_catalog(tables, nevents, seed, mask_frac=0.1, variances=(0.1, 0.25, 0.5), tmax=10.0):
    """C1/C4-style rectangular catalogue: every event has one pick per table (station x phase);
    tobs = table value at a random true node + origin time; a fraction is masked (luseObs = 0)."""
    rng = np.random.default_rng(seed)
    ntab, ngrd = tables.shape
    true_node = rng.integers(0, ngrd, nevents)
    tori = rng.uniform(0.0, tmax, nevents)
    tobs = tables[:, true_node].T.astype(np.float64) + tori[:, None]
    var = rng.choice(np.asarray(variances), size=(nevents, ntab))
    use = (rng.uniform(size=(nevents, ntab)) >= mask_frac).astype(np.int32)
    stat = np.broadcast_to((np.arange(ntab) // 2 + 1)[None, :], (nevents, ntab)).astype(np.int32)
    ptype = np.broadcast_to((np.arange(ntab) % 2 + 1)[None, :], (nevents, ntab)).astype(np.int32)
    return dict(nobs=ntab, nevents=nevents, true_node=true_node, tori=tori, tobs=tobs.ravel(), varobs=var.ravel(),
                luseObs=use.ravel(), statPtr=stat.ravel().copy(), pickType=ptype.ravel().copy(),
                statCor=np.zeros(ntab))
