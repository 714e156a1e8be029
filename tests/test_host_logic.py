"""CPU tests of the host-side logic of the product: the boundary-condition stencil records that the
C-ABI layer ships to the device (host_logic.hpp) against the oracle's EIKONAL3D_SETBCS, and the
tile execution order of the CUDA kernel against the global hyperplane order (emulated on the CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = tmp_path_factory.mktemp("shim") / "libshim.so"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                           "-I", os.path.join(ROOT, "mceik_b200", "csrc"), "-I", "/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "host_logic_shim.cpp"), "-o", str(so)])
    return C.CDLL(str(so))


def _records(shim, nx, ny, nz, h, ts, xs, ys, zs, org=(0.0, 0.0, 0.0)):
    ts, xs, ys, zs = (np.ascontiguousarray(a, np.float64) for a in (ts, xs, ys, zs))
    cap = 27 * len(ts)
    node, col = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    d, t = np.zeros(cap), np.zeros(cap)
    p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
    n = shim.shim_bc_records(nx, ny, nz, C.c_double(h), C.c_double(org[0]), C.c_double(org[1]), C.c_double(org[2]), len(ts),
                             p(ts, C.c_double), p(xs, C.c_double), p(ys, C.c_double), p(zs, C.c_double), cap,
                             p(node, C.c_int), p(d, C.c_double), p(t, C.c_double), p(col, C.c_int))
    return n, node[:max(n, 0)], d[:max(n, 0)], t[:max(n, 0)], col[:max(n, 0)]


def _oracle_bcs(nx, ny, nz, h, ts, xs, ys, zs, slow, org=(0.0, 0.0, 0.0)):
    n = nx * ny * nz
    lisbc = np.zeros(n, np.uint8)
    u = np.zeros(n)
    a = [np.ascontiguousarray(v, np.float64) for v in (ts, xs, ys, zs, slow)]
    p = lambda v: v.ctypes.data_as(O.c_dbl_p)
    ierr = O.lib().oracle_setbcs(nx, ny, nz, len(a[0]), C.c_double(h), C.c_double(h), C.c_double(h), C.c_double(org[0]),
                                 C.c_double(org[1]), C.c_double(org[2]), p(a[0]), p(a[1]), p(a[2]), p(a[3]), p(a[4]),
                                 lisbc.ctypes.data_as(C.POINTER(C.c_ubyte)), p(u))
    return ierr, lisbc, u


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_bc_records_reproduce_setbcs(shim, seed):
    rng = np.random.default_rng(seed)
    nx, ny, nz, h = 19, 23, 17, 75.0
    org = (100.0, -50.0, 12.5)
    slow = 1.0 / rng.uniform(2000, 6000, nx * ny * nz)
    ns = 3
    xs = org[0] + rng.uniform(h, (nx - 2) * h, ns)
    ys = org[1] + rng.uniform(h, (ny - 2) * h, ns)
    zs = org[2] + rng.uniform(h, (nz - 2) * h, ns)
    xs[1] = org[0] + 5 * h          # exactly on a node in x
    ys[2], zs[2] = ys[0], zs[0]     # overlapping stencils
    xs[2] = xs[0] + 0.25 * h
    ts = rng.uniform(0, 1, ns)
    n, node, d, t, col = _records(shim, nx, ny, nz, h, ts, xs, ys, zs, org)
    ierr, lisbc, u = _oracle_bcs(nx, ny, nz, h, ts, xs, ys, zs, slow, org)
    assert ierr == 0 and n > 0
    # replay the records the way apply_bcs_kernel does (sequential, min or assign)
    ur = np.full(nx * ny * nz, np.finfo(np.float64).max)
    for i in range(n):
        v = t[i] + d[i] * slow[node[i]]
        ur[node[i]] = v if col[i] else min(ur[node[i]], v)
    assert np.array_equal(ur, u)
    assert np.array_equal(np.unique(node), np.flatnonzero(lisbc))


def test_bc_records_error_cases(shim):
    nx = ny = nz = 10
    assert _records(shim, nx, ny, nz, 10.0, [0.0], [0.0], [45.0], [45.0])[0] == -1     # on node 1
    assert _records(shim, nx, ny, nz, 10.0, [0.0], [-4.0], [45.0], [45.0])[0] == -1    # left of the grid
    assert _records(shim, nx, ny, nz, 10.0, [0.0], [200.0], [45.0], [45.0])[0] == -1   # right of the grid
    n, node, d, t, col = _records(shim, nx, ny, nz, 10.0, [0.5], [80.0], [40.0], [40.0])  # node nx-1, on nodes in y,z
    assert n == 2 * 3 * 3 and col.sum() == 1 and d[col == 1][0] == 0.0
    assert _oracle_bcs(nx, ny, nz, 10.0, [0.0], [0.0], [45.0], [45.0], np.ones(1000))[0] == 1


@pytest.mark.parametrize("shape", [(37, 50, 21), (16, 16, 16), (33, 17, 48), (5, 70, 3), (64, 48, 32)])
def test_tile_order_equals_hyperplane_order(tmp_path, shape):
    """tests/tile_order_emulation.c: 16^3 tiles + clamped halo + tile-hyperplane order == global order, bit for bit."""
    O.build()
    exe = tmp_path / "tile_emu"
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-o", str(exe),
                           os.path.join(ROOT, "tests", "tile_order_emulation.c"), "-L", O.ORACLE_DIR, "-loracle", "-lm",
                           f"-Wl,-rpath,{O.ORACLE_DIR}"])
    out = subprocess.run([str(exe)] + [str(s) for s in shape], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("MATCH"), out.stdout
