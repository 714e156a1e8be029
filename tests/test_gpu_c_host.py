"""A plain C host (tests/c_host_example.c, compiled with gcc against include/mceik_b200.h and linked to the library)
makes the drop-in calls of INTEGRATION.md on a B200; its printed values must equal what the oracle computes for the
same inputs, bit for bit (%.17g round-trips a double)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_c_host(tmpdir):
    exe = os.path.join(str(tmpdir), "c_host_example")
    libdir = os.path.join(ROOT, "mceik_b200", "lib")
    subprocess.check_call(["/usr/bin/gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "c_host_example.c"), "-L", libdir, "-lmceik_b200", f"-Wl,-rpath,{libdir}"])
    return exe


@pytest.mark.gpu
def test_c_host_drop_in_calls(tmp_path):
    exe = build_c_host(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [l for l in out.stdout.splitlines() if "=" in l]
    vals = {}
    for l in lines:
        for tok in l.split():
            k, v = tok.split("=")
            vals[k] = float(v)
    nx, ny, nz, h = 24, 16, 20, 100.0
    n = nx * ny * nz
    slow = 1.0 / (3000.0 + 10.0 * (np.arange(n) % 97))
    u, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow, 0.0, 1130.0, 770.0, 810.0)
    assert ierr == 0
    assert vals["u[0]"] == u[0] and vals["u[n/2]"] == u[n // 2] and vals["u[n-1]"] == u[n - 1]
    ld = (n + 7) // 8 * 8
    test = O.aligned(3 * ld, np.float64)
    sx, sy, sz = [300.0, 1900.0, 1200.0], [200.0, 1300.0, 700.0], [1900.0, 1900.0, 100.0]
    for k in range(3):
        test[k * ld:k * ld + n] = O.homogeneous_traveltimes(nx, ny, nz, 0.0, 0.0, 0.0, h, h, h, sx[k], sy[k], sz[k], 5000.0)
    tobs = np.array([test[k * ld + 1234] + 2.5 for k in range(3)])
    rc, t0, obj = O.l2_gridsearch(ld, n, 3, 1, 0.0, np.zeros(3, np.int32), tobs, None, np.array([0.25, 0.1, 0.5]), test,
                                  np.float64, use_ref=O.ref() is not None)
    assert rc == 0
    iopt = O.minloc(obj[:n], use_ref=O.ref() is not None)
    assert int(vals["iopt"]) == iopt == 1234
    assert vals["t0"] == t0[iopt] and vals["obj"] == obj[iopt]
