"""CPU study (not a test): how many brick updates of a solve are provably idempotent?

Runs the oracle's Godunov update in lexicographic Gauss-Seidel order on the bench's checkerboard model and
prints, per sweep, the changed nodes / bricks and the bricks a sound skipping rule would leave out
(DESIGN.md section 8).  Usage: python tests/studies/skipstudy.py <n> <brick z extent> [S]
"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import oracle_lib as O  # noqa: E402

O.build()
so = os.path.join(HERE, "_skipstudy.so")
subprocess.check_call(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                       os.path.join(HERE, "skipstudy.c"), "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                       "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lm"])
n, bz = int(sys.argv[1]), int(sys.argv[2])
vs = len(sys.argv) > 3 and sys.argv[3] == "S"
L = C.CDLL(so)
slow = cases.checkerboard_slowness(n, n, n, vs=vs)
xs, ys, zs = cases.interior_sources(64, n, n, n, 100.0, seed=3)
L.study(C.c_int(n), C.c_double(100.0), slow.ctypes.data_as(C.c_void_p), C.c_double(xs[0]), C.c_double(ys[0]),
        C.c_double(zs[0]), C.c_double(1e-6), 20, bz)
