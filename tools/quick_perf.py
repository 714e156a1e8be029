"""Quick device-resident timing of the two hot kernels (development aid, not the bench)."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import mceik_b200  # noqa: E402
from mceik_b200 import _lib  # noqa: E402
from mceik_b200.eikonal import EikonalSolver  # noqa: E402
from mceik_b200.locate import Locator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--fields", type=int, default=16)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--gs-n", type=int, default=128)
ap.add_argument("--gs-stations", type=int, default=32)
ap.add_argument("--gs-events", type=int, default=512)
ap.add_argument("--skip-fsm", action="store_true")
ap.add_argument("--algo", type=int, default=2)  # 2 = brick kernels (the product default), 0 = tiles, 1 = levels
ap.add_argument("--skip-gs", action="store_true")
ap.add_argument("--maxit", type=int, default=20)
a = ap.parse_args()

torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = mceik_b200.Context(0, stream=stream.cuda_stream)

if not a.skip_fsm:
    n, h = a.n, 1000.0
    N = n ** 3
    slow = torch.from_numpy(cases.checkerboard_slowness(n, n, n, cell=max(n // 8, 1))).cuda()
    xs, ys, zs = cases.interior_sources(a.fields, n, n, n, h, seed=3)
    d_u = torch.empty((a.fields, N), dtype=torch.float64, device="cuda")
    sol = EikonalSolver(ctx, n, n, n, h, algo=a.algo, maxit=a.maxit)
    for rep in range(a.reps):
        torch.cuda.synchronize()
        t = time.time()
        iters, ferr = sol.solve_device(slow, np.zeros(a.fields, np.int32), np.zeros(a.fields), xs, ys, zs, d_u=d_u)
        torch.cuda.synchronize()
        dt = time.time() - t
        upd = sol.node_updates
        print(f"FSM {n}^3 x {a.fields} fields: iters={int(min(iters))}..{int(max(iters))} (mean {float(np.mean(iters)):.2f}) {dt*1e3:.1f} ms  {upd/dt/1e9:.2f} Gupd/s  "
              f"{24*upd/dt/1e9:.0f} GB/s algorithmic  sweep-kernel {sol.sweep_stats[0]:.1f} ms in {sol.sweep_stats[1]} launches", flush=True)
    del d_u

if not a.skip_gs:
    n, h = a.gs_n, 1000.0
    N = n ** 3
    ns = a.gs_stations
    rng = np.random.default_rng(5)
    sx, sy, sz = rng.uniform(0, (n - 1) * h, ns), rng.uniform(0, (n - 1) * h, ns), np.full(ns, (n - 1) * h)
    ntab = 2 * ns
    d_tab = torch.empty((ntab, N), dtype=torch.float32, device="cuda")
    lib = _lib.load()
    p = lambda x: x.ctypes.data_as(_lib.c_dbl_p)
    X = np.repeat(sx, 2); Y = np.repeat(sy, 2); Z = np.repeat(sz, 2)
    V = np.tile(np.array([5000.0, 5000.0 / np.sqrt(3.0)]), ns)
    rc = lib.mceik_homogeneous_tables_dev(ctx.handle, n, n, n, 0.0, 0.0, 0.0, h, h, h, ntab, p(X), p(Y), p(Z), p(V),
                                          C.c_void_p(d_tab.data_ptr()), N)
    assert rc == 0, _lib.last_error()
    ne = a.gs_events
    true_node = torch.from_numpy(rng.integers(0, N, ne)).cuda()
    tori = torch.from_numpy(rng.uniform(0, 10, ne)).cuda()
    tobs = (d_tab[:, true_node].T.double() + tori[:, None]).contiguous().view(-1)
    use = torch.from_numpy(rng.uniform(size=ne * ntab) >= 0.1).cuda()
    tid = torch.where(use, torch.arange(ntab, device="cuda").repeat(ne), torch.tensor(-1, device="cuda")).int()
    var = torch.from_numpy(rng.choice(np.array([0.1, 0.25, 0.5]), ne * ntab)).cuda()
    obs_ptr = (torch.arange(ne + 1, device="cuda") * ntab).int()
    iopt = torch.empty(ne, dtype=torch.int32, device="cuda")
    t0 = torch.empty(ne, dtype=torch.float64, device="cuda")
    obj = torch.empty(ne, dtype=torch.float64, device="cuda")
    loc = Locator(ctx)
    loc.set_tables_device(d_tab, N)
    for rep in range(a.reps):
        torch.cuda.synchronize()
        t = time.time()
        loc.locate_device(2, ne, ntab, obs_ptr, tid, tobs, var, None, iopt, t0, obj)
        torch.cuda.synchronize()
        dt = time.time() - t
        nuse = int(use.sum())
        print(f"GS {n}^3 x {ntab} tables x {ne} events: {dt*1e3:.1f} ms  {ne/dt:.1f} events/s  "
              f"{nuse*N*4/dt/1e9:.0f} GB/s algorithmic  {8*nuse*N/dt/1e12:.2f} TFLOP/s fp64  "
              f"hit={int((iopt.long() == true_node).sum())}/{ne}", flush=True)
