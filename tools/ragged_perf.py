"""Ragged catalogue timing: each event picks a random 70 % of the tables (sorted)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases, mceik_b200
from mceik_b200.locate import Locator
n, ns, ne = 128, 64, 256
h = 1000.0
rng = np.random.default_rng(5)
sx, sy, sz = rng.uniform(0, (n - 1) * h, ns), rng.uniform(0, (n - 1) * h, ns), np.full(ns, (n - 1) * h)
tables = np.concatenate([cases.homog_tables(n, n, n, h, sx, sy, sz, 5000.0), cases.homog_tables(n, n, n, h, sx, sy, sz, 2900.0)])
ntab, ngrd = tables.shape
ctx = mceik_b200.Context(0)
loc = Locator(ctx); loc.set_tables_host(tables, ngrd)
obs_ptr, tid, tc, var = [0], [], [], []
tori = rng.uniform(0, 5, ne)
for e in range(ne):
    ids = np.sort(rng.permutation(ntab)[:int(0.7 * ntab)])
    node = int(rng.integers(0, ngrd))
    tid += list(ids); tc += list(tables[ids, node] + tori[e]); var += [0.25] * len(ids)
    obs_ptr.append(len(tid))
obs_ptr, tid, tc, var = np.array(obs_ptr, np.int32), np.array(tid, np.int32), np.array(tc), np.array(var)
for mode in ("aligned", "general"):
    ctx.set_tuning("LOCATE_NO_ALIGN", 1 if mode == "general" else 0)
    for rep in range(3):
        t = time.time(); iopt, t0, obj = loc.locate_host(2, obs_ptr, tid, tc, var, tori); dt = time.time() - t
    print(mode, f"{ne/dt:.1f} events/s", int(iopt.sum()))
