"""BASELINE config 5 at test scale: the MCMC forward loop -- perturbed velocity models x stations solved in
one batched launch, fp32 tables left in HBM, and the catalogue misfit of every proposal evaluated against
them without a host round trip.  Checked bit for bit against the CPU oracle (fsm3d + locate restatements)."""
import numpy as np
import pytest

import cases
import oracle_lib as O

pytestmark = pytest.mark.gpu


def test_c5_proposals_tables_and_misfit_stay_on_device(gpu_ctx):
    import torch
    from mceik_b200.eikonal import EikonalSolver
    from mceik_b200.locate import Locator
    nx, ny, nz, h = 40, 32, 24, 500.0
    n = nx * ny * nz
    nprop, nstat = 3, 2
    rng = np.random.default_rng(77)
    base = cases.checkerboard_slowness(nx, ny, nz, cell=8)
    # proposal m: smooth +-3 % perturbation of vp; S model = P model * sqrt(3)
    slow = []
    for m in range(nprop):
        pert = 1.0 + 0.03 * np.sin(np.arange(n) * (m + 1) * 1e-3) if m else np.ones(n)
        slow += [base * pert, base * pert * np.sqrt(3.0)]
    slow = np.stack(slow)                                   # [2*nprop, N]: (P, S) per proposal
    sx, sy, sz = cases.interior_sources(nstat, nx, ny, nz, h, seed=5)
    # fields of proposal m: station 0 P, station 0 S, station 1 P, station 1 S  (table id = 2*(station-1) + phase-1)
    fmodel = np.array([2 * m + ph for m in range(nprop) for s in range(nstat) for ph in (0, 1)], dtype=np.int32)
    fx = np.array([sx[s] for m in range(nprop) for s in range(nstat) for ph in (0, 1)])
    fy = np.array([sy[s] for m in range(nprop) for s in range(nstat) for ph in (0, 1)])
    fz = np.array([sz[s] for m in range(nprop) for s in range(nstat) for ph in (0, 1)])
    nf = fmodel.size
    ntab = 2 * nstat

    # ---- oracle: every field, then the catalogue (synthesised from proposal 0) against every proposal
    ref_tab = np.empty((nf, n), dtype=np.float32)
    ref_it = np.empty(nf, dtype=np.int32)
    for f in range(nf):
        u, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow[fmodel[f]], 0.0, fx[f], fy[f], fz[f])
        assert ierr == 0
        ref_tab[f], ref_it[f] = u.astype(np.float32), it
    cat = cases.synthetic_catalog(ref_tab[:ntab], 12, seed=9, mask_frac=0.1)
    cat["tobs"] = cat["tobs"] + rng.normal(0.0, 0.02, cat["tobs"].size)
    X, Y, Z = cases.node_coords(nx, ny, nz, h, h, h)

    # ---- device: one launch for all proposals, tables stay in HBM
    d_slow = torch.from_numpy(slow).cuda()
    d_tab = torch.empty((nf, n), dtype=torch.float32, device="cuda")
    sol = EikonalSolver(gpu_ctx, nx, ny, nz, h)
    torch.cuda.synchronize()                                # torch's stream -> the context's private stream
    iters, ferr = sol.solve_device(d_slow, fmodel, np.zeros(nf), fx, fy, fz, d_tables=d_tab)
    gpu_ctx.synchronize()
    assert not ferr.any() and np.array_equal(iters, ref_it)
    assert np.array_equal(d_tab.cpu().numpy(), ref_tab)

    ne, nobs = cat["nevents"], cat["nobs"]
    obs_ptr = np.arange(ne + 1, dtype=np.int32) * nobs
    tid = np.where(cat["luseObs"] == 1, 2 * (cat["statPtr"] - 1) + (cat["pickType"] - 1), -1).astype(np.int32)
    dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
    d_ptr, d_tid = dev(obs_ptr, np.int32), dev(tid, np.int32)
    d_tobs, d_var, d_tori = dev(cat["tobs"], np.float64), dev(cat["varobs"], np.float64), dev(cat["tori"], np.float64)
    d_iopt = torch.empty(ne, dtype=torch.int32, device="cuda")
    d_t0 = torch.empty(ne, dtype=torch.float64, device="cuda")
    d_obj = torch.empty(ne, dtype=torch.float64, device="cuda")
    loc = Locator(gpu_ctx)
    torch.cuda.synchronize()
    misfit, misfit_ref = [], []
    for m in range(nprop):
        loc.set_tables_device(d_tab[m * ntab:(m + 1) * ntab], n)
        loc.locate_device(2, ne, nobs, d_ptr, d_tid, d_tobs, d_var, d_tori, d_iopt, d_t0, d_obj)
        gpu_ctx.synchronize()
        rc, hypo_ref, iopt_ref, obj_ref = O.locate3d_catalog(2, n, n, ref_tab[m * ntab:(m + 1) * ntab], nobs, ne,
                                                             cat["luseObs"], cat["statPtr"], cat["pickType"], cat["statCor"],
                                                             cat["tori"], cat["varobs"], cat["tobs"], X, Y, Z)
        assert rc == 0
        assert np.array_equal(d_iopt.cpu().numpy(), iopt_ref)
        assert np.array_equal(d_obj.cpu().numpy(), obj_ref)
        assert np.array_equal(d_t0.cpu().numpy(), hypo_ref.reshape(ne, 4)[:, 3])
        misfit.append(float(np.sum(d_obj.cpu().numpy())))
        misfit_ref.append(float(np.sum(obj_ref)))
    assert misfit == misfit_ref
