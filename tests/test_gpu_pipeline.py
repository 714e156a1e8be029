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


def test_c5_catalog_misfit_kernel(gpu_ctx):
    """mceik_catalog_misfit_dev (config 5: misfit of every proposal against a catalogue at fixed nodes) against a
    numpy statement of the same sums -- per event bit-equal arithmetic (locate.c:399-410, 500-513 order); the sum
    over events uses a fixed tree, so the total is compared to 1e-12 relative."""
    import torch
    from mceik_b200.locate import catalog_misfit_device
    rng = np.random.default_rng(31)
    nmod, ntab, ne, ngrd = 5, 7, 300, 4000
    tables = rng.uniform(0.5, 20.0, (nmod * ntab, ngrd)).astype(np.float32)
    node = rng.integers(0, ngrd, ne).astype(np.int32)
    tobs = rng.uniform(1.0, 25.0, (ne, ntab))
    var = rng.choice(np.array([0.1, 0.25, 0.5]), (ne, ntab))
    use = (rng.uniform(size=(ne, ntab)) >= 0.2).astype(np.int32)
    use[3] = 0  # an event without picks contributes nothing
    dev = lambda a: torch.from_numpy(a).cuda()
    out = torch.empty(nmod, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    catalog_misfit_device(gpu_ctx, dev(tables), ngrd, nmod, ntab, ne, dev(node), dev(tobs), dev(var), dev(use), out)
    gpu_ctx.synchronize()
    got = out.cpu().numpy()
    for m in range(nmod):
        tot = 0.0
        for e in range(ne):
            T = tables[m * ntab:(m + 1) * ntab, node[e]].astype(np.float64)
            xnorm, t0, obj = 0.0, 0.0, 0.0
            for j in range(ntab):
                if use[e, j]:
                    xnorm += 1.0 / var[e, j]
            for j in range(ntab):
                if use[e, j]:
                    t0 += ((1.0 / var[e, j]) / xnorm) * (tobs[e, j] - T[j])
            for j in range(ntab):
                if use[e, j]:
                    r = ((1.0 / var[e, j]) * 0.7071067811865475) * (tobs[e, j] - (T[j] + t0))
                    obj += r * r
            if xnorm > 0:
                tot += obj
        assert abs(got[m] - tot) <= 1e-12 * abs(tot), (m, got[m], tot)
