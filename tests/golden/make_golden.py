"""Generates tests/golden/*.npz (run in the build container, where /root/reference exists).

* locate_ref_small.npz  -- inputs and OUTPUTS OF THE REFERENCE'S OWN locate.c (compiled unmodified
  into oracle/_ref/libref_locate.so by `make -C oracle ref`) for a small ragged case, fp64 and
  fp32, so the oracle restatement stays pinned to reference output on boxes without /root/reference.
* fsm_oracle_small.npz  -- a small eikonal field produced by the oracle restatement itself
  (regression guard only: the reference has no FSM golden vector, "parity unpinned").
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
import refcases  # noqa: E402


def main():
    assert O.ref() is not None, "needs oracle/_ref (the reference tree)"
    rng = np.random.default_rng(20161018)
    nx, ny, nz, nobs = 13, 11, 7, 6
    n = nx * ny * nz
    ld = n + 64 - n % 64
    test = O.aligned(nobs * ld, np.float64)
    for i in range(nobs):
        test[i * ld:i * ld + n] = refcases._table(nx, ny, nz, 700.0, 700.0, 700.0, *(rng.random(3) * 4000.0))
    tobs = rng.random(nobs) * 2 + 1
    tcorr = rng.normal(0, 0.05, nobs)
    var = rng.uniform(0.1, 0.9, nobs)
    mask = np.array([0, 1, 0, 0, 1, 0], np.int32)
    out = dict(nx=nx, ny=ny, nz=nz, nobs=nobs, ld=ld, test=np.array(test), tobs=tobs, tcorr=tcorr, var=var, mask=mask)
    rc, t0, obj = O.l2_gridsearch(ld, n, nobs, 1, 0.0, mask, tobs, tcorr, var, test, np.float64, use_ref=True)
    assert rc == 0
    out.update(t0_f64=np.array(t0), obj_f64=np.array(obj), iopt_f64=O.minloc(obj, use_ref=True))
    rc, t0, obj = O.l2_gridsearch(ld, n, nobs, 0, 1.75, mask, tobs, None, var, test, np.float64, use_ref=True)
    out.update(t0_fix=np.array(t0), obj_fix=np.array(obj))
    t4 = O.aligned(nobs * ld, np.float32)
    t4[:] = test.astype(np.float32)
    rc, t0, obj = O.l2_gridsearch(ld, n, nobs, 1, 0.0, mask, tobs.astype(np.float32), tcorr.astype(np.float32),
                                  var.astype(np.float32), t4, np.float32, use_ref=True)
    assert rc == 0
    out.update(t0_f32=np.array(t0), obj_f32=np.array(obj), iopt_f32=O.minloc(obj, use_ref=True))
    np.savez_compressed(os.path.join(HERE, "locate_ref_small.npz"), **out)

    nx, ny, nz, h = 21, 18, 15, 120.0
    slow = 1.0 / np.random.default_rng(5).uniform(2500.0, 6000.0, nx * ny * nz)
    u, ierr, it = O.eikonal_serial(nx, ny, nz, h, slow, [0.0, 0.1], [h * 7.3, h * 15.0], [h * 9.9, h * 3.0],
                                   [h * 6.2, h * 11.5], tol=1e-7, maxit=20)
    assert ierr == 0
    np.savez_compressed(os.path.join(HERE, "fsm_oracle_small.npz"), nx=nx, ny=ny, nz=nz, h=h, slow=slow, u=u, iters=it)
    print("golden vectors written")


if __name__ == "__main__":
    main()
