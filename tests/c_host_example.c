/*
 * tests/c_host_example.c -- a plain C host of libmceik_b200.so (no Python, no ctypes): the calls a maintainer's C
 * driver makes after the link-time substitution of INTEGRATION.md.  Test infrastructure: tests/test_abi.py checks that
 * it compiles and links against include/mceik_b200.h + the library; tests/test_gpu_c_host.py runs it on a B200 and
 * compares the printed values with the Python mirror of the same calls.
 *
 *   eikonal3d_serial_driver jobs 1, 2, 3 (fsm3d.f90:1968-2052 argument list, everything by reference)
 *   computeHomogeneousTraveltimes (homog.c:594-599), locate_l2_gridSearch__double64 + locate_minLocDouble64
 *   (locate.c:923-934, 811)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mceik_b200.h"

int main(void)
{
    const int nx = 24, ny = 16, nz = 20, nsrc = 1, maxit = 20, iverb = 0;
    const double h = 100.0, tol = 1e-6, x0 = 0.0, y0 = 0.0, z0 = 0.0;
    const double ts = 0.0, xs = 1130.0, ys = 770.0, zs = 810.0;
    const int n = nx * ny * nz;
    double *slow = (double *)malloc(sizeof(double) * n), *u = (double *)malloc(sizeof(double) * n);
    int ierr = 0, job, i;
    for (i = 0; i < n; i++) slow[i] = 1.0 / (3000.0 + 10.0 * (i % 97));
    job = 1; eikonal3d_serial_driver(&job, &iverb, &maxit, &nsrc, &nx, &ny, &nz, &tol, &h, &x0, &y0, &z0, &ts, &xs, &ys, &zs, slow, u, &ierr);
    if (ierr) { printf("init ierr=%d (%s)\n", ierr, mceik_last_error()); return 1; }
    job = 2; eikonal3d_serial_driver(&job, &iverb, &maxit, &nsrc, &nx, &ny, &nz, &tol, &h, &x0, &y0, &z0, &ts, &xs, &ys, &zs, slow, u, &ierr);
    if (ierr) { printf("solve ierr=%d (%s)\n", ierr, mceik_last_error()); return 1; }
    job = 3; eikonal3d_serial_driver(&job, &iverb, &maxit, &nsrc, &nx, &ny, &nz, &tol, &h, &x0, &y0, &z0, &ts, &xs, &ys, &zs, slow, u, &ierr);
    printf("u[0]=%.17g u[n/2]=%.17g u[n-1]=%.17g\n", u[0], u[n / 2], u[n - 1]);

    /* three analytic tables, one event sitting on node 1234 with origin time 2.5 */
    {
        const int nobs = 3, ngrd = n, ldgrd = (n + 7) / 8 * 8, node = 1234;
        const double sx[3] = {300.0, 1900.0, 1200.0}, sy[3] = {200.0, 1300.0, 700.0}, sz[3] = {1900.0, 1900.0, 100.0};
        double *test, *t0, *obj, tobs[3], var[3] = {0.25, 0.1, 0.5};
        int mask[3] = {0, 0, 0}, k, iopt;
        if (posix_memalign((void **)&test, 64, sizeof(double) * nobs * ldgrd) || posix_memalign((void **)&t0, 64, sizeof(double) * ldgrd) ||
            posix_memalign((void **)&obj, 64, sizeof(double) * ldgrd)) return 1;
        memset(test, 0, sizeof(double) * nobs * ldgrd);
        for (k = 0; k < nobs; k++) {
            if (computeHomogeneousTraveltimes(nx, ny, nz, x0, y0, z0, h, h, h, sx[k], sy[k], sz[k], 5000.0, test + (size_t)k * ldgrd)) return 1;
            tobs[k] = test[(size_t)k * ldgrd + node] + 2.5;
        }
        if (locate_l2_gridSearch__double64(ldgrd, ngrd, nobs, 1, 0.0, mask, tobs, NULL, var, test, t0, obj)) { printf("search failed\n"); return 1; }
        iopt = locate_minLocDouble64(ngrd, obj);
        printf("iopt=%d t0=%.17g obj=%.17g\n", iopt, t0[iopt], obj[iopt]);
        free(test); free(t0); free(obj);
    }
    free(slow); free(u);
    return 0;
}
