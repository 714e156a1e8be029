// scratch: how many (brick, sweep) tasks are no-ops in later iterations?
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
double oracle_hamiltonian3d(double a, double b, double c, double f, int *ierr);
int oracle_setbcs(int nx, int ny, int nz, int nsrc, double dx, double dy, double dz, double x0, double y0, double z0,
                  const double *ts, const double *xs, const double *ys, const double *zs, const double *slow,
                  unsigned char *lisbc, double *u);
static int upd(int nx, int ny, int nz, int ix, int iy, int iz, double h, const double *slow, double *u) {
    long nxy = (long)nx * ny, ijk = (long)iz * nxy + (long)iy * nx + ix;
    double f = slow[ijk] * h, um, up, ux, uy, uz;
    um = ix > 0 ? u[ijk - 1] : u[ijk]; up = ix < nx - 1 ? u[ijk + 1] : u[ijk]; ux = um < up ? um : up;
    um = iy > 0 ? u[ijk - nx] : u[ijk]; up = iy < ny - 1 ? u[ijk + nx] : u[ijk]; uy = um < up ? um : up;
    um = iz > 0 ? u[ijk - nxy] : u[ijk]; up = iz < nz - 1 ? u[ijk + nxy] : u[ijk]; uz = um < up ? um : up;
    int ierr; double ub = oracle_hamiltonian3d(ux, uy, uz, f, &ierr);
    if (ub < u[ijk]) { u[ijk] = ub; return 1; }
    return 0;
}
// bz = brick z extent
int study(int n, double h, const double *slow, double xs, double ys, double zs, double tol, int maxit, int bz) {
    long N = (long)n * n * n;
    double *u = malloc(8 * N), *u0 = malloc(8 * N); unsigned char *bc = malloc(N);
    double ts = 0; oracle_setbcs(n, n, n, 1, h, h, h, 0, 0, 0, &ts, &xs, &ys, &zs, slow, bc, u);
    int nb = n / 8, nbz = (n + bz - 1) / bz; long NB = (long)nb * nb * nbz;
    // version stamp of last change per brick, and stamp at which each brick was last processed
    long *chg = calloc(NB, sizeof(long)), *seen = calloc(NB, sizeof(long)); unsigned char *c = malloc(NB);
    long stamp = 0;
    for (int it = 1; it <= maxit; it++) {
        memcpy(u0, u, 8 * N);
        for (int s = 0; s < 8; s++) {
            int rx = s & 1, ry = (s >> 1) & 1, rz = (s >> 2) & 1;
            memset(c, 0, NB); long nch = 0;
            for (int kz = 0; kz < n; kz++) { int iz = rz ? n - 1 - kz : kz;
              for (int ky = 0; ky < n; ky++) { int iy = ry ? n - 1 - ky : ky;
                for (int kx = 0; kx < n; kx++) { int ix = rx ? n - 1 - kx : kx;
                  long ijk = ((long)iz * n + iy) * n + ix;
                  if (bc[ijk]) continue;
                  if (upd(n, n, n, ix, iy, iz, h, slow, u)) { nch++; c[((long)(iz / bz) * nb + iy / 8) * nb + ix / 8] = 1; }
                } } }
            // skippable = brick unchanged this sweep and no neighbour (incl. itself) changed since its previous processing.
            // approximate accounting at sweep granularity: brick is "needed" in sweep S if it or a face neighbour changed in S-1 or S
            ++stamp;
            long nchb = 0, need = 0;
            for (long b = 0; b < NB; b++) if (c[b]) { chg[b] = stamp; nchb++; }
            for (int K = 0; K < nbz; K++) for (int J = 0; J < nb; J++) for (int I = 0; I < nb; I++) {
                long b = ((long)K * nb + J) * nb + I; long last = chg[b];
                if (I > 0 && chg[b - 1] > last) last = chg[b - 1];
                if (I < nb - 1 && chg[b + 1] > last) last = chg[b + 1];
                if (J > 0 && chg[b - nb] > last) last = chg[b - nb];
                if (J < nb - 1 && chg[b + nb] > last) last = chg[b + nb];
                if (K > 0 && chg[b - (long)nb * nb] > last) last = chg[b - (long)nb * nb];
                if (K < nbz - 1 && chg[b + (long)nb * nb] > last) last = chg[b + (long)nb * nb];
                if (last >= stamp - 1) need++;   // changed in this sweep or the one before
            }
            // rule R: skip(b,S) iff b and its 4 xy-neighbours made no change in S-1, and the upwind x/y neighbours made no change in S
            static unsigned char *cprev = NULL; if (!cprev) { cprev = calloc(NB, 1); memset(cprev, 1, NB); }
            long nskip = 0, unsound = 0;
            for (int K = 0; K < nbz; K++) for (int J = 0; J < nb; J++) for (int I = 0; I < nb; I++) {
                long b = ((long)K * nb + J) * nb + I;
                int dirty = cprev[b];
                if (I > 0) dirty |= cprev[b - 1]; if (I < nb - 1) dirty |= cprev[b + 1];
                if (J > 0) dirty |= cprev[b - nb]; if (J < nb - 1) dirty |= cprev[b + nb];
                if (K > 0) dirty |= cprev[b - (long)nb * nb]; if (K < nbz - 1) dirty |= cprev[b + (long)nb * nb];
                int ux = rx ? I + 1 : I - 1, uy = ry ? J + 1 : J - 1, uz = rz ? K + 1 : K - 1;
                if (ux >= 0 && ux < nb) dirty |= c[((long)K * nb + J) * nb + ux];
                if (uy >= 0 && uy < nb) dirty |= c[((long)K * nb + uy) * nb + I];
                if (uz >= 0 && uz < nbz) dirty |= c[((long)uz * nb + J) * nb + I];
                if (!dirty) { nskip++; if (c[b]) unsound++; }
            }
            memcpy(cprev, c, NB);
            printf("   rule R: skip %ld of %ld (%.3f) unsound %ld\n", nskip, NB, (double)nskip / NB, unsound);
            printf("it %2d sweep %d: changed nodes %9ld (%.4f)  changed bricks %6ld / %ld  needed(<=1 sweep old) %6ld (%.3f)\n", it, s, nch,
                   (double)nch / N, nchb, NB, need, (double)need / NB);
        }
        long lconv = 0; for (long i = 0; i < N; i++) if (fabs(u0[i] - u[i]) < tol) lconv++;
        printf("== it %d nonconv %ld\n", it, N - lconv);
        if (lconv == N) break;
    }
    return 0;
}
