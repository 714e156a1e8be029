/*
 * tests/tile_order_emulation.c -- CPU emulation of the tile-wavefront execution ORDER of
 * mceik_b200/csrc/fsm.cu (16^3 tiles with a clamped 1-node halo, tiles visited by tile
 * hyperplane, 46 local hyperplanes inside a tile, write-back of the interior), using the
 * oracle's local solver.  Test infrastructure: it proves on the CPU, without a GPU, that this
 * order reproduces the global hyperplane order of the reference (fsm3d.f90:62-85) bit for bit.
 * Usage: tile_order_emulation nx ny nz  -> prints "MATCH iters=<k>" or "MISMATCH ...".
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

double oracle_hamiltonian3d(double a, double b, double c, double f, int *ierr);
void oracle_eikonal3d_serial_driver(const int *, const int *, const int *, const int *, const int *, const int *,
                                    const int *, const double *, const double *, const double *, const double *,
                                    const double *, const double *, const double *, const double *, const double *,
                                    const double *, double *, int *);
int oracle_setbcs(int, int, int, int, double, double, double, double, double, double, const double *, const double *,
                  const double *, const double *, const double *, unsigned char *, double *);
int oracle_last_iterations(void);

#define T 16
#define S 18
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

int main(int argc, char **argv)
{
    int nx = argc > 1 ? atoi(argv[1]) : 37, ny = argc > 2 ? atoi(argv[2]) : 50, nz = argc > 3 ? atoi(argv[3]) : 21;
    long n = (long)nx * ny * nz, nxy = (long)nx * ny;
    double h = 100.0, tol = 1e-6, x0 = 0, y0 = 0, z0 = 0;
    int maxit = 20, nsrc = 2, iverb = 0, ierr;
    double ts[2] = {0.0, 0.3}, xs[2] = {h * (nx * 0.37), h * (nx * 0.8)}, ys[2] = {h * (ny * 0.61), h * 2.0},
           zs[2] = {h * (nz * 0.45), h * (nz - 2.5)};
    double *slow = malloc(sizeof(double) * n), *uref = malloc(sizeof(double) * n), *u = malloc(sizeof(double) * n);
    double *u0 = malloc(sizeof(double) * n);
    unsigned char *lisbc = malloc(n);
    srand(7);
    for (long i = 0; i < n; i++) slow[i] = 1.0 / (3000.0 + 2500.0 * (rand() / (double)RAND_MAX));
    int job = 1;
#define DRV() oracle_eikonal3d_serial_driver(&job, &iverb, &maxit, &nsrc, &nx, &ny, &nz, &tol, &h, &x0, &y0, &z0, ts, xs, ys, zs, slow, uref, &ierr)
    DRV(); job = 2; DRV();
    if (ierr) { printf("oracle ierr\n"); return 2; }
    int ref_iters = oracle_last_iterations();
    job = 3; DRV();

    if (oracle_setbcs(nx, ny, nz, nsrc, h, h, h, x0, y0, z0, ts, xs, ys, zs, slow, lisbc, u)) return 2;
    int ntx = (nx + T - 1) / T, nty = (ny + T - 1) / T, ntz = (nz + T - 1) / T;
    static double us[S * S * S];
    int it;
    for (it = 1; it <= maxit; it++) {
        memcpy(u0, u, sizeof(double) * n);
        for (int s = 0; s < 8; s++) {
            int revx = s & 1, revy = (s >> 1) & 1, revz = (s >> 2) & 1;
            for (int d = 0; d < ntx + nty + ntz - 2; d++)
                for (int Ko = 0; Ko < ntz; Ko++)
                    for (int Jo = 0; Jo < nty; Jo++) {
                        int Io = d - Ko - Jo;
                        if (Io < 0 || Io >= ntx) continue;
                        int I = revx ? ntx - 1 - Io : Io, J = revy ? nty - 1 - Jo : Jo, K = revz ? ntz - 1 - Ko : Ko;
                        int xl = I * T, yl = J * T, zl = K * T;
                        int ex = imin(T, nx - xl), ey = imin(T, ny - yl), ez = imin(T, nz - zl);
                        for (int kk = 0; kk < S; kk++)
                            for (int jj = 0; jj < S; jj++)
                                for (int ii = 0; ii < S; ii++) {
                                    int gx = imin(imax(xl + ii - 1, 0), nx - 1), gy = imin(imax(yl + jj - 1, 0), ny - 1),
                                        gz = imin(imax(zl + kk - 1, 0), nz - 1);
                                    us[(kk * S + jj) * S + ii] = u[gz * nxy + (long)gy * nx + gx];
                                }
                        for (int lev = 0; lev < ex + ey + ez - 2; lev++)
                            for (int c = 0; c < T; c++)
                                for (int b = 0; b < T; b++) {
                                    int a = lev - b - c;
                                    if (a < 0 || a >= T || a >= ex || b >= ey || c >= ez) continue;
                                    int i = revx ? ex - 1 - a : a, j = revy ? ey - 1 - b : b, k = revz ? ez - 1 - c : c;
                                    long gi = (long)(zl + k) * nxy + (long)(yl + j) * nx + (xl + i);
                                    if (lisbc[gi]) continue;
                                    int si = ((k + 1) * S + (j + 1)) * S + (i + 1), e;
                                    double ux = fmin(us[si - 1], us[si + 1]), uy = fmin(us[si - S], us[si + S]),
                                           uz = fmin(us[si - S * S], us[si + S * S]);
                                    double x = oracle_hamiltonian3d(ux, uy, uz, slow[gi] * h, &e);
                                    if (x < us[si]) us[si] = x;
                                }
                        for (int k = 0; k < ez; k++)
                            for (int j = 0; j < ey; j++)
                                for (int i = 0; i < ex; i++)
                                    u[(long)(zl + k) * nxy + (long)(yl + j) * nx + (xl + i)] = us[((k + 1) * S + (j + 1)) * S + (i + 1)];
                    }
        }
        long bad = 0;
        for (long i = 0; i < n; i++) if (!(fabs(u0[i] - u[i]) < tol)) bad++;
        if (bad == 0) break;
    }
    if (it > maxit) it = maxit;
    long diff = 0;
    for (long i = 0; i < n; i++) if (memcmp(&u[i], &uref[i], sizeof(double)) != 0) diff++;
    if (diff == 0 && it == ref_iters) printf("MATCH iters=%d\n", it);
    else printf("MISMATCH nodes=%ld iters=%d ref_iters=%d\n", diff, it, ref_iters);
    return diff != 0 || it != ref_iters;
}
