/*
 * oracle/fsm3d_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp64, no FMA contraction) of the *serial* fast-sweeping
 * eikonal path of the reference, used only as the parity checker by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 * Nothing under mceik_b200/ may link or call this file.
 *
 * PARITY STATUS: "parity unpinned by the reference".  The reference holds no golden
 * vector for the FSM (its only numeric check, fsm3d.f90:2149-2164, is dead code behind
 * `goto 500` at fsm3d.f90:2148) and there is no Fortran compiler in the build container,
 * so this restatement is pinned only by (a) by-construction facts checked in
 * tests/test_oracle_golden.py (source-node values, monotonicity, first-order agreement with
 * the analytic homogeneous field the reference meant to compare with), (b) a line by
 * line reading of the Fortran, and (c) a second reading of the same Fortran lines in
 * another language (tests/fsm_restatement.py, plain Python floats), with which it agrees
 * bit for bit on fields, error codes and iteration counts.  Each function cites the
 * lines it restates.
 *
 * Arithmetic contract (what "bit-exact" means for the CUDA path):
 *   - reference flags are gfortran -O2 without -march (Makefile.inc:4-13): separate
 *     IEEE multiply / add, IEEE sqrt, left-to-right evaluation, no reassociation;
 *     compile this file with -ffp-contract=off.
 *   - constants third = 1/3, two_third = 2/3 are fp64 constants that are MULTIPLIED
 *     (module.F90:7-8, fsm3d.f90:674-675).
 *   - u_nan = HUGE(1.d0) = DBL_MAX (module.F90:419).
 *
 * Indices: the Fortran is 1-based, ijk = (iz-1)*nx*ny + (iy-1)*nx + ix
 * (fsm3d.f90:448).  Here arrays are 0-based and every (ix,iy,iz) triple that is stored
 * or compared is kept 1-based so the boundary quirks of EIKONAL_INIT_GRID survive.
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define U_NAN DBL_MAX

static const double k_zero = 0.0, k_half = 0.5, k_two = 2.0, k_four = 4.0;
static const double k_third = 1.0 / 3.0;     /* module.F90:7 */
static const double k_two_third = 2.0 / 3.0; /* module.F90:8 */

/* ---- level structure: fsm3d.f90:226-308 (MAKE_LEVEL_STRUCT) ------------------------- */
typedef struct {
    int nlevels;
    int *level_ptr;  /* [nlevels+1], 0-based offsets in units of ints (3 per node) */
    int *nnl;        /* [nlevels] */
    int *ixyz_level; /* 1-based (ix,iy,iz) triples */
} oracle_levels_t;

int oracle_make_levels(int nx, int ny, int nz, oracle_levels_t *ls)
{
    long n = (long)nx * ny * nz;
    ls->nlevels = nx + ny + nz - 2;
    ls->level_ptr = (int *)calloc((size_t)ls->nlevels + 1, sizeof(int));
    ls->nnl = (int *)calloc((size_t)ls->nlevels, sizeof(int));
    ls->ixyz_level = (int *)malloc(sizeof(int) * 3 * (size_t)n);
    unsigned char *seen = (unsigned char *)calloc((size_t)n, 1);
    if (!ls->level_ptr || !ls->nnl || !ls->ixyz_level || !seen) return 1;
    int pos = 0;
    /* the Fortran's `level` runs 2..nx+ny+nz-1; s = level-2 is the 0-based coordinate sum */
    for (int s = 0; s <= nx + ny + nz - 3; s++) {
        int k1 = s - (nx - 1) - (ny - 1);
        if (k1 < 0) k1 = 0;
        if (k1 > nz - 1) k1 = nz - 1;
        int k2 = s < nz - 1 ? s : nz - 1;
        int np = 0;
        ls->level_ptr[s] = pos;
        for (int iz = k1; iz <= k2; iz++) {
            int j1 = s - iz - (nx - 1);
            if (j1 < 0) j1 = 0;
            if (j1 > ny - 1) j1 = ny - 1;
            int j2 = s - iz < ny - 1 ? s - iz : ny - 1;
            for (int iy = j1; iy <= j2; iy++) {
                int i1 = s - iz - iy;
                if (i1 < 0) i1 = 0;
                if (i1 > nx - 1) i1 = nx - 1;
                int i2 = s - iz - iy < nx - 1 ? s - iz - iy : nx - 1;
                for (int ix = i1; ix <= i2; ix++) {
                    ls->ixyz_level[pos++] = ix + 1;
                    ls->ixyz_level[pos++] = iy + 1;
                    ls->ixyz_level[pos++] = iz + 1;
                    seen[(long)iz * nx * ny + (long)iy * nx + ix]++;
                    np++;
                }
            }
        }
        ls->nnl[s] = np;
    }
    ls->level_ptr[ls->nlevels] = pos;
    int ierr = 0;
    for (long i = 0; i < n; i++)
        if (seen[i] != 1) ierr = 1; /* fsm3d.f90:301-304 */
    free(seen);
    return ierr;
}

void oracle_free_levels(oracle_levels_t *ls)
{
    free(ls->level_ptr);
    free(ls->nnl);
    free(ls->ixyz_level);
    memset(ls, 0, sizeof(*ls));
}

/* ---- local solvers: fsm3d.f90:562-693 ------------------------------------------------ */
/* SORT3 (fsm3d.f90:562-614): same comparison tree (matters only for equal inputs). */
static inline void sort3(double a, double b, double c, double *a1, double *a2, double *a3)
{
    int lab = !(a > b), lac = !(a > c), lbc = !(b > c);
    if (lab && lac) {
        *a1 = a;
        if (lbc) { *a2 = b; *a3 = c; } else { *a2 = c; *a3 = b; }
    } else if (!lab && lbc) {
        *a1 = b;
        if (lac) { *a2 = a; *a3 = c; } else { *a2 = c; *a3 = a; }
    } else {
        *a1 = c;
        if (lab) { *a2 = a; *a3 = b; } else { *a2 = b; *a3 = a; }
    }
}

/* SOLVE_HAMILTONIAN2D (fsm3d.f90:624-638) */
static inline double hamiltonian2d(double a, double b, double f)
{
    double amb = a - b;
    if (fabs(amb) < f) {
        double arg = k_two * f * f - amb * amb;
        return k_half * (a + b + sqrt(arg));
    }
    return (a < b ? a : b) + f;
}

/* SOLVE_HAMILTONIAN3D (fsm3d.f90:648-693); *ierr mirrors the Fortran codes 0..3 */
double oracle_hamiltonian3d(double a, double b, double c, double f, int *ierr)
{
    double a1, a2, a3;
    *ierr = 0;
    sort3(a, b, c, &a1, &a2, &a3);
    if (a1 == U_NAN) return U_NAN;
    double x = a1 + f;
    if (x > a2) {
        x = hamiltonian2d(a1, a2, f);
        if (x > a3) {
            double qb = -k_two_third * (a1 + a2 + a3);
            double qc = (a1 * a1 + a2 * a2 + a3 * a3 - f * f) * k_third;
            double disc = qb * qb - k_four * qc;
            if (disc < k_zero) *ierr = 1;
            x = k_half * (-qb + sqrt(disc));
            if (x < k_zero) *ierr = 2;
            if (x < U_NAN) return x;
        } else {
            return x;
        }
    } else {
        return x;
    }
    *ierr = 3;
    return U_NAN; /* fall-through: the function result keeps its u_nan initial value */
}

/* UPDATE3D + GET_U{X,Y,Z}MIN3D (fsm3d.f90:460-546); ix,iy,iz 1-based */
static inline int update3d(int nx, int ny, int nz, int ix, int iy, int iz, double h,
                           const double *slow, double *u)
{
    long nxy = (long)nx * ny;
    long ijk = (long)(iz - 1) * nxy + (long)(iy - 1) * nx + (ix - 1);
    double f = slow[ijk] * h;
    double um, up, ux, uy, uz;
    um = ix > 1 ? u[ijk - 1] : u[ijk];
    up = ix < nx ? u[ijk + 1] : u[ijk];
    ux = um < up ? um : up;
    um = iy > 1 ? u[ijk - nx] : u[ijk];
    up = iy < ny ? u[ijk + nx] : u[ijk];
    uy = um < up ? um : up;
    um = iz > 1 ? u[ijk - nxy] : u[ijk];
    up = iz < nz ? u[ijk + nxy] : u[ijk];
    uz = um < up ? um : up;
    int ierr;
    double ubar = oracle_hamiltonian3d(ux, uy, uz, f, &ierr);
    if (ubar < u[ijk]) u[ijk] = ubar;
    return ierr;
}

/* EVAL_UPDATE3D (fsm3d.f90:419-456): one hyperplane, OpenMP over its nodes as :437-440 */
static void eval_update3d(int nx, int ny, int nz, int revx, int revy, int revz, double h,
                          int nnl, const int *ixyz, const unsigned char *lupd,
                          const double *slow, double *u)
{
    long nxy = (long)nx * ny;
#pragma omp parallel for schedule(static) if (nnl > 256)
    for (int ip = 0; ip < nnl; ip++) {
        int ix = ixyz[3 * ip], iy = ixyz[3 * ip + 1], iz = ixyz[3 * ip + 2];
        if (revx) ix = nx + 1 - ix;
        if (revy) iy = ny + 1 - iy;
        if (revz) iz = nz + 1 - iz;
        long ijk = (long)(iz - 1) * nxy + (long)(iy - 1) * nx + (ix - 1);
        if (lupd[ijk]) update3d(nx, ny, nz, ix, iy, iz, h, slow, u);
    }
}

/* Thread count of the OpenMP region above (bench.py's CPU legs: torchrun exports OMP_NUM_THREADS=1).
 * n <= 0 leaves the setting alone; returns the number of threads the next solve will use. */
#ifdef _OPENMP
#include <omp.h>
int oracle_set_threads(int n)
{
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
#else
int oracle_set_threads(int n) { (void)n; return 1; }
#endif

/* sweep table, fsm3d.f90:46-53: (revx,revy,revz) per sweep */
static const int k_sweep[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0},
                                  {0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};

/* EIKONAL3D_FSM (fsm3d.f90:28-99). Returns the number of iterations executed. */
int oracle_fsm(int maxit, int nx, int ny, int nz, double h, double tol,
               const oracle_levels_t *ls, const unsigned char *lupd, const double *slow,
               double *u)
{
    long n = (long)nx * ny * nz;
    double *u0 = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(u0, u, sizeof(double) * (size_t)n);
    int k;
    for (k = 1; k <= maxit; k++) {
        for (int s = 0; s < 8; s++)
            for (int lev = 0; lev < ls->nlevels; lev++)
                eval_update3d(nx, ny, nz, k_sweep[s][0], k_sweep[s][1], k_sweep[s][2], h,
                              ls->nnl[lev], ls->ixyz_level + ls->level_ptr[lev], lupd, slow, u);
        long lconv = 0;
        for (long i = 0; i < n; i++) {
            if (fabs(u0[i] - u[i]) < tol) lconv++;
            u0[i] = u[i];
        }
        if (lconv == n) break;
    }
    free(u0);
    return k > maxit ? maxit : k;
}

/* ---- sources / boundary conditions: fsm3d.f90:697-840 -------------------------------- */
/* EIKONAL_SOURCE_INDEX (fsm3d.f90:697-711); returns the 1-based nearest node */
int oracle_source_index(int nx, double x0, double dx, double xs)
{
    if (xs <= x0) return 1;
    if (xs >= x0 + (double)(float)(nx - 1) * dx) return nx;
    return (int)((xs - x0) / dx + k_half) + 1;
}

/* EIKONAL_INIT_GRID (fsm3d.f90:716-755); ixloc entries are 1-based, -1 = unused */
int oracle_init_grid(int nx, int isx, double x0, double dx, double xs, int ixloc[3])
{
    int np, ierr = 0;
    ixloc[0] = ixloc[1] = ixloc[2] = -1;
    double xs_est = x0 + (double)(float)(isx - 1) * dx;
    if (xs_est > xs) {
        np = 2; ixloc[0] = isx - 1; ixloc[1] = isx;
    } else if (xs_est < xs) {
        np = 2; ixloc[0] = isx; ixloc[1] = isx + 1;
    } else {
        np = 0;
        if (isx > 0) ixloc[np++] = isx - 1;      /* always true for 1-based isx (:736) */
        ixloc[np++] = isx;
        if (isx < nx - 1) ixloc[np++] = isx + 1; /* drops the neighbour at isx = nx-1 (:742) */
    }
    for (int i = 0; i < np; i++)
        if (ixloc[i] < 1 || ixloc[i] > nx) ierr = 1;
    return ierr;
}

/* EIKONAL3D_SETBCS (fsm3d.f90:762-840).  lisbc is 1 byte per node here. */
int oracle_setbcs(int nx, int ny, int nz, int nsrc, double dx, double dy, double dz,
                  double x0, double y0, double z0, const double *ts, const double *xs,
                  const double *ys, const double *zs, const double *slow,
                  unsigned char *lisbc, double *u)
{
    long n = (long)nx * ny * nz, nxy = (long)nx * ny;
    memset(lisbc, 0, (size_t)n);
    for (long i = 0; i < n; i++) u[i] = U_NAN;
    for (int isrc = 0; isrc < nsrc; isrc++) {
        int ixloc[3], iyloc[3], izloc[3];
        int isx = oracle_source_index(nx, x0, dx, xs[isrc]);
        int isy = oracle_source_index(ny, y0, dy, ys[isrc]);
        int isz = oracle_source_index(nz, z0, dz, zs[isrc]);
        if (oracle_init_grid(nx, isx, x0, dx, xs[isrc], ixloc)) return 1;
        if (oracle_init_grid(ny, isy, y0, dy, ys[isrc], iyloc)) return 1;
        if (oracle_init_grid(nz, isz, z0, dz, zs[isrc], izloc)) return 1;
        for (int i = 0; i < 3; i++) {
            if (ixloc[i] == -1) continue;
            for (int j = 0; j < 3; j++) {
                if (iyloc[j] == -1) continue;
                for (int k = 0; k < 3; k++) {
                    if (izloc[k] == -1) continue;
                    int ix = ixloc[i], iy = iyloc[j], iz = izloc[k];
                    long ijk = (long)(iz - 1) * nxy + (long)(iy - 1) * nx + (ix - 1);
                    double x = x0 + (double)(float)(ix - 1) * dx;
                    double y = y0 + (double)(float)(iy - 1) * dy;
                    double z = z0 + (double)(float)(iz - 1) * dz;
                    double ex = xs[isrc] - x, ey = ys[isrc] - y, ez = zs[isrc] - z;
                    double d = sqrt(ex * ex + ey * ey + ez * ez);
                    double t = ts[isrc] + d * slow[ijk];
                    if (fabs(d) < 1.e-10)
                        u[ijk] = t;
                    else if (t < u[ijk])
                        u[ijk] = t;
                    lisbc[ijk] = 1;
                }
            }
        }
    }
    return 0;
}

/* ---- driver: fsm3d.f90:1968-2052 (eikonal3d_serial_driver), jobs 1/2/3 ---------------- */
static oracle_levels_t g_levels;
static int g_linit = 0;
static int g_last_iters = 0;

int oracle_last_iterations(void) { return g_last_iters; }

void oracle_eikonal3d_serial_driver(const int *job, const int *iverb, const int *maxit,
                                    const int *nsrc, const int *nx, const int *ny,
                                    const int *nz, const double *tol, const double *h,
                                    const double *x0, const double *y0, const double *z0,
                                    const double *ts, const double *xs, const double *ys,
                                    const double *zs, const double *slow, double *u, int *ierr)
{
    (void)iverb;
    *ierr = 0;
    if (*job == 1) {
        if (g_linit) { *ierr = 1; return; }
        *ierr = oracle_make_levels(*nx, *ny, *nz, &g_levels);
        if (*ierr) return;
        g_linit = 1;
    } else if (*job == 2) {
        if (!g_linit) { *ierr = 1; return; }
        long n = (long)*nx * *ny * *nz;
        unsigned char *lisbc = (unsigned char *)malloc((size_t)n);
        *ierr = oracle_setbcs(*nx, *ny, *nz, *nsrc, *h, *h, *h, *x0, *y0, *z0, ts, xs, ys, zs,
                              slow, lisbc, u);
        if (*ierr) { free(lisbc); return; }
        for (long i = 0; i < n; i++) lisbc[i] = !lisbc[i]; /* lupd = .NOT. lisbc (:2029-2032) */
        g_last_iters = oracle_fsm(*maxit, *nx, *ny, *nz, *h, *tol, &g_levels, lisbc, slow, u);
        free(lisbc);
    } else {
        if (g_linit) oracle_free_levels(&g_levels);
        g_linit = 0;
    }
}

/* double -> float table extraction: fsm3d.f90:1870-1872 (SNGL), homog.c:624-635 */
void oracle_double2float(long n, const double *x, float *x4)
{
    for (long i = 0; i < n; i++) x4[i] = (float)x[i];
}
