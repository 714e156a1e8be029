/*
 * tests/brick_pipeline_emulation.c -- CPU model of the streaming brick kernel's SCHEDULE
 * (mceik_b200/csrc/fsm_bricks16.cu), using the oracle's local solver.  Test infrastructure.
 *
 * What is modelled, with the kernel's own constants:
 *   - bricks of 8 x 8 x zc nodes; a brick (one warp) advances in steps; at step l it updates its nodes
 *     with i + j + k = l (sweep-frame brick coordinates);
 *   - cell (i,j,k), i in [-1,8], j in [-1,ey], k in [-1,ez], belongs to ring slot
 *     m = xgroup(i) + (j+1) + (k+1); slot m is LOADED from global memory (halo cells outside the grid
 *     clamped to the boundary node) when step m - 6 starts -- long before it is used -- and the in-brick
 *     cells of slot l - 4 (final since step l - 1) are STORED in step l;
 *   - a brick may load slot m only when its upwind y neighbour has completed m + By + 5 steps and its
 *     upwind x neighbour m + 7 (the fine-grained dependencies of the kernel on the [z][y][x] layout; on the
 *     blocked layout the x halo comes from face copies stored later and the kernel asks for m + By + 7),
 *     and starts only after its
 *     upwind z neighbour has finished;
 *   - the 8 sweeps of an iteration OVERLAP as in the kernel: a brick starts sweep s as soon as it and its six face
 *     neighbours have completed sweep s - 1 (and its upwind z neighbour sweep s), while other bricks are still in
 *     earlier sweeps; only the iterations are separated (one kernel launch each);
 *   - bricks that may advance are picked in RANDOM order, so every run explores another interleaving.
 * It asserts the ring lifetime the kernel relies on (a cell is read in steps [m-4, m+4] and written in
 * [m-3, m+3], i.e. 11 slots with the two prefetched ones) and that every value read had been loaded.
 * The result must equal the reference-ordered oracle bit for bit for every interleaving; with a lead
 * one step shorter (argv) it must not be relied upon -- the test checks that the model notices.
 *
 * Optional study (argv[7] = 1), not part of the kernel yet: skip a brick's sweep when it and its six face neighbours
 * changed nothing in the previous sweep and its upwind neighbours have finished the current one without a change
 * (an update whose inputs did not change is idempotent).  The model reports how many brick sweeps that skips and
 * must still give the oracle's bits.
 *
 * Blocked layout (argv[9] = 1; the kernel's default data path): the x halo of a brick is not read from the field but
 * from FACE COPIES -- per brick column and plane, the values of its memory columns 0 and 7 -- which the owning brick
 * stores itself, in half planes: the four rows its sweep enters first 11 (sweep column 0) / 12 (sweep column 7)
 * steps after the plane was entered, the other four 15 / 16 steps after (brick columns cut by the grid's y face:
 * all rows 15 / 16).  The upwind x neighbour must then have completed m + By + 3 steps before slot m is loaded
 * (m + By + 7 for cut columns); at the grid's x faces a brick reads its own face copy.  The model checks that every
 * face value stored was final (written back in an earlier step), that an upwind face value read belongs to the
 * current sweep, and that own / downwind face values read are still those of the previous sweep.
 *
 * Usage: brick_pipeline_emulation nx ny nz zc seed [dlead [skip [checkerboard cell [blocked]]]]  -> "MATCH iters=<k> stalls=<n> ..." or "MISMATCH ..."
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

#define BX 8
#define BY 8
#define PREFETCH 2
#define AHEAD (4 + PREFETCH)          /* slot l + AHEAD is loaded when step l starts */

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
static int xgroup(int i) { return (i + 4) >> 2; }

typedef struct {
    int I, J, K;            /* global brick coordinates */
    int x_lo, y_lo, y_hi, z_lo, z_hi, ey, ez;
    int sweep;              /* sweep being worked on (0..7), 8 when the iteration is finished */
    int progress;           /* steps completed in that sweep */
    int loaded;             /* highest slot loaded so far, -1 before the start */
    long last_changed;      /* global sweep (8 * iteration + sweep) in which the brick last changed a node */
    int changed;            /* ... in the sweep being worked on */
    double *L;              /* (BX+2) x (BY+2) x (ez+2) local cells */
    unsigned char *have;    /* cell has been loaded */
} brick_t;

static int nx, ny, nz, zc, nbx, nby, nbz;
static long nxy;
static double h;
static const double *slow;
static const unsigned char *lisbc;
static double *u;
static int lead_x, lead_y, dlead_g;
static int blocked;         /* 1 = x halo from face copies (the kernel's blocked layout) */
static double *face;        /* [side][brick column J * nbx + I][gz][row] values of memory columns 0 / 7 */
static long *face_sweep;    /* global sweep in which the entry was last stored (-1: by the boundary conditions) */
static long face_not_final, face_stale, face_early;
#define FIDX(side, I, J, gz, row) (((((long)(side) * nby + (J)) * nbx + (I)) * nz + (gz)) * BY + (row))
#define REVX(s) ((s) & 1)
#define REVY(s) (((s) >> 1) & 1)
#define REVZ(s) (((s) >> 2) & 1)   /* fsm3d.f90:46-53 */
static long ring_violations, unloaded_reads, gsweep0, skipped, tasks;
static int skip_mode;

#define LIDX(b, i, j, k) ((((long)(k) + 1) * (BY + 2) + ((j) + 1)) * (BX + 2) + ((i) + 1))

static long gnode(const brick_t *b, int i, int j, int k)
{ /* sweep-frame brick cell -> clamped global node */
    const int sw = b->sweep;
    int gx = REVX(sw) ? b->x_lo + BX - 1 - i : b->x_lo + i;
    int gy = REVY(sw) ? b->y_hi - j : b->y_lo + j;
    int gz = REVZ(sw) ? b->z_hi - k : b->z_lo + k;
    gx = imin(imax(gx, 0), nx - 1); gy = imin(imax(gy, 0), ny - 1); gz = imin(imax(gz, 0), nz - 1);
    return (long)gz * nxy + (long)gy * nx + gx;
}

static brick_t *brick_at(brick_t *bk, int I, int J, int K)
{
    if (I < 0 || I >= nbx || J < 0 || J >= nby || K < 0 || K >= nbz) return NULL;
    return bk + ((long)K * nby + J) * nbx + I;
}

/* progress word of the kernel: (sweep << 12) + steps completed in it; a finished sweep counts as (sweep + 1) << 12 */
static int word(const brick_t *b) { return (b->sweep << 12) + b->progress; }

static int may_load(brick_t *bk, const brick_t *b, int m)
{
    const int sw = b->sweep;
    const brick_t *ux = brick_at(bk, b->I + (REVX(sw) ? 1 : -1), b->J, b->K);
    const brick_t *uy = brick_at(bk, b->I, b->J + (REVY(sw) ? 1 : -1), b->K);
    /* blocked: the x halo comes out of the neighbour's face copies, stored in half planes (whole planes when cut in y) */
    const int lx = blocked ? (b->ey == BY ? BY + 3 : BY + 7) + dlead_g : lead_x;
    if (ux && word(ux) < (sw << 12) + m + lx) return 0;
    if (uy && word(uy) < (sw << 12) + m + lead_y) return 0;
    return 1;
}

static int may_start(brick_t *bk, const brick_t *b)
{ /* the brick and its six face neighbours have completed the previous sweep, the upwind z neighbour this one */
    const int sw = b->sweep;
    static const int d[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    for (int q = 0; q < 6; q++) {
        const brick_t *nb = brick_at(bk, b->I + d[q][0], b->J + d[q][1], b->K + d[q][2]);
        if (!nb) continue;
        int need = sw << 12;
        if (d[q][2] != 0 && d[q][2] == (REVZ(sw) ? 1 : -1)) need = (sw + 1) << 12;
        if (word(nb) < need) return 0;
    }
    return 1;
}

/* skip study: 1 = skip this sweep, 0 = run it, -1 = cannot tell yet (an upwind neighbour is still sweeping) */
static int may_skip(brick_t *bk, const brick_t *b)
{
    const int sw = b->sweep;
    const long G = gsweep0 + sw;
    static const int d[7][3] = {{0, 0, 0}, {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    for (int q = 0; q < 7; q++) {
        const brick_t *nb = brick_at(bk, b->I + d[q][0], b->J + d[q][1], b->K + d[q][2]);
        if (nb && nb->last_changed >= G - 1) return 0;   /* changed in the previous sweep, or already in this one */
    }
    const int up[3][3] = {{REVX(sw) ? 1 : -1, 0, 0}, {0, REVY(sw) ? 1 : -1, 0}, {0, 0, REVZ(sw) ? 1 : -1}};
    for (int q = 0; q < 3; q++) {
        const brick_t *nb = brick_at(bk, b->I + up[q][0], b->J + up[q][1], b->K + up[q][2]);
        if (nb && word(nb) < ((sw + 1) << 12)) return -1;
    }
    return 1;
}

static void load_slot(brick_t *b, int m)
{
    for (int k = -1; k <= b->ez; k++)
        for (int j = -1; j <= b->ey; j++)
            for (int i = -1; i <= BX; i++)
                if (xgroup(i) + (j + 1) + (k + 1) == m) {
                    double v = u[gnode(b, i, j, k)];
                    if (blocked && (i == -1 || i == BX) && j >= 0 && j < b->ey && k >= 0 && k < b->ez) {
                        const int sw = b->sweep;
                        const int gx = REVX(sw) ? b->x_lo + BX - 1 - i : b->x_lo + i;   /* not clamped */
                        const int gy = REVY(sw) ? b->y_hi - j : b->y_lo + j, gz = REVZ(sw) ? b->z_hi - k : b->z_lo + k;
                        int In = b->I, side;
                        if (gx >= 0 && gx < nx) { In = gx / BX; side = (gx % BX == 0) ? 0 : 1; }  /* the neighbour's adjacent column */
                        else side = gx < 0 ? 0 : 1;                                              /* grid face: own boundary column */
                        const long fi = FIDX(side, In, b->J, gz, gy - b->y_lo);
                        v = face[fi];
                        const long G = gsweep0 + sw;
                        if (In != b->I && i == -1) { if (face_sweep[fi] != G) face_stale++; }     /* upwind: this sweep's value */
                        else if (face_sweep[fi] == G) face_early++;                               /* own / downwind: not yet */
                    }
                    b->L[LIDX(b, i, j, k)] = v;
                    b->have[LIDX(b, i, j, k)] = 1;
                }
    b->loaded = m;
}

static double rd(brick_t *b, int i, int j, int k, int l)
{
    const int m = xgroup(i) + (j + 1) + (k + 1);
    if (l < m - 4 || l > m + 4) ring_violations++;
    if (!b->have[LIDX(b, i, j, k)] || m > l + 4) unloaded_reads++;
    return b->L[LIDX(b, i, j, k)];
}

static void step(brick_t *b)
{
    const int l = b->progress;
    for (int k = 0; k < b->ez; k++)
        for (int j = 0; j < b->ey; j++) {
            const int i = l - j - k;
            if (i < 0 || i >= BX) continue;
            const long g = gnode(b, i, j, k);
            if (lisbc[g]) continue;
            const int m = xgroup(i) + (j + 1) + (k + 1);
            if (l < m - 3 || l > m + 3) ring_violations++;
            const double self = rd(b, i, j, k, l);
            const double xm = rd(b, i - 1, j, k, l), xp = rd(b, i + 1, j, k, l);
            const double ym = rd(b, i, j - 1, k, l), yp = rd(b, i, j + 1, k, l);
            const double zm = rd(b, i, j, k - 1, l), zp = rd(b, i, j, k + 1, l);
            const double ux = xm < xp ? xm : xp, uy = ym < yp ? ym : yp, uz = zm < zp ? zm : zp;
            int ierr;
            const double ubar = oracle_hamiltonian3d(ux, uy, uz, slow[g] * h, &ierr);
            if (ubar < self) { b->L[LIDX(b, i, j, k)] = ubar; b->changed = 1; }
        }
    /* write back the in-brick cells of slot l - 4 */
    for (int k = 0; k < b->ez; k++)
        for (int j = 0; j < b->ey; j++)
            for (int i = 0; i < BX; i++)
                if (xgroup(i) + (j + 1) + (k + 1) == l - 4) u[gnode(b, i, j, k)] = b->L[LIDX(b, i, j, k)];
    if (blocked) {  /* face planes: half planes of each side, a fixed number of steps after the plane was entered */
        const int sw = b->sweep;
        for (int side = 0; side < 2; side++)
            for (int half = 0; half < 2; half++) {
                const int is = ((side == 0) == !REVX(sw)) ? 0 : BX - 1;          /* sweep column of this memory column */
                const int lag = (is == 0 ? 15 : 16) - ((half == 0 && b->ey == BY) ? 4 : 0);
                const int kf = l - lag;
                if (kf < 0 || kf >= b->ez) continue;
                const int j0 = b->ey == BY ? 4 * half : (half ? b->ey : 0), j1 = b->ey == BY ? 4 * half + 4 : b->ey;
                for (int j = j0; j < j1; j++) {
                    if (xgroup(is) + (j + 1) + (kf + 1) > l - 5) face_not_final++;  /* written back before this step? */
                    const int gy = REVY(sw) ? b->y_hi - j : b->y_lo + j, gz = REVZ(sw) ? b->z_hi - kf : b->z_lo + kf;
                    const long fi = FIDX(side, b->I, b->J, gz, gy - b->y_lo);
                    face[fi] = b->L[LIDX(b, is, j, kf)];
                    face_sweep[fi] = gsweep0 + sw;
                }
            }
    }
    b->progress = l + 1;
}

static long iteration(brick_t *bk, long nb, unsigned *rng)
{
    long stalls = 0, left = nb, idle = 0;
    for (long q = 0; q < nb; q++) { bk[q].sweep = 0; bk[q].progress = 0; bk[q].loaded = -1; }
    while (left > 0) {
        *rng = *rng * 1664525u + 1013904223u;
        brick_t *b = bk + (*rng >> 8) % nb;
        if (b->sweep == 8) continue;
        const int nsteps = b->ez + BY + 7 + (blocked ? 1 : 0);  /* blocked: one more step stores the last face plane */
        int moved = 0;
        if (b->loaded < 0) {  /* start of a sweep: slots 0 .. AHEAD-1 */
            const int sk = (skip_mode && may_start(bk, b)) ? may_skip(bk, b) : 0;
            if (sk == 1) {
                if (blocked)  /* nothing changes, so the face copies of the previous sweep are this sweep's too */
                    for (int side = 0; side < 2; side++)
                        for (int gz = b->z_lo; gz <= b->z_hi; gz++)
                            for (int row = 0; row < b->ey; row++) face_sweep[FIDX(side, b->I, b->J, gz, row)] = gsweep0 + b->sweep;
                b->sweep++; skipped++; tasks++;
                if (b->sweep == 8) left--;
                moved = 1;
            } else if (sk == 0 && may_start(bk, b) && may_load(bk, b, AHEAD - 1)) {
                b->changed = 0;
                memset(b->have, 0, (size_t)(BX + 2) * (BY + 2) * (b->ez + 2));
                for (int m = 0; m < AHEAD; m++) load_slot(b, m);
                moved = 1;
            }
        } else if (may_load(bk, b, b->progress + AHEAD)) {
            load_slot(b, b->progress + AHEAD);
            step(b);
            if (b->progress == nsteps) {  /* (sweep + 1) << 12 in the kernel's progress word */
                if (b->changed) b->last_changed = gsweep0 + b->sweep;
                b->sweep++; b->progress = 0; b->loaded = -1; tasks++;
                if (b->sweep == 8) left--;
            }
            moved = 1;
        }
        if (moved) idle = 0;
        else { stalls++; if (++idle > 200 * nb + 100000) { printf("DEADLOCK\n"); exit(3); } }
    }
    return stalls;
}

int main(int argc, char **argv)
{
    nx = argc > 1 ? atoi(argv[1]) : 24; ny = argc > 2 ? atoi(argv[2]) : 20; nz = argc > 3 ? atoi(argv[3]) : 22;
    zc = argc > 4 ? atoi(argv[4]) : 16;
    unsigned rng = argc > 5 ? (unsigned)atoi(argv[5]) : 1u;
    const int dlead = argc > 6 ? atoi(argv[6]) : 0;
    skip_mode = argc > 7 ? atoi(argv[7]) : 0;
    lead_x = 7 + dlead; lead_y = BY + 5 + dlead; dlead_g = dlead;
    blocked = argc > 9 ? atoi(argv[9]) : 0;
    if (nx % BX) { printf("nx must be a multiple of 8\n"); return 2; }
    nxy = (long)nx * ny;
    long n = nxy * nz;
    h = 100.0;
    double tol = 1e-6, x0 = 0, y0 = 0, z0 = 0;
    int maxit = 20, nsrc = 2, iverb = 0, ierr;
    double ts[2] = {0.0, 0.3}, xs[2] = {h * (nx * 0.37), h * (nx * 0.8)}, ys[2] = {h * (ny * 0.61), h * 2.0},
           zs[2] = {h * (nz * 0.45), h * (nz - 2.5)};
    double *sl = malloc(sizeof(double) * n), *uref = malloc(sizeof(double) * n), *u0 = malloc(sizeof(double) * n);
    unsigned char *bc = malloc(n);
    u = malloc(sizeof(double) * n);
    srand(7);
    const int checker = argc > 8 ? atoi(argv[8]) : 0;  /* cell size of a +-10 % checkerboard instead of the random model */
    for (long i = 0; i < n; i++) {
        if (checker) {
            const int ix = (int)(i % nx), iy = (int)((i / nx) % ny), iz = (int)(i / nxy);
            const int sgn = ((ix / checker + iy / checker + iz / checker) & 1) ? -1 : 1;
            sl[i] = 1.0 / (5000.0 * (1.0 + 0.1 * sgn));
        } else {
            sl[i] = 1.0 / (3000.0 + 2500.0 * (rand() / (double)RAND_MAX));
        }
    }
    slow = sl; lisbc = bc;
    int job = 1;
#define DRV() oracle_eikonal3d_serial_driver(&job, &iverb, &maxit, &nsrc, &nx, &ny, &nz, &tol, &h, &x0, &y0, &z0, ts, xs, ys, zs, sl, uref, &ierr)
    DRV(); job = 2; DRV();
    if (ierr) { printf("oracle ierr\n"); return 2; }
    const int ref_iters = oracle_last_iterations();
    job = 3; DRV();
    if (oracle_setbcs(nx, ny, nz, nsrc, h, h, h, x0, y0, z0, ts, xs, ys, zs, sl, bc, u)) return 2;

    nbx = nx / BX; nby = (ny + BY - 1) / BY; nbz = (nz + zc - 1) / zc;
    const long nb = (long)nbx * nby * nbz;
    brick_t *bk = calloc(nb, sizeof(brick_t));
    for (int K = 0; K < nbz; K++) for (int J = 0; J < nby; J++) for (int I = 0; I < nbx; I++) {
        brick_t *b = brick_at(bk, I, J, K);
        b->I = I; b->J = J; b->K = K;
        b->x_lo = I * BX; b->y_lo = J * BY; b->y_hi = imin(b->y_lo + BY, ny) - 1;
        b->z_lo = K * zc; b->z_hi = imin(b->z_lo + zc, nz) - 1;
        b->ey = b->y_hi - b->y_lo + 1; b->ez = b->z_hi - b->z_lo + 1;
        b->L = malloc(sizeof(double) * (BX + 2) * (BY + 2) * (b->ez + 2));
        b->have = malloc((size_t)(BX + 2) * (BY + 2) * (b->ez + 2));
        b->last_changed = 7;   /* "changed in the sweep before the first one": nothing is skipped at the start */
    }
    if (blocked) {  /* face copies of the boundary-condition state (apply_bcs_blocked_kernel writes them too) */
        face = malloc(sizeof(double) * 2 * nbx * nby * nz * BY);
        face_sweep = malloc(sizeof(long) * 2 * nbx * nby * nz * BY);
        for (int side = 0; side < 2; side++) for (int J = 0; J < nby; J++) for (int I = 0; I < nbx; I++)
            for (int gz = 0; gz < nz; gz++) for (int row = 0; row < BY; row++) {
                const int gy = J * BY + row;
                face[FIDX(side, I, J, gz, row)] = gy < ny ? u[(long)gz * nxy + (long)gy * nx + I * BX + (side ? BX - 1 : 0)] : DBL_MAX;
                face_sweep[FIDX(side, I, J, gz, row)] = -1;
            }
    }
    long stalls = 0;
    int it;
    for (it = 1; it <= maxit; it++) {
        memcpy(u0, u, sizeof(double) * n);
        gsweep0 = 8L * it;
        stalls += iteration(bk, nb, &rng);
        long lconv = 0;
        for (long i = 0; i < n; i++) if (fabs(u0[i] - u[i]) < tol) lconv++;
        if (lconv == n) break;
    }
    if (it > maxit) it = maxit;
    long bad = 0;
    for (long i = 0; i < n; i++) if (u[i] != uref[i]) bad++;
    const long face_bad = face_not_final + face_stale + face_early;
    if (bad || it != ref_iters || ring_violations || unloaded_reads || face_bad)
        printf("MISMATCH nodes=%ld iters=%d ref_iters=%d ring_violations=%ld unloaded_reads=%ld face: not_final=%ld stale=%ld early=%ld\n",
               bad, it, ref_iters, ring_violations, unloaded_reads, face_not_final, face_stale, face_early);
    else
        printf("MATCH iters=%d stalls=%ld skipped=%ld of %ld brick sweeps\n", it, stalls, skipped, tasks);
    return (bad || it != ref_iters || ring_violations || unloaded_reads || face_bad) ? 1 : 0;
}
