/*
 * oracle/locate_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's L2 travel-time-table grid search, in its three
 * flavours, plus the analytic homogeneous table generator.  Used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY STATUS: pinned.  oracle_l2_gridsearch_{f64,f32} and oracle_minloc_* are checked
 * bit-for-bit against the reference's own locate.c compiled unmodified into
 * oracle/_ref/libref_locate.so (recipe: oracle/Makefile `ref`), and against the known
 * answer of its main (locate.c:118-182: 145x145x45 grid, srand(4042), 20 picks -> flat
 * index 107312, t0 = 4.0).  The gridsearch.f90 flavour has no Fortran compiler to run
 * against and is pinned by its by-construction answer (gridsearch.f90:38-87: 79x71x15,
 * source node (31,55,4) -> 1-based index 21124, t0 = 4.0) and by equality with the C
 * flavour when every variance is 1.
 *
 * Compile with -O2 -ffp-contract=off: the reference is built without -march
 * (Makefile.inc:4-13) so it has no fused multiply-adds; accumulation runs over
 * observations in catalogue order, one grid point at a time.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* literal from locate.c:496 (1/sqrt(2) rounded DOWN to ...BCC, not M_SQRT1_2) */
#define SQRT2I_D 0.7071067811865475
#define SQRT2I_F 0.7071067811865475f

/* ---- flavour 1: locate.c:923-1047 (double) --------------------------------------------
 * mask[i]==0 -> used.  wt = 1/var, xnorm = sum wt (in pick order), tobsCor = tobs-tcorr.
 * t0[g]   = sum_i (wt_i/xnorm) * (tobsCor_i - test[i*ldgrd+g])          (locate.c:399,409)
 * obj[g]  = sum_i ( (wt_i*sqrt2i) * (tobsCor_i - (test[..]+t0[g])) )^2   (locate.c:500,511)
 * returns 1 on the argument errors of locate.c:948-974.                                  */
int oracle_l2_gridsearch_f64(int ldgrd, int ngrd, int nobs, int iwantOT, double t0use,
                             const int *mask, const double *tobs, const double *tcorr,
                             const double *varobs, const double *test, double *t0,
                             double *objfn)
{
    if ((sizeof(double) * (size_t)ldgrd) % 64 != 0 || ldgrd < ngrd || nobs < 1 || !mask ||
        !tobs || !varobs || !test || !t0 || !objfn)
        return 1;
    if ((uintptr_t)t0 % 64 || (uintptr_t)test % 64 || (uintptr_t)objfn % 64) return 1;
    double *tc = (double *)malloc(sizeof(double) * (size_t)nobs);
    double *wt = (double *)malloc(sizeof(double) * (size_t)nobs);
    int *ptr = (int *)malloc(sizeof(int) * (size_t)nobs);
    int nuse = 0;
    double xnorm = 0.0;
    for (int i = 0; i < nobs; i++) {
        if (mask[i] != 0) continue;
        tc[nuse] = tcorr ? tobs[i] - tcorr[i] : tobs[i];
        wt[nuse] = 1.0 / varobs[i];
        xnorm = xnorm + wt[nuse];
        ptr[nuse++] = i;
    }
    for (int g = 0; g < ngrd; g++) objfn[g] = 0.0;
    if (iwantOT == 1) {
        for (int g = 0; g < ngrd; g++) t0[g] = 0.0;
        for (int j = 0; j < nuse; j++) {
            const double *tt = test + (size_t)ldgrd * (size_t)ptr[j];
            double w = wt[j] / xnorm, to = tc[j];
            for (int g = 0; g < ngrd; g++) t0[g] = t0[g] + w * (to - tt[g]);
        }
    } else {
        for (int g = 0; g < ngrd; g++) t0[g] = t0use;
    }
    for (int j = 0; j < nuse; j++) {
        const double *tt = test + (size_t)ldgrd * (size_t)ptr[j];
        double w = wt[j] * SQRT2I_D, to = tc[j];
        for (int g = 0; g < ngrd; g++) {
            double res = w * (to - (tt[g] + t0[g]));
            objfn[g] = objfn[g] + res * res;
        }
    }
    free(tc); free(wt); free(ptr);
    return 0;
}

/* ---- flavour 1, single precision: locate.c:1079-1203 --------------------------------- */
int oracle_l2_gridsearch_f32(int ldgrd, int ngrd, int nobs, int iwantOT, float t0use,
                             const int *mask, const float *tobs, const float *tcorr,
                             const float *varobs, const float *test, float *t0, float *objfn)
{
    if ((sizeof(float) * (size_t)ldgrd) % 64 != 0 || ldgrd < ngrd || nobs < 1 || !mask ||
        !tobs || !varobs || !test || !t0 || !objfn)
        return 1;
    if ((uintptr_t)t0 % 64 || (uintptr_t)test % 64 || (uintptr_t)objfn % 64) return 1;
    float *tc = (float *)malloc(sizeof(float) * (size_t)nobs);
    float *wt = (float *)malloc(sizeof(float) * (size_t)nobs);
    int *ptr = (int *)malloc(sizeof(int) * (size_t)nobs);
    int nuse = 0;
    float xnorm = 0.0f;
    for (int i = 0; i < nobs; i++) {
        if (mask[i] != 0) continue;
        tc[nuse] = tcorr ? tobs[i] - tcorr[i] : tobs[i];
        wt[nuse] = 1.0f / varobs[i];
        xnorm = xnorm + wt[nuse];
        ptr[nuse++] = i;
    }
    for (int g = 0; g < ngrd; g++) objfn[g] = 0.0f;
    if (iwantOT == 1) {
        for (int g = 0; g < ngrd; g++) t0[g] = 0.0f;
        for (int j = 0; j < nuse; j++) {
            const float *tt = test + (size_t)ldgrd * (size_t)ptr[j];
            float w = wt[j] / xnorm, to = tc[j];
            for (int g = 0; g < ngrd; g++) t0[g] = t0[g] + w * (to - tt[g]);
        }
    } else {
        for (int g = 0; g < ngrd; g++) t0[g] = t0use;
    }
    for (int j = 0; j < nuse; j++) {
        const float *tt = test + (size_t)ldgrd * (size_t)ptr[j];
        float w = wt[j] * SQRT2I_F, to = tc[j];
        for (int g = 0; g < ngrd; g++) {
            float res = w * (to - (tt[g] + t0[g]));
            objfn[g] = objfn[g] + res * res;
        }
    }
    free(tc); free(wt); free(ptr);
    return 0;
}

/* locate_minLoc{Double64,Float64}: locate.c:811-851 -- first strict minimum, 0-based.
 * A NaN in x[0] is sticky (every later `<` is false), as in the reference.            */
int oracle_minloc_f64(int n, const double *x)
{
    double xmin = x[0];
    int imin = 0;
    for (int i = 1; i < n; i++)
        if (x[i] < xmin) { imin = i; xmin = x[i]; }
    return imin;
}
int oracle_minloc_f32(int n, const float *x)
{
    float xmin = x[0];
    int imin = 0;
    for (int i = 1; i < n; i++)
        if (x[i] < xmin) { imin = i; xmin = x[i]; }
    return imin;
}

/* ---- flavour 2: gridsearch.f90:382-459 (double), :463-540 (float) ----------------------
 * mask==1 skips.  xnorm = sum of *variances* of unmasked picks (:431-434); t0 weight
 * 1/(var_i*xnorm) (:185); objective weight sqrt2i/var_i with sqrt2i = one/SQRT(two)
 * (:272-273) -- evaluated here with the same two IEEE operations.  iwantOT != 1 leaves
 * t0 = 0.  t0_out (may be NULL) exposes the internal t0 the Fortran only prints (:454).  */
void oracle_gridsearch_f90_f64(int ldgrd, int ngrd, int nobs, int iwantOT, const int *mask,
                               const double *tobs, const double *varobs, const double *test,
                               double *logPDF, double *t0_out, int *ierr)
{
    *ierr = 0;
    if (ldgrd % 64 != 0) { *ierr = 1; return; }
    if (ngrd > ldgrd) { *ierr = 1; return; }
    int msum = 0;
    double vsum = 0.0;
    for (int i = 0; i < nobs; i++) { msum += mask[i]; vsum = vsum + varobs[i]; }
    if (msum == nobs) { *ierr = 1; return; }
    if (fabs(vsum - 0.0) < 2.220446049250313e-16) { *ierr = 1; return; }
    double *t0 = (double *)calloc((size_t)ngrd, sizeof(double));
    for (int g = 0; g < ngrd; g++) logPDF[g] = 0.0;
    if (iwantOT == 1) {
        double xnorm = 0.0;
        for (int i = 0; i < nobs; i++)
            if (mask[i] != 1) xnorm = xnorm + varobs[i];
        for (int i = 0; i < nobs; i++) {
            if (mask[i] == 1) continue;
            const double *tt = test + (size_t)ldgrd * (size_t)i;
            double w = 1.0 / (varobs[i] * xnorm), to = tobs[i];
            for (int g = 0; g < ngrd; g++) t0[g] = t0[g] + w * (to - tt[g]);
        }
    }
    const double sqrt2i = 1.0 / sqrt(2.0);
    for (int i = 0; i < nobs; i++) {
        if (mask[i] == 1) continue;
        const double *tt = test + (size_t)ldgrd * (size_t)i;
        double w = sqrt2i / varobs[i], to = tobs[i];
        for (int g = 0; g < ngrd; g++) {
            double res = w * (to - (tt[g] + t0[g]));
            logPDF[g] = logPDF[g] + res * res;
        }
    }
    if (t0_out) memcpy(t0_out, t0, sizeof(double) * (size_t)ngrd);
    free(t0);
}

void oracle_gridsearch_f90_f32(int ldgrd, int ngrd, int nobs, int iwantOT, const int *mask,
                               const float *tobs, const float *varobs, const float *test,
                               float *logPDF, float *t0_out, int *ierr)
{
    *ierr = 0;
    if (ldgrd % 64 != 0) { *ierr = 1; return; }
    if (ngrd > ldgrd) { *ierr = 1; return; }
    int msum = 0;
    float vsum = 0.0f;
    for (int i = 0; i < nobs; i++) { msum += mask[i]; vsum = vsum + varobs[i]; }
    if (msum == nobs) { *ierr = 1; return; }
    if (fabsf(vsum - 0.0f) < 1.1920929e-07f) { *ierr = 1; return; }
    float *t0 = (float *)calloc((size_t)ngrd, sizeof(float));
    for (int g = 0; g < ngrd; g++) logPDF[g] = 0.0f;
    if (iwantOT == 1) {
        float xnorm = 0.0f;
        for (int i = 0; i < nobs; i++)
            if (mask[i] != 1) xnorm = xnorm + varobs[i];
        for (int i = 0; i < nobs; i++) {
            if (mask[i] == 1) continue;
            const float *tt = test + (size_t)ldgrd * (size_t)i;
            float w = 1.0f / (varobs[i] * xnorm), to = tobs[i];
            for (int g = 0; g < ngrd; g++) t0[g] = t0[g] + w * (to - tt[g]);
        }
    }
    const float sqrt2i = 1.0f / sqrtf(2.0f);
    for (int i = 0; i < nobs; i++) {
        if (mask[i] == 1) continue;
        const float *tt = test + (size_t)ldgrd * (size_t)i;
        float w = sqrt2i / varobs[i], to = tobs[i];
        for (int g = 0; g < ngrd; g++) {
            float res = w * (to - (tt[g] + t0[g]));
            logPDF[g] = logPDF[g] + res * res;
        }
    }
    if (t0_out) memcpy(t0_out, t0, sizeof(float) * (size_t)ngrd);
    free(t0);
}

/* ---- flavour 3: the catalogue contract of locate.f90:322-519 ---------------------------
 * Interface and data flow of LOCATE3D_GRIDSEARCH (rectangular [nevents x nobs] inputs,
 * table = (statPtr, pickType), tobs - statCor(iobs), job 1 = fixed tori, job 2 = analytic
 * t0, fp32 tables promoted to fp64 (:414,:459), logPDF = -objective, first-index MAXLOC,
 * hypo = (x,y,z,t0)) with the arithmetic of flavour 1 (SURVEY.md section 8a "canonical
 * variant": the t0 scaling of locate.f90:410-426, the un-offset luseObs index of :396,:439
 * and the hypo(:)=0 wipe of :471 are reference defects that are NOT reproduced).
 *
 *   tables  : fp32 [ntables][ldgrd], table id = 2*(station-1) + (pickType-1)
 *   luseObs, statPtr, pickType, varobs, tobs : [nevents*nobs], event-major
 *   xlocs,ylocs,zlocs : fp32 node coordinates [ngrd] (locate.f90:646-652)
 *   out: hypo[4*nevents], iopt[nevents] (0-based), objmin[nevents] (= -logPDF(iopt))
 * Events with no usable pick get iopt=-1 and hypo=0.                                     */
int oracle_locate3d_catalog(int job, int ngrd, size_t ldgrd, int ntables, const float *tables,
                            int nobs, int nevents, const int *luseObs, const int *statPtr,
                            const int *pickType, const double *statCor, const double *tori,
                            const double *varobs, const double *tobs, const float *xlocs,
                            const float *ylocs, const float *zlocs, double *hypo, int *iopt,
                            double *objmin)
{
    if (job != 1 && job != 2) return 1; /* jobs 3,5: "Not yet done" (locate.f90:502-512) */
    double *t0 = (double *)malloc(sizeof(double) * (size_t)ngrd);
    double *obj = (double *)malloc(sizeof(double) * (size_t)ngrd);
    int rc = 0;
    for (int e = 0; e < nevents; e++) {
        const size_t o0 = (size_t)e * (size_t)nobs;
        double xnorm = 0.0;
        int nuse = 0;
        for (int i = 0; i < nobs; i++) {
            if (luseObs[o0 + i] == 0) continue;
            int tid = 2 * (statPtr[o0 + i] - 1) + (pickType[o0 + i] - 1);
            if (tid < 0 || tid >= ntables) { rc = 1; goto done; }
            xnorm = xnorm + 1.0 / varobs[o0 + i];
            nuse++;
        }
        if (nuse == 0) {
            iopt[e] = -1; objmin[e] = 0.0;
            hypo[4 * e] = hypo[4 * e + 1] = hypo[4 * e + 2] = hypo[4 * e + 3] = 0.0;
            continue;
        }
        for (int g = 0; g < ngrd; g++) { obj[g] = 0.0; t0[g] = job == 2 ? 0.0 : tori[e]; }
        if (job == 2) {
            for (int i = 0; i < nobs; i++) {
                if (luseObs[o0 + i] == 0) continue;
                int tid = 2 * (statPtr[o0 + i] - 1) + (pickType[o0 + i] - 1);
                const float *tt = tables + ldgrd * (size_t)tid;
                double to = tobs[o0 + i] - statCor[i];
                double w = (1.0 / varobs[o0 + i]) / xnorm;
                for (int g = 0; g < ngrd; g++) t0[g] = t0[g] + w * (to - (double)tt[g]);
            }
        }
        for (int i = 0; i < nobs; i++) {
            if (luseObs[o0 + i] == 0) continue;
            int tid = 2 * (statPtr[o0 + i] - 1) + (pickType[o0 + i] - 1);
            const float *tt = tables + ldgrd * (size_t)tid;
            double to = tobs[o0 + i] - statCor[i];
            double w = (1.0 / varobs[o0 + i]) * SQRT2I_D;
            for (int g = 0; g < ngrd; g++) {
                double res = w * (to - ((double)tt[g] + t0[g]));
                obj[g] = obj[g] + res * res;
            }
        }
        int im = oracle_minloc_f64(ngrd, obj); /* == MAXLOC(-obj), first occurrence */
        iopt[e] = im;
        objmin[e] = obj[im];
        hypo[4 * e] = (double)xlocs[im];
        hypo[4 * e + 1] = (double)ylocs[im];
        hypo[4 * e + 2] = (double)zlocs[im];
        hypo[4 * e + 3] = t0[im];
    }
done:
    free(t0); free(obj);
    return rc;
}

/* Per-event posterior volume of the catalogue flavour: logPDF[g] = -objective[g] (locate.f90:458-461)
 * and t0[g], one event, picks given as (table id < 0 = unused, corrected pick time, variance).      */
int oracle_event_logpdf(int job, int ngrd, size_t ldgrd, const float *tables, int npicks, const int *table_id,
                        const double *tobs_cor, const double *varobs, double tori, double *logpdf, double *t0)
{
    if (job != 1 && job != 2) return 1;
    double xnorm = 0.0;
    for (int i = 0; i < npicks; i++)
        if (table_id[i] >= 0) xnorm = xnorm + 1.0 / varobs[i];
    for (int g = 0; g < ngrd; g++) { logpdf[g] = 0.0; t0[g] = job == 2 ? 0.0 : tori; }
    if (job == 2)
        for (int i = 0; i < npicks; i++) {
            if (table_id[i] < 0) continue;
            const float *tt = tables + ldgrd * (size_t)table_id[i];
            double w = (1.0 / varobs[i]) / xnorm, to = tobs_cor[i];
            for (int g = 0; g < ngrd; g++) t0[g] = t0[g] + w * (to - (double)tt[g]);
        }
    for (int i = 0; i < npicks; i++) {
        if (table_id[i] < 0) continue;
        const float *tt = tables + ldgrd * (size_t)table_id[i];
        double w = (1.0 / varobs[i]) * SQRT2I_D, to = tobs_cor[i];
        for (int g = 0; g < ngrd; g++) {
            double res = w * (to - ((double)tt[g] + t0[g]));
            logpdf[g] = logpdf[g] + res * res;
        }
    }
    for (int g = 0; g < ngrd; g++) logpdf[g] = -logpdf[g];
    return 0;
}

/* ---- L1 flavour: locate.c:1205-1335.  PARITY UNPINNED BY THE REFERENCE: weightedMedian__double is declared at
 * locate.c:73 and defined nowhere, so the reference cannot link or run this function; only its demo inputs
 * survive (locate.c:228-237).  The arithmetic restated here is the reference's; the median definition is the one
 * written in include/mceik_b200.h, implemented independently of the product code (index sort, not pair insertion). */
static const double *g_wm_x;
static int wm_cmp(const void *a, const void *b)
{
    int ia = *(const int *)a, ib = *(const int *)b;
    if (g_wm_x[ia] < g_wm_x[ib]) return -1;
    if (g_wm_x[ia] > g_wm_x[ib]) return 1;
    return ia < ib ? -1 : (ia > ib ? 1 : 0);
}
double oracle_weighted_median(int n, const double *x, const double *w)
{
    if (n < 1) return 0.0;
    int *perm = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) perm[i] = i;
    g_wm_x = x;
    qsort(perm, (size_t)n, sizeof(int), wm_cmp);
    double W = 0.0, cum = 0.0, med = x[perm[n - 1]];
    for (int k = 0; k < n; k++) W = W + w[perm[k]];
    double half = 0.5 * W;
    for (int k = 0; k < n; k++) {
        cum = cum + w[perm[k]];
        if (cum > half) { med = x[perm[k]]; break; }
        if (cum == half) { med = k + 1 < n ? 0.5 * (x[perm[k]] + x[perm[k + 1]]) : x[perm[k]]; break; }
    }
    free(perm);
    return med;
}

int oracle_l1_gridsearch_f64(int ldgrd, int ngrd, int nobs, int iwantOT, double t0use, const int *mask,
                             const double *tobs, const double *varobs, const double *test, double *t0,
                             double *objfn)
{
    if (ldgrd < ngrd || nobs < 1 || !mask || !tobs || !varobs || !test || !t0 || !objfn) return 1;
    double *tc = (double *)malloc(sizeof(double) * (size_t)nobs);
    double *wt = (double *)malloc(sizeof(double) * (size_t)nobs);
    double *res = (double *)malloc(sizeof(double) * (size_t)nobs);
    int *ptr = (int *)malloc(sizeof(int) * (size_t)nobs);
    int nuse = 0;
    double wtsum = 0.0;
    for (int i = 0; i < nobs; i++) {           /* locate.c:1236-1246 */
        if (mask[i] != 0) continue;
        wt[nuse] = 1.0 / varobs[i];
        tc[nuse] = tobs[i];
        wtsum = wtsum + wt[nuse];
        ptr[nuse++] = i;
    }
    for (int g = 0; g < ngrd; g++) objfn[g] = 0.0;
    if (iwantOT == 1) {
        if (fabs(wtsum - 1.0) > 1.e-14) {      /* locate.c:1263-1273 */
            double wtsumi = 1.0 / wtsum;
            for (int j = 0; j < nuse; j++) wt[j] = wt[j] * wtsumi;
        }
        for (int g = 0; g < ngrd; g++) {       /* locate.c:1284-1300 */
            for (int j = 0; j < nuse; j++) res[j] = tc[j] - test[(size_t)ldgrd * (size_t)ptr[j] + g];
            t0[g] = oracle_weighted_median(nuse, res, wt);
        }
    } else {
        for (int g = 0; g < ngrd; g++) t0[g] = t0use;
    }
    for (int j = 0; j < nuse; j++) {           /* locate.c:1312-1324 */
        const double *tt = test + (size_t)ldgrd * (size_t)ptr[j];
        for (int g = 0; g < ngrd; g++) objfn[g] = objfn[g] + wt[j] * fabs(tc[j] - tt[g] - t0[g]);
    }
    free(tc); free(wt); free(res); free(ptr);
    return 0;
}

/* ---- analytic homogeneous tables: homog.c:594-621 -------------------------------------
 * t[iz*nx*ny+iy*nx+ix] = sqrt((xs-x)^2+(ys-y)^2+(zs-z)^2) * (1/vel), x = x0 + ix*dx     */
int oracle_homogeneous_traveltimes(int nx, int ny, int nz, double x0, double y0, double z0,
                                   double dx, double dy, double dz, double xs, double ys,
                                   double zs, double vel, double *ttimes)
{
    const double slow = 1.0 / vel;
    const size_t nxy = (size_t)nx * (size_t)ny;
    for (int iz = 0; iz < nz; iz++)
        for (int iy = 0; iy < ny; iy++)
            for (int ix = 0; ix < nx; ix++) {
                double x = x0 + (double)ix * dx, y = y0 + (double)iy * dy,
                       z = z0 + (double)iz * dz;
                double ex = xs - x, ey = ys - y, ez = zs - z;
                double dist = sqrt(ex * ex + ey * ey + ez * ez);
                ttimes[(size_t)iz * nxy + (size_t)iy * nx + ix] = dist * slow;
            }
    return 0;
}
