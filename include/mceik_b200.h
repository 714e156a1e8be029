/*
 * mceik_b200.h -- C ABI of libmceik_b200.so: the B200 (sm_100a) implementation of mceik's
 * forward-model / location hot path (fast-sweeping eikonal solver + L2 travel-time-table
 * grid search).  Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Two groups of entry points:
 *   (1) DROP-IN symbols: the names, argument order, by-reference/by-value convention and
 *       ierr behaviour of the reference's own C / Fortran BIND(C) interface for this path.
 *       Each declaration cites the reference interface it replaces (file:line under the
 *       reference tree).  An unmodified module.F90 / homog.c links against them.
 *   (2) BATCHED extensions (mceik_*): many fields / many events per call, host- or
 *       device-resident buffers.  The drop-in symbols are batch-of-one wrappers over these.
 *
 * There is NO CPU fallback: every compute entry point returns ierr/-rc != 0 (and
 * mceik_last_error() says why) when no CUDA device is usable.
 *
 * Grid convention (fsm3d.f90:1936-1938, homog.c:585-587): x fastest,
 *   flat (0-based) index = iz*nx*ny + iy*nx + ix ; z is up, z=0 is the model base.
 */
#ifndef MCEIK_B200_H
#define MCEIK_B200_H 1

#include <stdbool.h>
#include <stddef.h>
#include "mceik_b200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ===================================================================================== */
/* (1) DROP-IN ENTRY POINTS                                                              */
/* ===================================================================================== */

/* --- eikonal, serial life-cycle.  Replaces EIKONAL3D_SERIAL_DRIVER, declared at
 *     module.F90:369-382, defined at fsm3d.f90:1968-2052.  job 1 = initialise for (nx,ny,nz),
 *     job 2 = boundary conditions + fast sweeping, job 3 = finalise.  ierr: 0 ok; 1 on double
 *     init, solve-before-init, a source whose stencil leaves the grid (fsm3d.f90:716-755), or
 *     any CUDA failure.  Non-convergence after maxit iterations is NOT an error.             */
void eikonal3d_serial_driver(const int *job, const int *iverb, const int *maxit, const int *nsrc,
                             const int *nx, const int *ny, const int *nz,
                             const double *tol, const double *h,
                             const double *x0, const double *y0, const double *z0,
                             const double *ts, const double *xs, const double *ys, const double *zs,
                             const double *slow, double *u, int *ierr);

/* --- eikonal, "distributed" life-cycle.  Replaces EIKONAL3D_INITIALIZE (module.F90:344-356,
 *     fsm3d.f90:1583-1674), EIKONAL3D_SOLVE (module.F90:358-367, fsm3d.f90:1754-1840) and
 *     EIKONAL3D_FINALIZE (module.F90:384-389, fsm3d.f90:1891-1929).  `comm` (a Fortran MPI
 *     handle in the reference) is accepted and ignored: a field is never split across GPUs
 *     here, so ndivx/ndivy/ndivz/noverlap are recorded but unused and the solve has the
 *     SERIAL sweep semantics (the fixed point the reference's block-Jacobi MPI variant
 *     converges to).  `n` must equal nx*ny*nz on the calling process.                        */
void eikonal3d_initialize(const int *comm, const int *iverb, const int *nx, const int *ny, const int *nz,
                          const int *ndivx, const int *ndivy, const int *ndivz,
                          const int *noverlap, const int *maxit,
                          const double *x0, const double *y0, const double *z0,
                          const double *h, const double *tol, int *ierr);
void eikonal3d_solve(const int *comm, const int *nsrc, const int *n,
                     const double *ts, const double *xs, const double *ys, const double *zs,
                     const double *slow, double *u, int *ierr);
void eikonal3d_finalize(const int *comm, int *ierr);

/* --- L2 grid search, C flavour.  Replaces locate_l2_gridSearch__double64 (locate.c:923-1047)
 *     and locate_l2_gridSearch__float64 (locate.c:1079-1203): same argument checks (ldgrd
 *     bytes % 64, ldgrd >= ngrd, nobs >= 1, NULLs, 64-byte alignment of t0/test/objfn), same
 *     return code (0 ok, 1 error).  mask[i]==0 means "use pick i"; tcorr may be NULL.        */
int locate_l2_gridSearch__double64(const int ldgrd, const int ngrd, const int nobs, const int iwantOT,
                                   const double t0use, const int *mask, const double *tobs,
                                   const double *tcorr, const double *varobs, const double *test,
                                   double *t0, double *objfn);
int locate_l2_gridSearch__float64(const int ldgrd, const int ngrd, const int nobs, const int iwantOT,
                                  const float t0use, const int *mask, const float *tobs,
                                  const float *tcorr, const float *varobs, const float *test,
                                  float *t0, float *objfn);
/* L1 flavour of the single-event search -- replaces locate_l1_gridSearch__double64 (locate.c:1205-1335; SURVEY.md
 * section 8f row 3).  PARITY UNPINNED: the reference cannot link this function (its weighted median is declared at
 * locate.c:73 and defined nowhere), so there is no reference behaviour to compare with; the arithmetic below is
 * the reference's own, the median definition is ours, and the oracle restates both.
 *   used picks: mask[i] == 0, catalogue order; w_i = 1/varobs[i]; when iwantOT == 1 and |sum w - 1| > 1e-14 the
 *   weights are scaled by 1/sum w (locate.c:1263-1273; the reference's running sum indexes the packed array with
 *   the unpacked index, :1245 -- with masked picks that is a defect and is not reproduced).
 *   iwantOT == 1: t0[g] = weighted median of r_i = tobs[i] - test[i*ldgrd+g]; else t0[g] = t0use.
 *   objfn[g] = sum_i w_i * |(tobs[i] - test[i*ldgrd+g]) - t0[g]|, accumulated in pick order (locate.c:1320-1321).
 * Weighted median of (x_k, w_k): sort by (x, original index) ascending; W = sum of w in that order; the first k whose
 * running sum exceeds W/2 gives x_k; a running sum exactly equal to W/2 gives (x_k + x_{k+1})/2 (the ordinary
 * median for equal weights).  At most 128 used picks.  Returns 0, or 1 on bad arguments. */
int locate_l1_gridSearch__double64(int ldgrd, int ngrd, int nobs, int iwantOT, double t0use, const int *mask,
                                   const double *tobs, const double *varobs, const double *test, double *t0,
                                   double *objfn);
/* The weighted median above as a host function with the prototype the reference declares (locate.c:73-77): perm
 * (may be NULL) is an ordering the caller keeps between calls, *lsort tells whether it had to be rebuilt. */
double weightedMedian__double(int n, const double *x, const double *w, int *perm, bool *lsort, int *ierr);
/* first index of the strict minimum (locate.c:811-851) */
int locate_minLocDouble64(const int n, const double *x);
int locate_minLocFloat64(const int n, const float *x);

/* --- L2 grid search, Fortran flavour.  Replaces LOCATE3D_GRIDSEARCH_DOUBLE64 /
 *     LOCATE3D_GRIDSEARCH_FLOAT64 (interface gridsearch.f90:2-26, bodies :382-459, :463-540):
 *     by-reference arguments, mask==1 skips, t0 weight 1/(var_i * sum var), result in logPDF
 *     (a positive misfit despite its name).  ierr = 1 when ldgrd % 64 != 0 (in ELEMENTS),
 *     ngrd > ldgrd, every pick masked, or |sum varobs| < epsilon.                            */
void locate3d_gridsearch__double64(const int *ldgrd, const int *ngrd, const int *nobs, const int *iwantOT,
                                   const int *mask, const double *tobs, const double *varobs,
                                   const double *test, double *logPDF, int *ierr);
void locate3d_gridsearch__float64(const int *ldgrd, const int *ngrd, const int *nobs, const int *iwantOT,
                                  const int *mask, const float *tobs, const float *varobs,
                                  const float *test, float *logPDF, int *ierr);

/* --- catalogue locator.  Replaces LOCATE3D_INITIALIZE / LOCATE3D_GRIDSEARCH /
 *     LOCATE3D_FINALIZE (include/locate.h:10-23, locate.f90:322-519, :562-689).
 *     The reference pulls every table from HDF5 inside its event loop (locate.f90:400,443);
 *     HDF5 stays host-side, so here the tables and node coordinates are handed over once with
 *     mceik_locate_set_tables() / mceik_locate_set_grid() (by the host's h5io reader) between
 *     locate3d_initialize and locate3d_gridsearch; tttFileID/locFileID are recorded only.
 *     job 1 = location with fixed origin time tori, job 2 = location + analytic origin time,
 *     other jobs -> ierr = 1 ("Not yet done", locate.f90:502-512).  hypo = (x,y,z,t0) per
 *     event.  `test` (declared OUT, never written by the reference) is left untouched.       */
void locate3d_initialize(const int *comm, const int *iverb, const long *tttFileID, const long *locFileID,
                         const int *ndivx, const int *ndivy, const int *ndivz, int *ierr);
void locate3d_gridsearch(const int *model, const int *job, const int *nobs, const int *nevents,
                         const int *luseObs, const int *statPtr, const int *pickType,
                         const double *statCor, const double *tori, const double *varobs,
                         const double *tobs, double *test, double *hypo, int *ierr);
void locate3d_finalize(void);

/* --- analytic homogeneous table.  Replaces computeHomogeneousTraveltimes (homog.c:594-621). */
int computeHomogeneousTraveltimes(const int nx, const int ny, const int nz,
                                  double x0, double y0, double z0,
                                  const double dx, double dy, const double dz,
                                  const double xs, const double ys, double zs,
                                  const double vel, double *ttimes);

/* ===================================================================================== */
/* (2) BATCHED EXTENSIONS                                                                */
/* ===================================================================================== */

typedef struct mceik_ctx mceik_ctx; /* opaque: device id, stream, workspaces, resident tables */

/* Last error text of the calling thread ("" if none). */
const char *mceik_last_error(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
long long mceik_kernel_launch_count(void);

/* device < 0 selects the current CUDA device.  stream is a cudaStream_t passed as void*
 * (NULL = a private non-blocking stream owned by the context). */
int mceik_ctx_create(int device, void *stream, mceik_ctx **ctx);
void mceik_ctx_destroy(mceik_ctx *ctx);
int mceik_ctx_synchronize(mceik_ctx *ctx);

/* Geometry + solver parameters of one batch of eikonal fields (solverParametersType,
 * module.F90:117-131, minus the MPI decomposition). */
typedef struct mceik_fsm_grid {
    int nx, ny, nz;
    double h;          /* isotropic spacing (m); the solver has a single h (fsm3d.f90:2021) */
    double x0, y0, z0; /* origin (m) */
    double tol;        /* convergence tolerance (s) */
    int maxit;         /* iteration cap; one iteration = 8 sweeps */
} mceik_fsm_grid;

enum { MCEIK_FSM_ALGO_TILES = 0, MCEIK_FSM_ALGO_LEVELS = 1, MCEIK_FSM_ALGO_BRICKS = 2 };

/*
 * Solve `nfields` independent eikonal fields in one call.
 *   slow        [nmodels][N] fp64 slowness (s/m), N = nx*ny*nz
 *   field_model [nfields]    which slowness model field f uses
 *   src_ptr     [nfields+1]  CSR into ts/xs/ys/zs (HOST arrays in both variants): the sources
 *                            (normally one station) that seed field f
 *   u           [nfields][N] fp64 travel times out (may be NULL if only tables are wanted)
 *   tables      [nfields][ldtab] fp32 travel-time tables out (may be NULL), ldtab >= N
 *   iters       [nfields]    iterations executed per field (host, may be NULL)
 *   field_ierr  [nfields]    per-field status: 0 ok, 1 = source stencil out of grid (host, may be NULL)
 * Returns 0 when every field solved, 1 when some field failed (field_ierr says which), <0 on
 * argument / CUDA errors.  _host takes host pointers and stages through the context's device
 * workspace; _dev takes device pointers for slow/u/tables and runs on the context stream
 * without synchronising the caller except for the small per-iteration convergence read-back.
 */
int mceik_fsm_solve_batched_host(mceik_ctx *ctx, const mceik_fsm_grid *grid, int nmodels, const double *slow,
                                 int nfields, const int *field_model, const int *src_ptr,
                                 const double *ts, const double *xs, const double *ys, const double *zs,
                                 double *u, float *tables, size_t ldtab, int *iters, int *field_ierr);
int mceik_fsm_solve_batched_dev(mceik_ctx *ctx, const mceik_fsm_grid *grid, int nmodels, const double *d_slow,
                                int nfields, const int *field_model, const int *src_ptr,
                                const double *ts, const double *xs, const double *ys, const double *zs,
                                double *d_u, float *d_tables, size_t ldtab, int *iters, int *field_ierr);
/* Select the sweep kernel: BRICKS (default; warp-per-brick streaming kernel, fsm_bricks.cu), TILES
 * (CTA-per-16^3-tile kernel, fsm.cu) or LEVELS (one launch per hyperplane; cross-check path). */
int mceik_fsm_set_algo(mceik_ctx *ctx, int algo);
/* Development switches of the sweep / search kernels (none is needed in production: the defaults are the measured
 * best).  Keys: "ZC", "BY", "NO16", "PUBLISH", "PUBLISHER", "NO_STAGGER", "NATURAL", "L2PF", "BATCH", "TRACE",
 * "STATS", "DEBUG", "LOCATE_NO_ALIGN" (DESIGN.md section 5).  A context reads MCEIK_FSM_<KEY> / MCEIK_LOCATE_NO_ALIGN
 * from the environment once, when it is created.  Returns 0, or -1 for an unknown key. */
int mceik_fsm_set_tuning(mceik_ctx *ctx, const char *key, int value);
/* Node-updates executed by the last solve on this context (N * 8 * iterations, summed over fields). */
long long mceik_fsm_last_node_updates(mceik_ctx *ctx);
/* Device time (ms, CUDA events on the context stream) and count of the sweep-kernel launches of the
 * last solve: the numerator/denominator of the roofline figure bench.py reports. */
int mceik_fsm_last_sweep_stats(mceik_ctx *ctx, double *sweep_ms, int *launches);

/* Device self-test of the sweep kernels' arithmetic building blocks: the branch-free square root
 * and the straight-line local solver must equal __dsqrt_rn and the reference-ordered solver bit
 * for bit on `samples` pseudo-random inputs.  Outputs the number of mismatches (expected 0). */
int mceik_selftest_solver(mceik_ctx *ctx, unsigned long long seed, long long samples, long long *bad_sqrt,
                          long long *bad_solve);

/*
 * Multi-GPU (one process per GPU): the fields of a batched solve are shared out over the ranks of a communicator,
 * every rank solves its fields without any communication, and one in-place NCCL all-gather over NVLink replicates
 * the fp32 tables.  This replaces the reference's decomposition of ONE field over the ranks of an MPI communicator
 * with a ghost exchange per sweep (eikonal3d_initialize / eikonal3d_solve, fsm3d.f90:1583-1840, communicators from
 * mpiutils.f90:99-264): as there, the host hands in a communicator and nothing else.
 *   mceik_comm_unique_id   rank 0 obtains the 128-byte NCCL id and distributes it with the host's own means
 *                          (MPI_Bcast in the reference's drivers, torch.distributed in bench.py);
 *   mceik_comm_init        every rank: joins (collective); mceik_comm_destroy leaves.
 *   mceik_fsm_assign_fields  the deterministic assignment every rank computes: fields of a slowness model are dealt
 *                          over the ranks holding that model, heaviest first when `cost` (e.g. the iteration counts of
 *                          the previous solve of an MCMC loop; NULL = equal) is given.  table_row[f] = rank * slots + k.
 *   mceik_fsm_solve_sharded_dev  collective.  All arguments describe ALL nfields fields and are identical on every
 *                          rank (d_slow: every model resident on every rank).  d_tables_all [world * slots][ldtab]
 *                          receives every field's fp32 table at row table_row[f] on every rank; iters / field_ierr
 *                          [nfields] are filled on every rank.  Returns 1 when any field failed its boundary conditions.
 *                          With more ranks than fields some ranks solve nothing and still receive every table.  A
 *                          rank whose own solve fails (< 0, mceik_last_error) still takes part in the closing
 *                          collectives, so the others return (1, with field_ierr = -1 for that rank's fields)
 *                          instead of waiting for it.
 *   mceik_tables_allgather the collective alone (in place: rank r's rows are [r * slots, (r + 1) * slots)).
 *   mceik_tables_alloc_replicated  collective: one buffer of rows x ldtab floats per rank, every rank mapping the
 *                          buffers of all the others (CUDA IPC).  When d_tables_all of mceik_fsm_solve_sharded_dev is this
 *                          buffer, every table is PUT into the peers' copies over NVLink by the copy engines as soon as
 *                          its field has converged, under the sweeps of the remaining fields: no collective kernel and
 *                          no rank waiting for another during the solve (the closing all-gather of the iteration counts
 *                          is the barrier).  The buffer holds a complete set of tables from the return of the solve
 *                          until any rank starts the next sharded solve into it (the host separates the two with its
 *                          own barrier when the tables are still in use).  mceik_tables_free_replicated releases it (collective in effect: no rank may
 *                          still be putting).
 * NCCL is loaded at run time (libnccl.so.2); without it these entry points return -2 and the rest of the library works.
 */
int mceik_comm_unique_id(void *id128);
int mceik_comm_init(mceik_ctx *ctx, int world, int rank, const void *id128);
int mceik_comm_destroy(mceik_ctx *ctx);
int mceik_fsm_assign_fields(int nfields, const int *field_model, const int *cost, int world, int *rank_of_field,
                            int *table_row, int *slots);
int mceik_fsm_solve_sharded_dev(mceik_ctx *ctx, const mceik_fsm_grid *grid, int nmodels, const double *d_slow, int nfields,
                                const int *field_model, const int *src_ptr, const double *ts, const double *xs,
                                const double *ys, const double *zs, const int *cost, float *d_tables_all, size_t ldtab,
                                int *iters, int *field_ierr, int *table_row);
int mceik_tables_allgather(mceik_ctx *ctx, float *d_tables_all, size_t ldtab, int slots);
int mceik_tables_alloc_replicated(mceik_ctx *ctx, size_t rows, size_t ldtab, float **d_tables_all);
int mceik_tables_free_replicated(mceik_ctx *ctx);

/* Forward-loop misfit of many proposals (BASELINE config 5; the reference has no code for it, the definition is the
 * build's): model m owns tables [m * ntab, (m + 1) * ntab) of d_tables [nmodels * ntab][ldgrd]; every
 * event e sits at its catalogue node d_node[e] (0-based, < ngrd) with picks d_tobs / d_varobs / d_use [nevents][ntab]
 * (use != 0: the pick counts); misfit[m] = sum_e sum_j (w_j/sqrt2 (tobs_j - (T_j + t0_e)))^2 with the analytic
 * weighted-mean origin time t0_e of locate.c:399-410.  Device pointers, asynchronous on the context stream. */
int mceik_catalog_misfit_dev(mceik_ctx *ctx, const float *d_tables, size_t ldgrd, int ngrd, int nmodels, int ntab, int nevents,
                             const int *d_node, const double *d_tobs, const double *d_varobs, const int *d_use,
                             double *d_misfit);

/* Analytic homogeneous tables on the device: fp32 table t = dist/vel per station (homog.c:594-621
 * followed by homog.c:624-635).  d_tables [nstations][ldtab]; xs,ys,zs,vel are host arrays. */
int mceik_homogeneous_tables_dev(mceik_ctx *ctx, int nx, int ny, int nz, double x0, double y0, double z0,
                                 double dx, double dy, double dz, int nstations, const double *xs,
                                 const double *ys, const double *zs, const double *vel,
                                 float *d_tables, size_t ldtab);

/* Travel-time tables resident for the locator: fp32 [ntables][ldgrd], table id =
 * 2*(station-1) + (pickType-1) for the catalogue entry points.  _host copies, _dev borrows the
 * caller's device buffer (it must outlive the context or the next set call). */
int mceik_locate_set_tables_host(mceik_ctx *ctx, int ntables, int ngrd, size_t ldgrd, const float *tables);
int mceik_locate_set_tables_dev(mceik_ctx *ctx, int ntables, int ngrd, size_t ldgrd, const float *d_tables);
/* Node coordinates (fp32, as read from /Model/{x,y,z}locs, locate.f90:646-652), host arrays [ngrd]. */
int mceik_locate_set_grid(mceik_ctx *ctx, int ngrd, const float *xlocs, const float *ylocs, const float *zlocs);
/* Same for the process-global context used by the drop-in locate3d_* symbols. */
int mceik_locate3d_set_tables(int ntables, int ngrd, size_t ldgrd, const float *tables);
int mceik_locate3d_set_grid(int ngrd, const float *xlocs, const float *ylocs, const float *zlocs);

/*
 * Locate a batch of events against the resident tables (CSR picks).
 *   obs_ptr  [nevents+1]; per pick: table id (<0 = unused), corrected pick time
 *   (tobs - static correction), variance.  job 1: t0 fixed to tori[e]; job 2: analytic t0.
 *   Outputs per event: iopt (0-based flat node, -1 when the event has no usable pick),
 *   t0opt (origin time at iopt), objopt (objective at iopt).
 * _host: all pointers host.  _dev: all pointers device, asynchronous on the context stream;
 *   nobs_total = obs_ptr[nevents], max_picks = an upper bound on picks per event (sizes smem).
 */
int mceik_locate_batched_host(mceik_ctx *ctx, int job, int nevents, const int *obs_ptr, const int *table_id,
                              const double *tobs_cor, const double *varobs, const double *tori,
                              int *iopt, double *t0opt, double *objopt);
int mceik_locate_batched_dev(mceik_ctx *ctx, int job, int nevents, int nobs_total, int max_picks, const int *d_obs_ptr,
                             const int *d_table_id, const double *d_tobs_cor, const double *d_varobs,
                             const double *d_tori, int *d_iopt, double *d_t0opt, double *d_objopt);
/* Full posterior volume of ONE event against the resident tables: the
 * per-event logPDF the reference accumulates, logPDF[g] = -sum_i (w_i/sqrt2 * (tobs_i - (T_i[g] + t0[g])))^2
 * (locate.f90:436-463, fp32 tables promoted to fp64), optionally its fp32 copy (what the location file
 * stores, h5io.c:714-819) and the origin-time grid t0[g].  Picks as in mceik_locate_batched_host
 * (table_id < 0 = unused); job 1 uses t0 = tori everywhere.  Host output arrays [ngrd]; any may be NULL. */
int mceik_locate_event_logpdf_host(mceik_ctx *ctx, int job, int npicks, const int *table_id, const double *tobs_cor,
                                   const double *varobs, double tori, double *logpdf, float *logpdf4, double *t0grid);
/* LOCATE_OPTNODE (locate.f90:75-117) on one device: 0-based first index of the maximum of pdf[ngrd] (MAXLOC). */
int mceik_locate_optnode_host(mceik_ctx *ctx, int ngrd, const double *pdf, int *node);
/* LOCATE_NORMALIZE_PDF (locate.f90:43-64): pdf *= 1/sum(pdf) in place; *sum (may be NULL) receives the sum.
 * Returns 1 and leaves pdf untouched when the sum is exactly zero.  The sum uses a fixed reduction tree
 * (bit-reproducible run to run); its order differs from the Fortran SUM, so it agrees with a sequential
 * sum to ~1e-13 relative, not bit for bit (tests use 1e-12). */
int mceik_locate_normalize_pdf_host(mceik_ctx *ctx, int ngrd, double *pdf, double *sum);
/* Catalogue form (mceik_struct.h layouts): picks are catalog->obsPtr CSR, table from
 * (statPtr, pickType), static correction from stations->pcorr/scorr; hypo[4*nevents] needs
 * mceik_locate_set_grid().  iopt/obj may be NULL. */
int mceik_locate_catalog(mceik_ctx *ctx, const struct mceik_catalog_struct *catalog,
                         const struct mceik_stations_struct *stations, int job,
                         double *hypo, int *iopt, double *obj);

#ifdef __cplusplus
}
#endif
#endif /* MCEIK_B200_H */
