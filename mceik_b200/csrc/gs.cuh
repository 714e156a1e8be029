// gs.cuh -- L2 travel-time-table grid search kernels: host-side launch interface.
//
// Reference semantics: locate_l2_gridSearch__double64/float64 locate.c:923-1203 with the
// stack kernels locate.c:388-567 and locate_minLoc* locate.c:811-851; catalogue contract of
// LOCATE3D_GRIDSEARCH locate.f90:385-499 (fp32 tables promoted to fp64, first-index optimum).
#pragma once
#include <cstdint>
#include "common.cuh"

namespace mceik {
namespace gs {

constexpr int kEventsPerBlock = 8;   // events sharing one pass over the tables
// pick slots of a block the fast (uniform) search kernel holds: 260 B of shared memory per slot, two CTAs per SM
// (100 KB each); blocks with more slots are searched by the general kernel
constexpr int kUniformMaxPicks = 392;
constexpr int kPointsPerThread = 2;  // grid points per thread
constexpr int kThreads = 256;
constexpr int kChunk = kThreads * kPointsPerThread;

// One partial optimum: best grid point of one event inside the chunks one CTA visited.
struct Partial {
    double val;   // objective at idx
    double t0;    // origin time at idx
    int idx;      // flat node, INT_MAX when nothing finite was seen
    int nan0;     // 1 when the objective at node 0 is NaN (sticky in the reference's scan)
};

struct LocateArgs {
    int job;           // 1: t0 = tori[e]; 2: analytic t0
    int nevents;
    int ngrd;
    size_t ldgrd;
    int maxpicks;      // max picks of any event (smem is sized from it)
    const float *tables;     // [ntables][ldgrd]
    const int *obs_ptr;      // [nevents+1]
    const int *table_id;     // [npicks], < 0 = unused pick
    const double *tobs_cor;  // [npicks]
    const double *w_t0;      // [npicks] (1/var)/sum(1/var)
    const double *w_obj;     // [npicks] (1/var)*sqrt2i
    const double *tori;      // [nevents] (job 1)
    const int *blk_uniform;  // [ceil(nevents/8)] 1 = every pick slot of the block uses one table (fast kernel)
    Partial *partials;       // [nevents][nlanes]
    int nlanes;              // gridDim.x of the main kernel
};

// Per-event weights: xnorm = sum 1/var over used picks in catalogue order (locate.c:981-1013),
// w_t0 = (1/var)/xnorm (locate.c:399), w_obj = (1/var)*0.7071067811865475 (locate.c:500).
void launch_prepare(int nevents, const int *d_obs_ptr, const int *d_table_id, const double *d_varobs,
                    double *d_w_t0, double *d_w_obj, int *d_nuse, cudaStream_t st);
// blk_uniform[b] = 1 when the 8 events of block b use the same table in every pick slot
void launch_classify(int nevents, const int *d_obs_ptr, const int *d_table_id, int *d_blk_uniform, cudaStream_t st);
size_t locate_smem_bytes(int maxpicks);
int locate_lanes(int nevents, int ngrd);
void launch_locate(const LocateArgs &a, cudaStream_t st);
void launch_finalize(int nevents, int nlanes, const Partial *d_partials, const int *d_nuse, int *d_iopt,
                     double *d_t0opt, double *d_objopt, cudaStream_t st);

// misfit of nmodels proposals against a catalogue located at fixed nodes (config 5), out[nmodels]
void launch_catalog_misfit(const float *d_tables, size_t ldgrd, int nmodels, int ntab, int nevents, const int *d_node,
                           const double *d_tobs, const double *d_var, const int *d_use, double *d_out, cudaStream_t st);

// Single-event, full-grid outputs (the locate.c / gridsearch.f90 contract): t0[g], objfn[g].
// ptr[j] is the row of used pick j in `test`; weights and corrected picks are already compressed.
template <typename T>
void launch_full_grid(int ngrd, size_t ldgrd, int nuse, const int *d_row, const T *d_tobs, const T *d_w_t0,
                      const T *d_w_obj, int want_ot, T t0use, const T *d_test, T *d_t0, T *d_obj,
                      cudaStream_t st);
// one event, resident fp32 tables: logPDF = -objective (fp64 and/or fp32) and the origin-time grid
void launch_event_grid(int ngrd, size_t ldgrd, int npicks, const int *d_table_id, const double *d_tobs, const double *d_w_t0,
                       const double *d_w_obj, int want_ot, double t0use, const float *d_tables, double *d_logpdf,
                       float *d_logpdf4, double *d_t0, cudaStream_t st);
// L1 flavour of the single-event search (locate.c:1205-1335): weighted-median origin time + weighted L1 misfit.
// d_test holds the nuse used rows packed in pick order; d_wt the (normalised) weights.
constexpr int kL1MaxObs = 128;
void launch_l1_grid(int ngrd, size_t ldgrd, int nuse, const double *d_tobs, const double *d_wt, int want_ot, double t0use,
                    const double *d_test, double *d_t0, double *d_obj, cudaStream_t st);
// first index of the strict minimum (locate.c:811-851); result written to d_out[0]
template <typename T>
void launch_minloc(int n, const T *d_x, int *d_out, void *d_scratch, size_t scratch_bytes, cudaStream_t st,
                   bool maxloc = false);
// deterministic grid sum (scratch: minloc_scratch_bytes()) and in-place scaling (LOCATE_NORMALIZE_PDF, locate.f90:43-64)
void launch_sum(int n, const double *d_x, double *d_out, void *d_scratch, size_t scratch_bytes, cudaStream_t st);
void launch_scale(int n, double factor, double *d_x, cudaStream_t st);
size_t minloc_scratch_bytes();

}  // namespace gs
}  // namespace mceik
