// fsm.cuh -- batched fast-sweeping eikonal kernels: host-side launch interface.
//
// Reference semantics (serial path): EIKONAL3D_SETBCS fsm3d.f90:762-840, EIKONAL3D_FSM
// fsm3d.f90:28-99, EVAL_UPDATE3D/UPDATE3D fsm3d.f90:419-546, SOLVE_HAMILTONIAN3D :648-693.
#pragma once
#include <cstdint>
#include <vector>
#include "common.cuh"

namespace mceik {
namespace fsm {

constexpr int kTile = 16;                     // tile edge (nodes)
constexpr int kTileNodes = kTile * kTile * kTile;
constexpr int kHalo = kTile + 2;              // smem row stride of a tile with its 1-node halo
constexpr int kHaloNodes = kHalo * kHalo * kHalo;
constexpr int kTileLevels = 3 * kTile - 2;    // hyperplanes inside one tile
constexpr int kMaxSlots = 4;                  // fields (sharing one slowness model) per CTA
constexpr int kGroupThreads = 128;            // threads working on one slot

// One boundary-condition record: a stencil node of one source (fsm3d.f90:810-833).
struct BcRecord {
    double d;       // distance source -> node (m), computed on the host exactly as :823
    double ts;      // source time (s)
    int node;       // flat 0-based node index
    int collocated; // |d| < 1e-10 -> assign instead of min (:825-829)
};

// Static description of the tile decomposition of an (nx,ny,nz) grid.
struct TilePlan {
    int nx = 0, ny = 0, nz = 0;
    int ntx = 0, nty = 0, ntz = 0, ntiles = 0, ntlevels = 0;
    DevBuf tile_order;  // int[ntiles]: I | J<<10 | K<<20, sorted by I+J+K
    DevBuf tlevel_ptr;  // int[ntlevels+1]
    DevBuf lvl_nodes;   // uint16[kTileNodes]: a | b<<4 | c<<8 sorted by a+b+c, then c, then b
    std::vector<int> h_tlevel_ptr;
    void build(int nx_, int ny_, int nz_, cudaStream_t st);
    void release();
};

// Arguments of one launch of the tile-wavefront sweep kernel (= one FSM iteration = 8 sweeps).
struct SweepArgs {
    int nx, ny, nz;
    int ntx, nty, ntz, ntiles, ntlevels;
    int ngroups, nslots;
    double h, tol;
    const int *group_fields;  // [ngroups][kMaxSlots] field index or -1
    const int *group_model;   // [ngroups]
    const double *slow;       // [nmodels][N]
    double *u;                // [nfields][N]
    double *u0;               // [nfields][N] values at the start of the iteration
    const int *tile_order;
    const int *tlevel_ptr;
    const uint16_t *lvl_nodes;
    int *done;                  // [ngroups][ntiles] sweeps completed per tile, zeroed per launch
    int *queue;                 // [1] ticket counter, zeroed per launch
    unsigned long long *nonconv; // [nfields] nodes with !(|u0-u| < tol), zeroed per launch
    const int *bc_ptr;          // [nfields+1] CSR into bc_tile / bc_local
    const int *bc_tile;         // tile id of each (unique) boundary-condition node
    const uint16_t *bc_local;   // i | j<<4 | k<<8 inside that tile
};

// Brick decomposition (8 x 8 x zc nodes) used by the streaming sweep kernel (fsm_bricks.cu).
struct BrickPlan {
    int nx = 0, ny = 0, nz = 0, zc = 0, by = 0;  // brick = 8 x by x zc nodes (by = 8 or 16)
    int nbx = 0, nby = 0, nbz = 0, nbricks = 0, nblevels = 0;
    DevBuf brick_order;  // int[nbricks]: I | J<<10 | K<<20 sorted by I+J+K
    DevBuf blevel_ptr;   // int[nblevels+1]
    std::vector<int> h_blevel_ptr;  // host copy (ticket tables are built per launch)
    void build(int nx_, int ny_, int nz_, int by_, int zc_, cudaStream_t st);
    void release();
};

// Arguments of one launch of the brick sweep kernel (= 8 sweeps of every active field).
struct BrickArgs {
    int nx, ny, nz;
    int nbx, nby, nbz, nbricks, nblevels, zc, by;
    int nfields_active;
    int publish;              // a sweeping warp publishes its progress every `publish` steps (4, 8 or 16; a power of two)
    double h;
    const int *active;        // [nfields_active] field ids
    const int *field_model;   // [nfields]
    const double *slow;       // [nmodels][N]; bricks16: slow*h (slow_is_fh), the product UPDATE3D forms per visit
    int slow_is_fh;
    double *u;                // [nfields][N]
    const int *brick_order;
    const int *blevel_ptr;
    int *done;                    // [nfields][nbricks] sweeps completed per brick, zeroed per launch
    unsigned long long *queue;    // [1] ticket counter, zeroed per launch
    // bricks16 ticket order: the active fields form two groups; group 1 runs `stagger` brick levels behind group 0,
    // so one group is in the wide middle of a sweep while the other ramps up or drains.  Virtual level V holds
    // (sweep, level) = divmod(V, nblevels) of group 0 and divmod(V - stagger, nblevels) of group 1.
    const long long *vptr;        // [8 * nblevels + stagger + 1] tickets before virtual level V
    int nf0, stagger;             // group 0 = active[0, nf0), group 1 = active[nf0, nfields_active)
    // bricks16, blocked layout: u is [field][brick column J * nbx + I][z][80] -- the 8 x 8 nodes of one brick plane
    // (row-major, 512 contiguous bytes) followed by copies of its columns 0 and 7 (8 + 8 values), which the x
    // neighbours read as their halo instead of 8 separate sectors of u; slow is [model][brick column][z][64].  Rows
    // beyond ny hold u_nan.  launch_block_fields / launch_unblock_fields convert from / to the [z][y][x] layout.
    int blocked;
    int l2_prefetch;              // bricks16 experiment (MCEIK_FSM_L2PF): brick records bulk-prefetched into the L2 this many planes ahead
    int batch;                    // bricks16 experiment (MCEIK_FSM_BATCH): > 0 = fields run in sequential batches of this size
    int publisher;                // bricks16: 1 = the last warp of every CTA publishes progress for the others
    const int *bc_ptr;            // [nfields+1] CSR into bc_node
    const int *bc_node;           // flat node index of each (unique) boundary-condition node
    unsigned long long *stats;    // optional [4] cycle counters (MCEIK_FSM_STATS=1), else nullptr
    int debug;                    // bottleneck experiments only (MCEIK_FSM_DEBUG): 1 no solver, 2 no stores, 4 no loads
};
void launch_iteration_bricks(const BrickArgs &a, cudaStream_t st);

// Ticket of the bricks16 queue -> (sweep, brick level, index of the brick in its level, index into `active`).
// vptr[V] = tickets before virtual level V (built by host::build_ticket_table); shared by the kernel and the CPU test.
struct TicketTask {
    int sweep, level, bidx, fidx;
};
__host__ __device__ inline TicketTask decode_ticket(long long t, const long long *vptr, const int *blevel_ptr, int nblevels,
                                                    int stagger, int nf0, int nf1) {
    const int nvl = 8 * nblevels + stagger;
    int lo = 0, hi = nvl;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (vptr[mid] <= t) lo = mid; else hi = mid;
    }
    long long r = t - vptr[lo];
    int vl = lo, gnf = nf0, gofs = 0;
    if (lo < 8 * nblevels) {
        const int l0 = lo % nblevels;
        const long long cnt0 = (long long)nf0 * (blevel_ptr[l0 + 1] - blevel_ptr[l0]);
        if (r >= cnt0) { r -= cnt0; vl = lo - stagger; gnf = nf1; gofs = nf0; }
    } else {
        vl = lo - stagger; gnf = nf1; gofs = nf0;
    }
    TicketTask k;
    k.sweep = vl / nblevels;
    k.level = vl - k.sweep * nblevels;
    k.bidx = (int)(r / gnf);
    k.fidx = gofs + (int)(r - (long long)k.bidx * gnf);
    return k;
}
// 16-byte-pair variant (fsm_bricks16.cu): requires nx % 8 == 0 and by == 8
void launch_iteration_bricks16(const BrickArgs &a, cudaStream_t st);
// most planes of one brick that may hold boundary-condition nodes of one field (more: use launch_iteration_bricks)
int bricks16_max_bc_planes();

// device self-test of sqrt_fast / local_solve_sl against __dsqrt_rn / local_solve; d_bad[2] counts mismatches
void launch_selftest(unsigned long long seed, int blocks, int per_thread, unsigned long long *d_bad, cudaStream_t st);
void launch_fill(double *d_u, size_t n, double value, cudaStream_t st);
// d_out[i] = d_slow[i] * h (fsm3d.f90:470 evaluates this product at every node visit; same bits every time)
void launch_scale_slowness(size_t n, double h, const double *d_slow, double *d_out, cudaStream_t st);
// One thread per field applies its BcRecords in source order.
void launch_apply_bcs(int nfields, size_t n, const int *d_field_model, const int *d_rec_ptr,
                      const BcRecord *d_recs, const double *d_slow, double *d_u, cudaStream_t st);
void launch_iteration_tiles(const SweepArgs &a, cudaStream_t st);
size_t tiles_smem_bytes(int nslots);
// Debug / cross-check path: one launch per hyperplane, global memory only.
void launch_iteration_levels(int nx, int ny, int nz, double h, int nfields, const int *d_active_fields,
                             const int *d_field_model, const double *d_slow, const uint8_t *d_lupd,
                             double *d_u, cudaStream_t st);
// u0 = u and count !(|u0-u| < tol) per field (used by the levels path; fsm3d.f90:86-90).
void launch_convergence(size_t n, int nfields, const int *d_active_fields, double tol, const double *d_u,
                        double *d_u0, unsigned long long *d_nonconv, cudaStream_t st);
void launch_mark_bcs(int nrec, const int *d_rec_field, const int *d_rec_node, size_t n, uint8_t *d_lupd,
                     cudaStream_t st);
// Blocked layout of the bricks16 kernel (BrickArgs::blocked).  block: d_ub[f] = records of d_u[f] for every field
// (u_nan in rows beyond ny); unblock: nodes of d_ub[f] for the listed fields (d_fields == nullptr: all).
size_t blocked_field_doubles(int nx, int ny, int nz);
size_t blocked_slowness_doubles(int nx, int ny, int nz);
void launch_block_fields(int nx, int ny, int nz, int nfields, const double *d_u, double *d_ub, cudaStream_t st);
void launch_unblock_fields(int nx, int ny, int nz, int nfields, const int *d_fields, const double *d_ub, double *d_u,
                           float *d_tables, size_t ldtab, cudaStream_t st);  // fp64 field and / or fp32 table (either may be null)
void launch_apply_bcs_blocked(int nfields, int nx, int ny, int nz, const int *d_field_model, const int *d_rec_ptr,
                              const BcRecord *d_recs, const double *d_slow, double *d_ub, cudaStream_t st);
void launch_convergence_blocked(int nx, int ny, int nz, int nfields, const int *d_active_fields, double tol, const double *d_ub,
                                double *d_u0b, unsigned long long *d_nonconv, cudaStream_t st);
// d_out[model][brick column][z][64] = d_slow[model][z][y][x] * h (1.0 in rows beyond ny)
void launch_scale_slowness_blocked(int nx, int ny, int nz, int nmodels, double h, const double *d_slow, double *d_out,
                                   cudaStream_t st);
// fp64 field -> fp32 table (fsm3d.f90:1870-1872, homog.c:624-635)
void launch_pack_tables(int nfields, size_t n, size_t ldtab, const double *d_u, float *d_tables,
                        cudaStream_t st);
// analytic homogeneous field(s): OutT = float (locator tables) or double (homog.c:594-621 as is)
template <typename OutT>
void launch_homog_tables(int nx, int ny, int nz, double x0, double y0, double z0, double dx, double dy,
                         double dz, int nstations, const double *d_xyzv, OutT *d_tables, size_t ldtab,
                         cudaStream_t st);

}  // namespace fsm
}  // namespace mceik
