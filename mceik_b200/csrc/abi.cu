// abi.cu -- the C ABI of libmceik_b200.so (declared in include/mceik_b200.h).
//
// Host-side orchestration only: argument checks that mirror the reference entry points,
// staging between caller memory and HBM, the per-iteration convergence loop of the eikonal
// solver, and the batch-of-one wrappers behind the drop-in symbols.  All arithmetic on grid-
// sized data happens in the kernels of fsm.cu / gs.cu; there is no CPU fallback.
#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/mceik_b200.h"
#include "comm.cuh"
#include "common.cuh"
#include "fsm.cuh"
#include "gs.cuh"
#include "host_logic.hpp"

namespace mceik {

static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *last_error() { return g_err; }
static std::atomic<long long> g_launches{0};
void count_launch(long long n) { g_launches += n; }
long long launch_count() { return g_launches.load(); }

}  // namespace mceik

using namespace mceik;

struct mceik_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int fsm_algo = MCEIK_FSM_ALGO_BRICKS;
    long long last_updates = 0;
    // device time of the sweep kernel launches of the last solve (CUDA events on ctx->stream)
    double last_sweep_ms = 0.0;
    int last_sweep_launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fin = nullptr;
    // mceik_fsm_solve_batched_host with pinned output: a converged field is copied back on copy_stream while
    // the remaining fields keep iterating (early_u = host destination, early_done[f] = already on its way)
    comm::Comm *comm = nullptr;  // mceik_comm_init: the ranks sharing the sources (one process per GPU)
    // sharded solve into the replicated buffer (mceik_tables_alloc_replicated): the table of local field f lands in row
    // put_row0 + f and is put into the peers' buffers on put_stream as soon as the field has converged
    bool put_enabled = false;
    size_t put_row0 = 0, put_ldtab = 0, put_n = 0;
    cudaStream_t put_stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    double *early_u = nullptr;
    std::vector<char> early_done;
    fsm::TilePlan plan;
    fsm::BrickPlan bplan;
    // development switches of the sweep kernels: defaults are the measured best; read from the environment once when
    // the context is created (MCEIK_FSM_<KEY>), changed per context with mceik_fsm_set_tuning().  -1 = automatic.
    struct Tuning {
        int zc = 256;         // planes per brick (<= 256)
        int by = 8;           // 16: 8x16 cross-section, generic brick kernel
        int no16 = 0;         // 1: generic brick kernel even when nx % 8 == 0
        int publish = -1;     // steps between progress publications (power of two)
        int publisher = -1;   // 0 / 1: force one publication flavour
        int no_stagger = 0;   // 1: one field group in the ticket order
        int natural = 0;      // 1: bricks16 works on the caller's [z][y][x] layout instead of the blocked one
        int l2pf = 0;         // brick records bulk-prefetched into the L2 this many planes ahead (experiment; 0 = off)
        int batch = 0;        // sequential field batches (experiment)
        int trace = 0, stats = 0, debug = 0;
        int locate_no_align = 0;  // 1: keep ragged event blocks on the general search kernel (MCEIK_LOCATE_NO_ALIGN)
    } tune;
    // eikonal workspaces
    DevBuf ws_slow, ws_u, ws_u0, ws_tab, ws_meta, ws_ctrl, ws_lupd, ws_xyzv, ws_fh, ws_ub, ws_u0b, ws_comm;
    // locator state
    const float *d_tables = nullptr;
    DevBuf own_tables;
    int ntables = 0, ngrd = 0;
    size_t ldgrd = 0;
    std::vector<float> xlocs, ylocs, zlocs;
    DevBuf ws_gs_in, ws_gs_w, ws_gs_part, ws_gs_out, ws_gs_misc;
};

namespace {

int *tuning_slot(mceik_ctx *c, const char *key) {
    struct { const char *name; int *p; } const tab[] = {
        {"ZC", &c->tune.zc}, {"BY", &c->tune.by}, {"NO16", &c->tune.no16}, {"PUBLISH", &c->tune.publish},
        {"PUBLISHER", &c->tune.publisher}, {"NO_STAGGER", &c->tune.no_stagger}, 
        {"NATURAL", &c->tune.natural}, {"L2PF", &c->tune.l2pf}, {"BATCH", &c->tune.batch}, {"TRACE", &c->tune.trace},
        {"STATS", &c->tune.stats}, {"DEBUG", &c->tune.debug}, {"LOCATE_NO_ALIGN", &c->tune.locate_no_align}};
    for (const auto &e : tab)
        if (strcmp(e.name, key) == 0) return e.p;
    return nullptr;
}
void tuning_from_env(mceik_ctx *c) {
    for (const char *k : {"ZC", "BY", "NO16", "PUBLISH", "PUBLISHER", "NO_STAGGER", "NATURAL", "L2PF", "BATCH", "TRACE",
                          "STATS", "DEBUG"}) {
        const std::string name = std::string("MCEIK_FSM_") + k;
        if (const char *e = getenv(name.c_str())) *tuning_slot(c, k) = atoi(e);
    }
    if (const char *e = getenv("MCEIK_LOCATE_NO_ALIGN")) c->tune.locate_no_align = atoi(e);
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) MCEIK_CUDA(cudaSetDevice(dev));
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <class F>
int guarded(F &&f) {
    try {
        return f();
    } catch (const std::exception &e) {
        set_error("%s", e.what());
        return -2;
    }
}

template <class T>
T *upload(DevBuf &buf, size_t offset_bytes, const std::vector<T> &v, cudaStream_t st) {
    T *d = reinterpret_cast<T *>(static_cast<char *>(buf.p) + offset_bytes);
    if (!v.empty()) MCEIK_CUDA(cudaMemcpyAsync(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, st));
    return d;
}
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------
// eikonal: the batched solve on device-resident slow / u
// ------------------------------------------------------------------------------------------
int fsm_solve_dev(mceik_ctx *ctx, const mceik_fsm_grid *g, int nmodels, const double *d_slow, int nfields,
                  const int *field_model, const int *src_ptr, const double *ts, const double *xs,
                  const double *ys, const double *zs, double *d_u, float *d_tables, size_t ldtab, int *iters,
                  int *field_ierr) {
    if (!ctx || !g || !d_slow || !field_model || !src_ptr || (nfields > 0 && (!ts || !xs || !ys || !zs))) {
        set_error("mceik_fsm_solve_batched: NULL argument");
        return -1;
    }
    if (g->nx < 1 || g->ny < 1 || g->nz < 1 || nmodels < 1 || nfields < 0 || nfields > 65535) {
        set_error("mceik_fsm_solve_batched: bad sizes (nx,ny,nz >= 1, 0 <= nfields <= 65535)");
        return -1;
    }
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    if (N > (size_t)INT_MAX) {
        set_error("mceik_fsm_solve_batched: grid has more than 2^31-1 nodes");
        return -1;
    }
    if (d_tables && ldtab < N) {
        set_error("mceik_fsm_solve_batched: ldtab < nx*ny*nz");
        return -1;
    }
    if ((reinterpret_cast<uintptr_t>(d_u) | reinterpret_cast<uintptr_t>(d_slow)) % 16 != 0) {
        set_error("mceik_fsm_solve_batched: d_u and d_slow must be 16-byte aligned (128-bit accesses)");
        return -1;
    }
    for (int f = 0; f < nfields; ++f)
        if (field_model[f] < 0 || field_model[f] >= nmodels) {
            set_error("mceik_fsm_solve_batched: field_model[%d] out of range", f);
            return -1;
        }
    ctx->last_updates = 0;
    ctx->last_sweep_ms = 0.0;
    ctx->last_sweep_launches = 0;
    if (nfields == 0) return 0;
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->stream;
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    if (!ctx->ev0) {
        MCEIK_CUDA(cudaEventCreate(&ctx->ev0));
        MCEIK_CUDA(cudaEventCreate(&ctx->ev1));
        if (!ctx->ev_fin) MCEIK_CUDA(cudaEventCreateWithFlags(&ctx->ev_fin, cudaEventDisableTiming));
    }

    ctx->plan.build(nx, ny, nz, st);
    const fsm::TilePlan &pl = ctx->plan;
    const bool bricks = ctx->fsm_algo == MCEIK_FSM_ALGO_BRICKS;
    if (bricks) ctx->bplan.build(nx, ny, nz, ctx->tune.by == 16 ? 16 : 8, std::max(1, std::min(256, ctx->tune.zc)), st);
    const fsm::BrickPlan &bp = ctx->bplan;

    // ---- boundary conditions: stencil records per field (host), unique node lists per field
    std::vector<fsm::BcRecord> recs;
    std::vector<int> rec_ptr(nfields + 1, 0), ferr(nfields, 0);
    std::vector<int> bc_ptr(nfields + 1, 0), bc_tile, rec_field, rec_node;
    std::vector<uint16_t> bc_local;
    for (int f = 0; f < nfields; ++f) {
        const int s0 = src_ptr[f], ns = src_ptr[f + 1] - src_ptr[f];
        const size_t before = recs.size();
        ferr[f] = ns < 0 ? 1 : host::build_bc_records(nx, ny, nz, g->h, g->x0, g->y0, g->z0, ns, ts + s0, xs + s0,
                                                      ys + s0, zs + s0, recs);
        if (ferr[f]) recs.resize(before);  // SETBCS failed: the field is not solved (fsm3d.f90:2022-2026)
        rec_ptr[f + 1] = (int)recs.size();
        std::vector<int> nodes;
        for (size_t r = before; r < recs.size(); ++r) nodes.push_back(recs[r].node);
        std::sort(nodes.begin(), nodes.end());
        nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
        for (int node : nodes) {
            const int ix = node % nx, iy = (node / nx) % ny, iz = node / (nx * ny);
            const int I = ix / fsm::kTile, J = iy / fsm::kTile, K = iz / fsm::kTile;
            bc_tile.push_back((K * pl.nty + J) * pl.ntx + I);
            bc_local.push_back((uint16_t)((ix % fsm::kTile) | ((iy % fsm::kTile) << 4) | ((iz % fsm::kTile) << 8)));
            rec_field.push_back(f);
            rec_node.push_back(node);
        }
        bc_ptr[f + 1] = (int)bc_tile.size();
    }
    std::vector<int> fmodel(field_model, field_model + nfields);

    // meta buffer layout (all offsets 256-byte aligned)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    const size_t o_fmodel = take(sizeof(int) * nfields), o_recptr = take(sizeof(int) * (nfields + 1));
    const size_t o_recs = take(sizeof(fsm::BcRecord) * std::max<size_t>(recs.size(), 1));
    const size_t o_bcptr = take(sizeof(int) * (nfields + 1)), o_bctile = take(sizeof(int) * std::max<size_t>(bc_tile.size(), 1));
    const size_t o_bcloc = take(sizeof(uint16_t) * std::max<size_t>(bc_local.size(), 1));
    const size_t o_recf = take(sizeof(int) * std::max<size_t>(rec_field.size(), 1));
    const size_t o_recn = take(sizeof(int) * std::max<size_t>(rec_node.size(), 1));
    const size_t o_gfields = take(sizeof(int) * fsm::kMaxSlots * nfields), o_gmodel = take(sizeof(int) * nfields);
    const size_t o_active = take(sizeof(int) * nfields), o_units = take(sizeof(int) * nfields);
    const size_t o_vptr = take(sizeof(long long) * (8 * (size_t)std::max(bp.nblevels, 1) + std::max(bp.nblevels, 1) + 2));
    ctx->ws_meta.ensure(off);
    const int *d_fmodel = upload(ctx->ws_meta, o_fmodel, fmodel, st);
    const int *d_recptr = upload(ctx->ws_meta, o_recptr, rec_ptr, st);
    const fsm::BcRecord *d_recs = upload(ctx->ws_meta, o_recs, recs, st);
    const int *d_bcptr = upload(ctx->ws_meta, o_bcptr, bc_ptr, st);
    const int *d_bctile = upload(ctx->ws_meta, o_bctile, bc_tile, st);
    const uint16_t *d_bcloc = upload(ctx->ws_meta, o_bcloc, bc_local, st);
    const int *d_recf = upload(ctx->ws_meta, o_recf, rec_field, st);
    const int *d_recn = upload(ctx->ws_meta, o_recn, rec_node, st);


    // ---- iterations (fsm3d.f90:62-96).  One launch = 8 sweeps of every still-active field.
    std::vector<int> active, it(nfields, 0);
    for (int f = 0; f < nfields; ++f)
        if (!ferr[f]) active.push_back(f);
    // control block: [queue int (256 B)] [nonconv ull x nfields] [done int x ngroups*ntiles]
    const size_t c_nonconv = 256, c_done = align_up(c_nonconv + sizeof(unsigned long long) * nfields);
    ctx->ws_ctrl.ensure(c_done + sizeof(int) * (size_t)nfields * std::max(pl.ntiles, bricks ? bp.nbricks : 0));
    char *ctrl = static_cast<char *>(ctx->ws_ctrl.p);
    unsigned long long *d_nonconv = reinterpret_cast<unsigned long long *>(ctrl + c_nonconv);
    std::vector<unsigned long long> h_nonconv(nfields);

    // the 16-byte-pair brick kernel reads slow*h, formed once per solve instead of once per node visit
    bool bricks16 = bricks && nx % 8 == 0 && bp.by == 8 && !ctx->tune.no16;
    if (bricks16) {  // its per-brick list of boundary-condition planes is short: fields with many sources stacked in one brick
        std::map<std::pair<int, int>, std::vector<int>> planes;  // (field, brick) -> planes holding stencil nodes
        for (size_t r = 0; r < rec_node.size(); ++r) {
            const int node = rec_node[r], gx = node % nx, gy = (node / nx) % ny, gz = node / (nx * ny);
            std::vector<int> &v = planes[{rec_field[r], ((gz / bp.zc) * bp.nby + gy / bp.by) * bp.nbx + gx / 8}];
            if (std::find(v.begin(), v.end(), gz) == v.end()) v.push_back(gz);
        }
        for (auto &kv : planes)
            if ((int)kv.second.size() > fsm::bricks16_max_bc_planes()) bricks16 = false;  // -> generic brick kernel
    }
    // bricks16 works on its own blocked copy of the fields (BrickArgs::blocked: 512 contiguous bytes per brick plane
    // + x-face copies; the [z][y][x] walk tops out at ~4.4 TB/s of DRAM traffic in 64-byte pieces, the blocked one at
    // ~6.1 TB/s, tools/stream_bench.cu) and reads slow*h, formed once per solve instead of once per node visit.
    // Converged fields are converted back into d_u.
    const bool blocked = bricks16 && !ctx->tune.natural;
    const size_t Nb = blocked ? fsm::blocked_field_doubles(nx, ny, nz) : N;  // doubles per field the iterations work on
    const double *d_fh = nullptr;
    double *d_w = nullptr, *d_w0 = nullptr;  // fields / start-of-iteration copy the sweeps and the convergence test use
    // ---- u = HUGE everywhere, then the stencil values (fsm3d.f90:782-834)
    if (blocked) {  // the caller's d_u (may be NULL: tables only) is written once, when a field has converged
        double *fh = static_cast<double *>(ctx->ws_fh.ensure(sizeof(double) * fsm::blocked_slowness_doubles(nx, ny, nz) * nmodels));
        fsm::launch_scale_slowness_blocked(nx, ny, nz, nmodels, g->h, d_slow, fh, st);
        d_fh = fh;
        d_w = static_cast<double *>(ctx->ws_ub.ensure(sizeof(double) * Nb * nfields));
        d_w0 = static_cast<double *>(ctx->ws_u0b.ensure(sizeof(double) * Nb * nfields));
        fsm::launch_fill(d_w, Nb * nfields, DBL_MAX, st);
        fsm::launch_apply_bcs_blocked(nfields, nx, ny, nz, d_fmodel, d_recptr, d_recs, d_slow, d_w, st);
    } else {
        if (!d_u) d_u = static_cast<double *>(ctx->ws_u.ensure(sizeof(double) * N * nfields));
        d_w = d_u;
        fsm::launch_fill(d_u, N * nfields, DBL_MAX, st);
        fsm::launch_apply_bcs(nfields, N, d_fmodel, d_recptr, d_recs, d_slow, d_u, st);
        if (bricks16) {
            double *fh = static_cast<double *>(ctx->ws_fh.ensure(sizeof(double) * N * nmodels));
            fsm::launch_scale_slowness(N * nmodels, g->h, d_slow, fh, st);
            d_fh = fh;
        }
        d_w0 = static_cast<double *>(ctx->ws_u0.ensure(sizeof(double) * N * nfields));
    }
    double *d_u0 = d_w0;
    if (bricks)  // u0 = u before the first iteration (fsm3d.f90:60); refreshed by the convergence kernel
        MCEIK_CUDA(cudaMemcpyAsync(d_w0, d_w, sizeof(double) * Nb * nfields, cudaMemcpyDeviceToDevice, st));
    uint8_t *d_lupd = nullptr;
    if (ctx->fsm_algo == MCEIK_FSM_ALGO_LEVELS) {
        d_lupd = static_cast<uint8_t *>(ctx->ws_lupd.ensure(N * nfields));
        MCEIK_CUDA(cudaMemsetAsync(d_lupd, 1, N * nfields, st));
        fsm::launch_mark_bcs((int)rec_field.size(), d_recf, d_recn, N, d_lupd, st);
        MCEIK_CUDA(cudaMemcpyAsync(d_u0, d_u, sizeof(double) * N * nfields, cudaMemcpyDeviceToDevice, st));
    }

    for (int k = 1; k <= g->maxit && !active.empty(); ++k) {
        MCEIK_CUDA(cudaMemsetAsync(ctrl, 0, c_done, st));
        if (ctx->fsm_algo == MCEIK_FSM_ALGO_LEVELS) {
            const int *d_active = upload(ctx->ws_meta, o_active, active, st);
            fsm::launch_iteration_levels(nx, ny, nz, g->h, (int)active.size(), d_active, d_fmodel, d_slow, d_lupd,
                                         d_u, st);
            fsm::launch_convergence(N, (int)active.size(), d_active, g->tol, d_u, d_u0, d_nonconv, st);
        } else if (bricks) {
            const int *d_active = upload(ctx->ws_meta, o_active, active, st);
            fsm::BrickArgs a;
            a.nx = nx; a.ny = ny; a.nz = nz;
            a.nbx = bp.nbx; a.nby = bp.nby; a.nbz = bp.nbz; a.nbricks = bp.nbricks; a.nblevels = bp.nblevels; a.zc = bp.zc; a.by = bp.by;
            a.nfields_active = (int)active.size();
            // few fields: short publication interval (tight pipelining of the brick wavefront);
            // many fields: parallelism is plentiful, publish less often (each publication costs a fence)
            a.publish = active.size() >= 48 ? 16 : (active.size() > 16 ? 8 : 4);  // measured, profiles/kernel_evolution_r1.md
            if (ctx->tune.publish >= 2 && (ctx->tune.publish & (ctx->tune.publish - 1)) == 0) a.publish = ctx->tune.publish;  // a power of two
            a.h = g->h;
            a.active = d_active; a.field_model = d_fmodel; a.slow = bricks16 ? d_fh : d_slow; a.slow_is_fh = bricks16 ? 1 : 0;
            a.u = d_w; a.blocked = blocked ? 1 : 0;
            a.l2_prefetch = ctx->tune.l2pf;
            a.brick_order = ctx->bplan.brick_order.as<int>();
            a.blevel_ptr = ctx->bplan.blevel_ptr.as<int>();
            a.queue = reinterpret_cast<unsigned long long *>(ctrl);
            a.done = reinterpret_cast<int *>(ctrl + c_done);
            a.bc_ptr = d_bcptr; a.bc_node = d_recn;
            a.vptr = nullptr; a.nf0 = a.nfields_active; a.stagger = 0; a.batch = 0;
            // few active fields: a publisher warp per CTA takes the release fences off the sweeping warps (+8-12 % up to
            // 11 fields, +1 % at 16, nothing beyond; profiles/kernel_evolution_r1.md)
            a.publisher = active.size() <= 16 ? 1 : 0;
            if (ctx->tune.publisher >= 0) a.publisher = ctx->tune.publisher != 0;
            if (bricks16) {  // two field groups half a sweep apart (see BrickArgs)
                const int nl = bp.nblevels, nfa = a.nfields_active;
                a.nf0 = (nfa + 1) / 2;
                a.stagger = (nfa >= 12 && !ctx->tune.no_stagger) ? nl / 2 : 0;  // measured: +2.5 % at 16 fields, -3.6 % at 4
                if (a.stagger == 0) a.nf0 = nfa;  // one group holds every field
                a.batch = ctx->tune.batch;
                if (a.batch >= nfa) a.batch = 0;
                if (a.batch > 0) { a.stagger = 0; a.nf0 = nfa; }
                const std::vector<long long> vptr = a.batch > 0 ? host::build_ticket_table(nl, bp.h_blevel_ptr.data(), 1, 1, 0)
                                                                : host::build_ticket_table(nl, bp.h_blevel_ptr.data(), nfa, a.nf0, a.stagger);
                a.vptr = upload(ctx->ws_meta, o_vptr, vptr, st);
            }
            a.debug = ctx->tune.debug;
            a.stats = ctx->tune.stats ? reinterpret_cast<unsigned long long *>(ctrl + 64) : nullptr;
            MCEIK_CUDA(cudaMemsetAsync(a.done, 0, sizeof(int) * (size_t)nfields * bp.nbricks, st));
            MCEIK_CUDA(cudaEventRecord(ctx->ev0, st));
            if (bricks16) fsm::launch_iteration_bricks16(a, st);
            else fsm::launch_iteration_bricks(a, st);
            MCEIK_CUDA(cudaEventRecord(ctx->ev1, st));
            ctx->last_sweep_launches += 1;
            if (blocked) fsm::launch_convergence_blocked(nx, ny, nz, (int)active.size(), d_active, g->tol, d_w, d_w0, d_nonconv, st);
            else fsm::launch_convergence(N, (int)active.size(), d_active, g->tol, d_w, d_w0, d_nonconv, st);
        } else {
            // group the active fields by slowness model, up to kMaxSlots per CTA
            std::map<int, std::vector<int>> by_model;
            for (int f : active) by_model[fmodel[f]].push_back(f);
            size_t widest = 0;
            for (auto &kv : by_model) widest = std::max(widest, kv.second.size());
            const int B = widest >= 3 ? 4 : (widest == 2 ? 2 : 1);
            std::vector<int> gfields, gmodel;
            for (auto &kv : by_model)
                for (size_t i = 0; i < kv.second.size(); i += B) {
                    gmodel.push_back(kv.first);
                    for (int b = 0; b < fsm::kMaxSlots; ++b)
                        gfields.push_back(b < B && i + b < kv.second.size() ? kv.second[i + b] : -1);
                }
            const int ngroups = (int)gmodel.size();
            if ((long long)ngroups * pl.ntiles * 8 > (long long)INT_MAX) throw CudaError("too many tile tasks in one launch");
            fsm::SweepArgs a;
            a.nx = nx; a.ny = ny; a.nz = nz;
            a.ntx = pl.ntx; a.nty = pl.nty; a.ntz = pl.ntz; a.ntiles = pl.ntiles; a.ntlevels = pl.ntlevels;
            a.ngroups = ngroups; a.nslots = B;
            a.h = g->h; a.tol = g->tol;
            a.group_fields = upload(ctx->ws_meta, o_gfields, gfields, st);
            a.group_model = upload(ctx->ws_meta, o_gmodel, gmodel, st);
            a.slow = d_slow; a.u = d_u; a.u0 = d_u0;
            a.tile_order = ctx->plan.tile_order.as<int>();
            a.tlevel_ptr = ctx->plan.tlevel_ptr.as<int>();
            a.lvl_nodes = ctx->plan.lvl_nodes.as<uint16_t>();
            a.queue = reinterpret_cast<int *>(ctrl);
            a.nonconv = d_nonconv;
            a.done = reinterpret_cast<int *>(ctrl + c_done);
            a.bc_ptr = d_bcptr; a.bc_tile = d_bctile; a.bc_local = d_bcloc;
            MCEIK_CUDA(cudaMemsetAsync(a.done, 0, sizeof(int) * (size_t)ngroups * pl.ntiles, st));
            MCEIK_CUDA(cudaEventRecord(ctx->ev0, st));
            fsm::launch_iteration_tiles(a, st);
            MCEIK_CUDA(cudaEventRecord(ctx->ev1, st));
            ctx->last_sweep_launches += 1;
        }
        MCEIK_CUDA(cudaMemcpyAsync(h_nonconv.data(), d_nonconv, sizeof(unsigned long long) * nfields,
                                   cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaStreamSynchronize(st));
        if (bricks && ctx->tune.stats) {
            unsigned long long hs[8];
            MCEIK_CUDA(cudaMemcpy(hs, ctrl + 64, sizeof(hs), cudaMemcpyDeviceToHost));
            if (hs[3]) printf("[fsm stats] iter %d: tasks %llu  avg cycles: start-wait %.0f  upwind-wait %.0f  run %.0f\n", k, hs[3],
                              (double)hs[0] / hs[3], (double)hs[1] / hs[3], (double)hs[2] / hs[3]);
            if (hs[3] && hs[4])  // -DMCEIK_B16_PROFILE builds: phases of a step, cycles per task
                printf("[fsm stats]   step phases per task: issue %.0f  copy-wait+sync %.0f  reads+solve+writes %.0f  write-back %.0f\n",
                       (double)hs[4] / hs[3], (double)hs[5] / hs[3], (double)hs[6] / hs[3], (double)hs[7] / hs[3]);
        }
        if (ctx->fsm_algo != MCEIK_FSM_ALGO_LEVELS) {
            float ms = 0.f;
            MCEIK_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            ctx->last_sweep_ms += ms;
            if (ctx->tune.trace) printf("[fsm trace] iter %d: %zu active fields, sweep kernel %.2f ms\n", k, active.size(), ms);
        }
        std::vector<int> still;
        for (int f : active) {
            it[f] = k;
            if (h_nonconv[f] != 0) still.push_back(f);  // lconv /= nxyz -> next iteration (fsm3d.f90:95)
        }
        std::vector<int> finished;  // fields that have just converged (or run out of iterations): final
        {
            size_t j = 0;
            for (int f : active) {
                if (k < g->maxit && j < still.size() && still[j] == f) { ++j; continue; }
                finished.push_back(f);
            }
        }
        if (blocked && !finished.empty()) {  // into the caller's layout: fp64 field and / or fp32 table, one pass
            const int *d_fin = upload(ctx->ws_meta, o_units, finished, st);
            fsm::launch_unblock_fields(nx, ny, nz, (int)finished.size(), d_fin, d_w, d_u, d_tables, ldtab, st);
        }
        if (ctx->put_enabled && !finished.empty()) {  // their tables go to the other ranks while the others keep iterating
            MCEIK_CUDA(cudaEventRecord(ctx->ev_fin, st));
            MCEIK_CUDA(cudaStreamWaitEvent(ctx->put_stream, ctx->ev_fin, 0));
            for (int f : finished)
                comm::put_to_peers(ctx->comm, sizeof(float) * (ctx->put_row0 + (size_t)f) * ctx->put_ldtab, sizeof(float) * ctx->put_n,
                                   ctx->put_stream);
        }
        if (ctx->early_u && d_u && !finished.empty()) {  // copy them back while the others keep iterating
            if (blocked) {
                MCEIK_CUDA(cudaEventRecord(ctx->ev_fin, st));
                MCEIK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fin, 0));
            }  // else: the stream is idle here (synchronised above) and d_u[f] is final
            for (int f : finished) {
                MCEIK_CUDA(cudaMemcpyAsync(ctx->early_u + (size_t)f * N, d_u + (size_t)f * N, sizeof(double) * N,
                                           cudaMemcpyDeviceToHost, ctx->copy_stream));
                ctx->early_done[f] = 1;
            }
        }
        active.swap(still);
    }
    if (blocked && g->maxit < 1)  // no iteration ran: the fields are their boundary conditions
        fsm::launch_unblock_fields(nx, ny, nz, nfields, nullptr, d_w, d_u, d_tables, ldtab, st);
    int rc = 0;
    for (int f = 0; f < nfields; ++f) {
        ctx->last_updates += (long long)N * 8 * it[f];
        if (iters) iters[f] = it[f];
        if (field_ierr) field_ierr[f] = ferr[f];
        if (ferr[f]) rc = 1;
    }
    if (d_tables && !blocked) fsm::launch_pack_tables(nfields, N, ldtab, d_u, d_tables, st);
    return rc;
}

// ------------------------------------------------------------------------------------------
// locator: batched search on device-resident pick arrays
// ------------------------------------------------------------------------------------------
int locate_dev(mceik_ctx *ctx, int job, int nevents, int nobs_total, int max_picks, const int *d_obs_ptr,
               const int *d_table_id, const double *d_tobs_cor, const double *d_varobs, const double *d_tori,
               int *d_iopt, double *d_t0opt, double *d_objopt) {
    if (!ctx || !ctx->d_tables) {
        set_error("mceik_locate_batched: no travel-time tables set");
        return -1;
    }
    if (job != 1 && job != 2) {
        set_error("mceik_locate_batched: job %d not supported (1 = fixed origin time, 2 = analytic origin time)", job);
        return 1;
    }
    if (nevents < 0 || nobs_total < 0 || max_picks < 0 || !d_obs_ptr || (job == 1 && !d_tori) || !d_iopt || !d_t0opt ||
        !d_objopt) {
        set_error("mceik_locate_batched: bad argument");
        return -1;
    }
    if (nevents == 0) return 0;
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->stream;
    const size_t np = std::max(nobs_total, 1);
    const int nblk_all = (nevents + gs::kEventsPerBlock - 1) / gs::kEventsPerBlock;
    double *d_w = static_cast<double *>(ctx->ws_gs_w.ensure(align_up(sizeof(double) * np) * 2 + align_up(sizeof(int) * nevents) +
                                                            sizeof(int) * nblk_all));
    double *d_w0 = d_w;
    double *d_w1 = reinterpret_cast<double *>(reinterpret_cast<char *>(d_w) + align_up(sizeof(double) * np));
    int *d_nuse = reinterpret_cast<int *>(reinterpret_cast<char *>(d_w) + 2 * align_up(sizeof(double) * np));
    int *d_uniform = reinterpret_cast<int *>(reinterpret_cast<char *>(d_nuse) + align_up(sizeof(int) * nevents));
    gs::launch_prepare(nevents, d_obs_ptr, d_table_id, d_varobs, d_w0, d_w1, d_nuse, st);
    gs::launch_classify(nevents, d_obs_ptr, d_table_id, d_uniform, st);
    const int max_events_per_launch = 65535 * gs::kEventsPerBlock;
    for (int e0 = 0; e0 < nevents; e0 += max_events_per_launch) {
        const int ne = std::min(max_events_per_launch, nevents - e0);
        gs::LocateArgs a;
        a.job = job; a.nevents = ne; a.ngrd = ctx->ngrd; a.ldgrd = ctx->ldgrd; a.maxpicks = std::max(max_picks, 1);
        a.tables = ctx->d_tables;
        a.obs_ptr = d_obs_ptr + e0; a.table_id = d_table_id; a.tobs_cor = d_tobs_cor; a.w_t0 = d_w0; a.w_obj = d_w1;
        a.tori = d_tori ? d_tori + e0 : nullptr;
        a.blk_uniform = d_uniform + e0 / gs::kEventsPerBlock;
        a.nlanes = gs::locate_lanes(ne, ctx->ngrd);
        a.partials = static_cast<gs::Partial *>(ctx->ws_gs_part.ensure(sizeof(gs::Partial) * (size_t)ne * a.nlanes));
        gs::launch_locate(a, st);
        gs::launch_finalize(ne, a.nlanes, a.partials, d_nuse + e0, d_iopt + e0, d_t0opt + e0, d_objopt + e0, st);
    }
    return 0;
}

std::mutex g_mutex;
mceik_ctx *g_default_ctx = nullptr;

mceik_ctx *default_ctx() {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!g_default_ctx) {
        mceik_ctx *c = nullptr;
        if (mceik_ctx_create(-1, nullptr, &c) != 0) return nullptr;
        g_default_ctx = c;
    }
    return g_default_ctx;
}

// hidden state of the drop-in life-cycles (the reference keeps SAVE / module variables:
// fsm3d.f90:1985-1988, module.F90:419-423, 469-470)
struct SerialState { bool init = false; int nx = 0, ny = 0, nz = 0; } g_serial;
struct SolveState { bool init = false; mceik_fsm_grid grid{}; } g_solve;
struct LocState { bool init = false; int iverb = 0; } g_loc;

}  // namespace

// ============================================================================================
extern "C" {

const char *mceik_last_error(void) { return mceik::last_error(); }
long long mceik_kernel_launch_count(void) { return mceik::launch_count(); }

int mceik_ctx_create(int device, void *stream, mceik_ctx **out) {
    return guarded([&]() -> int {
        if (!out) { set_error("mceik_ctx_create: NULL out"); return -1; }
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            set_error("mceik_ctx_create: no CUDA device (%s); libmceik_b200 has no CPU fallback", cudaGetErrorString(e));
            cudaGetLastError();
            return -3;
        }
        if (device < 0) MCEIK_CUDA(cudaGetDevice(&device));
        if (device >= ndev) { set_error("mceik_ctx_create: device %d of %d", device, ndev); return -1; }
        cudaDeviceProp prop;
        MCEIK_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) {
            set_error("mceik_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
            return -3;
        }
        DeviceGuard dg(device);
        // The sweep kernel reads isolated 32-byte sectors (x-halo columns); the default 64-byte L2 fetch
        // granularity would double their DRAM traffic.
        if (!getenv("MCEIK_L2_FETCH_DEFAULT")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
        mceik_ctx *c = new mceik_ctx();
        c->device = device;
        tuning_from_env(c);
        if (stream) {
            c->stream = static_cast<cudaStream_t>(stream);
        } else {
            MCEIK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
            c->own_stream = true;
        }
        *out = c;
        return 0;
    });
}

void mceik_ctx_destroy(mceik_ctx *c) {
    if (!c) return;
    try {
        DeviceGuard dg(c->device);
        cudaStreamSynchronize(c->stream);
        c->plan.release();
        c->bplan.release();
        for (DevBuf *b : {&c->ws_slow, &c->ws_u, &c->ws_u0, &c->ws_tab, &c->ws_meta, &c->ws_ctrl, &c->ws_lupd, &c->ws_xyzv, &c->ws_fh, &c->ws_ub, &c->ws_u0b, &c->ws_comm,
                          &c->own_tables, &c->ws_gs_in, &c->ws_gs_w, &c->ws_gs_part, &c->ws_gs_out, &c->ws_gs_misc})
            b->release();
        if (c->ev0) cudaEventDestroy(c->ev0);
        if (c->ev1) cudaEventDestroy(c->ev1);
        if (c->ev_fin) cudaEventDestroy(c->ev_fin);
        if (c->put_stream) cudaStreamDestroy(c->put_stream);
        try { comm::destroy(c->comm); } catch (...) {}
        if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
        if (c->own_stream) cudaStreamDestroy(c->stream);
    } catch (...) {
    }
    delete c;
}

int mceik_ctx_synchronize(mceik_ctx *c) {
    return guarded([&]() -> int {
        if (!c) return -1;
        DeviceGuard dg(c->device);
        MCEIK_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    });
}

int mceik_fsm_set_algo(mceik_ctx *c, int algo) {
    if (!c || (algo != MCEIK_FSM_ALGO_TILES && algo != MCEIK_FSM_ALGO_LEVELS && algo != MCEIK_FSM_ALGO_BRICKS)) return -1;
    c->fsm_algo = algo;
    return 0;
}

int mceik_fsm_set_tuning(mceik_ctx *c, const char *key, int value) {
    int *p = (c && key) ? tuning_slot(c, key) : nullptr;
    if (!p) {
        set_error("mceik_fsm_set_tuning: unknown key");
        return -1;
    }
    *p = value;
    return 0;
}
long long mceik_fsm_last_node_updates(mceik_ctx *c) { return c ? c->last_updates : 0; }
int mceik_fsm_last_sweep_stats(mceik_ctx *c, double *sweep_ms, int *launches) {
    if (!c) return -1;
    if (sweep_ms) *sweep_ms = c->last_sweep_ms;
    if (launches) *launches = c->last_sweep_launches;
    return 0;
}

int mceik_selftest_solver(mceik_ctx *ctx, unsigned long long seed, long long samples, long long *bad_sqrt,
                          long long *bad_solve) {
    return guarded([&]() -> int {
        if (!ctx || !bad_sqrt || !bad_solve) return -1;
        DeviceGuard dg(ctx->device);
        unsigned long long *d = static_cast<unsigned long long *>(ctx->ws_ctrl.ensure(256));
        MCEIK_CUDA(cudaMemsetAsync(d, 0, 16, ctx->stream));
        const int blocks = 148 * 8, per = (int)std::max<long long>(1, samples / (blocks * 256LL));
        fsm::launch_selftest(seed, blocks, per, d, ctx->stream);
        unsigned long long h[2] = {0, 0};
        MCEIK_CUDA(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
        MCEIK_CUDA(cudaStreamSynchronize(ctx->stream));
        *bad_sqrt = (long long)h[0];
        *bad_solve = (long long)h[1];
        return 0;
    });
}

int mceik_fsm_solve_batched_dev(mceik_ctx *ctx, const mceik_fsm_grid *grid, int nmodels, const double *d_slow,
                                int nfields, const int *field_model, const int *src_ptr, const double *ts,
                                const double *xs, const double *ys, const double *zs, double *d_u, float *d_tables,
                                size_t ldtab, int *iters, int *field_ierr) {
    return guarded([&]() -> int {
        return fsm_solve_dev(ctx, grid, nmodels, d_slow, nfields, field_model, src_ptr, ts, xs, ys, zs, d_u, d_tables,
                             ldtab, iters, field_ierr);
    });
}

int mceik_fsm_solve_batched_host(mceik_ctx *ctx, const mceik_fsm_grid *grid, int nmodels, const double *slow, int nfields,
                                 const int *field_model, const int *src_ptr, const double *ts, const double *xs,
                                 const double *ys, const double *zs, double *u, float *tables, size_t ldtab, int *iters,
                                 int *field_ierr) {
    return guarded([&]() -> int {
        if (!ctx || !grid || !slow) { set_error("mceik_fsm_solve_batched_host: NULL argument"); return -1; }
        if (grid->nx < 1 || grid->ny < 1 || grid->nz < 1 || nmodels < 1 || nfields < 0) {
            set_error("mceik_fsm_solve_batched_host: bad sizes");
            return -1;
        }
        DeviceGuard dg(ctx->device);
        const size_t N = (size_t)grid->nx * grid->ny * grid->nz;
        double *d_slow = static_cast<double *>(ctx->ws_slow.ensure(sizeof(double) * N * nmodels));
        MCEIK_CUDA(cudaMemcpyAsync(d_slow, slow, sizeof(double) * N * nmodels, cudaMemcpyHostToDevice, ctx->stream));
        double *d_u = u ? static_cast<double *>(ctx->ws_u.ensure(sizeof(double) * N * std::max(nfields, 1))) : nullptr;
        float *d_tab = nullptr;
        if (tables) d_tab = static_cast<float *>(ctx->ws_tab.ensure(sizeof(float) * ldtab * std::max(nfields, 1)));
        // pinned destination: copy every field back as soon as it has converged, overlapped with the sweeps of
        // the others (pageable memory would make those copies block the host between iterations)
        cudaPointerAttributes pa{};
        const bool pinned = u && nfields > 1 && cudaPointerGetAttributes(&pa, u) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
        ctx->early_u = nullptr;
        if (pinned) {
            if (!ctx->copy_stream) MCEIK_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            ctx->early_u = u;
            ctx->early_done.assign(nfields, 0);
        }
        int rc;
        try {
            rc = fsm_solve_dev(ctx, grid, nmodels, d_slow, nfields, field_model, src_ptr, ts, xs, ys, zs, d_u, d_tab, ldtab, iters,
                               field_ierr);
        } catch (...) {
            ctx->early_u = nullptr;
            if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);  // no copy into the caller's buffer may outlive the call
            throw;
        }
        ctx->early_u = nullptr;
        if (rc < 0) {
            if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
            return rc;
        }
        if (u && pinned) {
            for (int f = 0; f < nfields; ++f)  // fields that stopped at maxit or failed their boundary conditions
                if (!ctx->early_done[f])
                    MCEIK_CUDA(cudaMemcpyAsync(u + (size_t)f * N, d_u + (size_t)f * N, sizeof(double) * N, cudaMemcpyDeviceToHost,
                                               ctx->stream));
            MCEIK_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        } else if (u) {
            MCEIK_CUDA(cudaMemcpyAsync(u, d_u, sizeof(double) * N * nfields, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (tables)
            MCEIK_CUDA(cudaMemcpyAsync(tables, d_tab, sizeof(float) * ldtab * nfields, cudaMemcpyDeviceToHost, ctx->stream));
        MCEIK_CUDA(cudaStreamSynchronize(ctx->stream));
        return rc;
    });
}

// ------------------------------------------------------------------------------------------
// multi-GPU: sources sharded over the ranks of a communicator, tables replicated (comm.cu)
// ------------------------------------------------------------------------------------------
int mceik_comm_unique_id(void *id128) {
    return guarded([&]() -> int {
        if (!id128) { set_error("mceik_comm_unique_id: NULL"); return -1; }
        comm::unique_id(id128);
        return 0;
    });
}

int mceik_comm_init(mceik_ctx *ctx, int world, int rank, const void *id128) {
    return guarded([&]() -> int {
        if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) { set_error("mceik_comm_init: bad argument"); return -1; }
        if (ctx->comm) { set_error("mceik_comm_init: the context already has a communicator"); return 1; }
        DeviceGuard dg(ctx->device);
        ctx->comm = comm::create(world, rank, id128);
        return 0;
    });
}

int mceik_comm_destroy(mceik_ctx *ctx) {
    return guarded([&]() -> int {
        if (!ctx) return -1;
        DeviceGuard dg(ctx->device);
        comm::destroy(ctx->comm);
        ctx->comm = nullptr;
        return 0;
    });
}

int mceik_fsm_assign_fields(int nfields, const int *field_model, const int *cost, int world, int *rank_of_field, int *table_row,
                            int *slots) {
    return guarded([&]() -> int {
        if (nfields < 0 || world < 1 || (nfields > 0 && !field_model)) { set_error("mceik_fsm_assign_fields: bad argument"); return -1; }
        std::vector<int> rk, row;
        int sl = 0;
        comm::assign_fields(nfields, field_model, cost, world, rk, row, sl);
        for (int f = 0; f < nfields; ++f) {
            if (rank_of_field) rank_of_field[f] = rk[f];
            if (table_row) table_row[f] = row[f];
        }
        if (slots) *slots = sl;
        return 0;
    });
}

int mceik_tables_alloc_replicated(mceik_ctx *ctx, size_t rows, size_t ldtab, float **d_tables_all) {
    return guarded([&]() -> int {
        if (!ctx || !d_tables_all || rows == 0 || ldtab == 0) { set_error("mceik_tables_alloc_replicated: bad argument"); return -1; }
        DeviceGuard dg(ctx->device);
        *d_tables_all = static_cast<float *>(comm::alloc_replicated(ctx->comm, sizeof(float) * rows * ldtab, ctx->stream));
        return 0;
    });
}

int mceik_tables_free_replicated(mceik_ctx *ctx) {
    return guarded([&]() -> int {
        if (!ctx) return -1;
        DeviceGuard dg(ctx->device);
        if (ctx->put_stream) MCEIK_CUDA(cudaStreamSynchronize(ctx->put_stream));
        comm::free_replicated(ctx->comm);
        return 0;
    });
}

int mceik_tables_allgather(mceik_ctx *ctx, float *d_tables_all, size_t ldtab, int slots) {
    return guarded([&]() -> int {
        if (!ctx || !d_tables_all || slots < 0) { set_error("mceik_tables_allgather: bad argument"); return -1; }
        DeviceGuard dg(ctx->device);
        comm::all_gather_inplace(ctx->comm, d_tables_all, sizeof(float) * ldtab * (size_t)slots, ctx->stream);
        return 0;
    });
}

int mceik_fsm_solve_sharded_dev(mceik_ctx *ctx, const mceik_fsm_grid *grid, int nmodels, const double *d_slow, int nfields,
                                const int *field_model, const int *src_ptr, const double *ts, const double *xs, const double *ys,
                                const double *zs, const int *cost, float *d_tables_all, size_t ldtab, int *iters, int *field_ierr,
                                int *table_row) {
    return guarded([&]() -> int {
        if (!ctx || !grid || !d_slow || !field_model || !src_ptr || !d_tables_all || nfields < 0) {
            set_error("mceik_fsm_solve_sharded_dev: bad argument");
            return -1;
        }
        DeviceGuard dg(ctx->device);
        const int world = comm::world(ctx->comm), rank = comm::rank(ctx->comm);
        std::vector<int> rk, row;
        int slots = 0;
        comm::assign_fields(nfields, field_model, cost, world, rk, row, slots);
        // this rank's fields, in increasing field order (= increasing table row)
        std::vector<int> mine, fm, sp(1, 0);
        std::vector<double> lts, lxs, lys, lzs;
        for (int f = 0; f < nfields; ++f) {
            if (rk[f] != rank) continue;
            mine.push_back(f);
            fm.push_back(field_model[f]);
            for (int q = src_ptr[f]; q < src_ptr[f + 1]; ++q) { lts.push_back(ts[q]); lxs.push_back(xs[q]); lys.push_back(ys[q]); lzs.push_back(zs[q]); }
            sp.push_back((int)lts.size());
        }
        const int nl = (int)mine.size();
        std::vector<int> stat(2 * (size_t)slots * world, 0);  // per rank: [iters x slots][ierr x slots]
        int *my = stat.data() + 2 * (size_t)slots * rank;
        // the tables of the local fields land in this rank's rows of the replicated buffer.  In a buffer from
        // mceik_tables_alloc_replicated every table is put into the peers' copies by the copy engines as soon as its
        // field has converged (no collective, no rank waits for another); any other buffer is completed by one in-place
        // all-gather after the solve.
        const size_t N = (size_t)grid->nx * grid->ny * grid->nz;
        const bool one_sided = world > 1 && d_tables_all == comm::replicated_local(ctx->comm) &&
                               sizeof(float) * ldtab * (size_t)slots * world <= comm::replicated_bytes(ctx->comm);
        if (one_sided) {
            if (!ctx->put_stream) MCEIK_CUDA(cudaStreamCreateWithFlags(&ctx->put_stream, cudaStreamNonBlocking));
            if (!ctx->ev_fin) MCEIK_CUDA(cudaEventCreateWithFlags(&ctx->ev_fin, cudaEventDisableTiming));  // a rank without fields never solves
            ctx->put_enabled = true;
            ctx->put_row0 = (size_t)rank * slots; ctx->put_ldtab = ldtab; ctx->put_n = N;
        }
        // A rank whose own solve fails (bad argument, CUDA error) still takes part in the collectives below -- the
        // others would wait for it forever -- and marks its fields with ierr = -1; it reports the failure afterwards.
        int rc = 0;
        std::string local_error;
        if (nl > 0) {  // (a rank without fields has nothing to solve: more ranks than fields)
            try {
                rc = fsm_solve_dev(ctx, grid, nmodels, d_slow, nl, fm.data(), sp.data(), lts.data(), lxs.data(), lys.data(), lzs.data(),
                                   nullptr, d_tables_all + (size_t)rank * slots * ldtab, ldtab, my, my + slots);
                if (rc < 0) local_error = mceik_last_error();
            } catch (const std::exception &e) {
                rc = -2;
                local_error = e.what();
            }
        }
        ctx->put_enabled = false;
        if (rc < 0)
            for (int i = 0; i < nl; ++i) my[slots + i] = -1;
        if (one_sided) {  // my puts are complete before I contribute to the all-gather below, which therefore is the barrier
            MCEIK_CUDA(cudaEventRecord(ctx->ev_fin, ctx->put_stream));
            MCEIK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_fin, 0));
        } else {
            comm::all_gather_inplace(ctx->comm, d_tables_all, sizeof(float) * ldtab * (size_t)slots, ctx->stream);
        }
        if (world > 1) {  // iteration counts and error flags of every field, on every rank
            int *d_stat = static_cast<int *>(ctx->ws_comm.ensure(sizeof(int) * stat.size()));
            MCEIK_CUDA(cudaMemcpyAsync(d_stat + 2 * (size_t)slots * rank, my, sizeof(int) * 2 * slots, cudaMemcpyHostToDevice, ctx->stream));
            comm::all_gather_inplace(ctx->comm, d_stat, sizeof(int) * 2 * (size_t)slots, ctx->stream);
            MCEIK_CUDA(cudaMemcpyAsync(stat.data(), d_stat, sizeof(int) * stat.size(), cudaMemcpyDeviceToHost, ctx->stream));
        }
        MCEIK_CUDA(cudaStreamSynchronize(ctx->stream));
        int any = 0;
        for (int f = 0; f < nfields; ++f) {
            const int r = rk[f], pos = row[f] - r * slots;
            if (iters) iters[f] = stat[2 * (size_t)slots * r + pos];
            const int e = stat[2 * (size_t)slots * r + slots + pos];
            if (field_ierr) field_ierr[f] = e;
            any |= e;
            if (table_row) table_row[f] = row[f];
        }
        if (rc < 0) {
            set_error("mceik_fsm_solve_sharded_dev: the solve of rank %d failed: %s", rank, local_error.c_str());
            return rc;
        }
        return any ? 1 : 0;
    });
}

int mceik_catalog_misfit_dev(mceik_ctx *ctx, const float *d_tables, size_t ldgrd, int ngrd, int nmodels, int ntab, int nevents,
                             const int *d_node, const double *d_tobs, const double *d_varobs, const int *d_use, double *d_misfit) {
    return guarded([&]() -> int {
        if (!ctx || !d_tables || !d_node || !d_tobs || !d_varobs || !d_use || !d_misfit || nmodels < 0 || ntab < 1 || nevents < 0 ||
            ngrd < 1 || ldgrd < (size_t)ngrd) {
            set_error("mceik_catalog_misfit_dev: bad argument");
            return -1;
        }
        DeviceGuard dg(ctx->device);
        gs::launch_catalog_misfit(d_tables, ldgrd, nmodels, ntab, nevents, d_node, d_tobs, d_varobs, d_use, d_misfit, ctx->stream);
        return 0;
    });
}

int mceik_homogeneous_tables_dev(mceik_ctx *ctx, int nx, int ny, int nz, double x0, double y0, double z0, double dx,
                                 double dy, double dz, int nstations, const double *xs, const double *ys, const double *zs,
                                 const double *vel, float *d_tables, size_t ldtab) {
    return guarded([&]() -> int {
        if (!ctx || !xs || !ys || !zs || !vel || !d_tables || nx < 1 || ny < 1 || nz < 1 || nstations < 0 ||
            nstations > 65535 || ldtab < (size_t)nx * ny * nz) {
            set_error("mceik_homogeneous_tables_dev: bad argument");
            return -1;
        }
        DeviceGuard dg(ctx->device);
        std::vector<double> xyzv(4 * (size_t)nstations);
        for (int s = 0; s < nstations; ++s) {
            xyzv[4 * s] = xs[s]; xyzv[4 * s + 1] = ys[s]; xyzv[4 * s + 2] = zs[s];
            xyzv[4 * s + 3] = 1.0 / vel[s];  // homog.c:604
        }
        ctx->ws_xyzv.ensure(sizeof(double) * std::max<size_t>(xyzv.size(), 1));
        const double *d_xyzv = upload(ctx->ws_xyzv, 0, xyzv, ctx->stream);
        fsm::launch_homog_tables<float>(nx, ny, nz, x0, y0, z0, dx, dy, dz, nstations, d_xyzv, d_tables, ldtab, ctx->stream);
        MCEIK_CUDA(cudaStreamSynchronize(ctx->stream));
        return 0;
    });
}

// ---------------------------------------------------------------- locator state
int mceik_locate_set_tables_dev(mceik_ctx *ctx, int ntables, int ngrd, size_t ldgrd, const float *d_tables) {
    if (!ctx || ntables < 1 || ngrd < 1 || ldgrd < (size_t)ngrd || !d_tables) {
        set_error("mceik_locate_set_tables: bad argument");
        return -1;
    }
    ctx->d_tables = d_tables; ctx->ntables = ntables; ctx->ngrd = ngrd; ctx->ldgrd = ldgrd;
    return 0;
}

int mceik_locate_set_tables_host(mceik_ctx *ctx, int ntables, int ngrd, size_t ldgrd, const float *tables) {
    return guarded([&]() -> int {
        if (!ctx || ntables < 1 || ngrd < 1 || ldgrd < (size_t)ngrd || !tables) {
            set_error("mceik_locate_set_tables: bad argument");
            return -1;
        }
        DeviceGuard dg(ctx->device);
        float *d = static_cast<float *>(ctx->own_tables.ensure(sizeof(float) * ldgrd * ntables));
        MCEIK_CUDA(cudaMemcpyAsync(d, tables, sizeof(float) * ldgrd * ntables, cudaMemcpyHostToDevice, ctx->stream));
        MCEIK_CUDA(cudaStreamSynchronize(ctx->stream));
        return mceik_locate_set_tables_dev(ctx, ntables, ngrd, ldgrd, d);
    });
}

int mceik_locate_set_grid(mceik_ctx *ctx, int ngrd, const float *xl, const float *yl, const float *zl) {
    if (!ctx || ngrd < 1 || !xl || !yl || !zl) { set_error("mceik_locate_set_grid: bad argument"); return -1; }
    ctx->xlocs.assign(xl, xl + ngrd); ctx->ylocs.assign(yl, yl + ngrd); ctx->zlocs.assign(zl, zl + ngrd);
    return 0;
}

int mceik_locate3d_set_tables(int ntables, int ngrd, size_t ldgrd, const float *tables) {
    mceik_ctx *c = default_ctx();
    return c ? mceik_locate_set_tables_host(c, ntables, ngrd, ldgrd, tables) : -3;
}
int mceik_locate3d_set_grid(int ngrd, const float *xl, const float *yl, const float *zl) {
    mceik_ctx *c = default_ctx();
    return c ? mceik_locate_set_grid(c, ngrd, xl, yl, zl) : -3;
}

int mceik_locate_batched_dev(mceik_ctx *ctx, int job, int nevents, int nobs_total, int max_picks, const int *d_obs_ptr,
                             const int *d_table_id, const double *d_tobs_cor, const double *d_varobs,
                             const double *d_tori, int *d_iopt, double *d_t0opt, double *d_objopt) {
    return guarded([&]() -> int {
        return locate_dev(ctx, job, nevents, nobs_total, max_picks, d_obs_ptr, d_table_id, d_tobs_cor, d_varobs, d_tori,
                          d_iopt, d_t0opt, d_objopt);
    });
}

static int locate_host_arrays(mceik_ctx *ctx, int job, int nevents, const std::vector<int> &optr, int np, int maxp,
                              const int *table_id, const double *tobs_cor, const double *varobs, const double *tori, int *iopt,
                              double *t0opt, double *objopt);

int mceik_locate_batched_host(mceik_ctx *ctx, int job, int nevents, const int *obs_ptr, const int *table_id,
                              const double *tobs_cor, const double *varobs, const double *tori, int *iopt,
                              double *t0opt, double *objopt) {
    return guarded([&]() -> int {
        if (!ctx || nevents < 0 || !obs_ptr || !iopt || !t0opt || !objopt) {
            set_error("mceik_locate_batched_host: bad argument");
            return -1;
        }
        if (nevents == 0) return 0;
        if (job != 1 && job != 2) { set_error("mceik_locate_batched_host: job %d not supported", job); return 1; }
        const int np = obs_ptr[nevents] - obs_ptr[0];
        if (np > 0 && (!table_id || !tobs_cor || !varobs)) { set_error("mceik_locate_batched_host: NULL pick arrays"); return -1; }
        int maxp = 0;
        std::vector<int> optr(nevents + 1);
        for (int e = 0; e <= nevents; ++e) optr[e] = obs_ptr[e] - obs_ptr[0];
        for (int e = 0; e < nevents; ++e) {
            if (optr[e + 1] < optr[e]) { set_error("mceik_locate_batched_host: obs_ptr not monotone"); return -1; }
            maxp = std::max(maxp, optr[e + 1] - optr[e]);
        }
        for (int p = 0; p < np; ++p)
            if (table_id[obs_ptr[0] + p] >= ctx->ntables) { set_error("mceik_locate_batched_host: table id out of range"); return -1; }
        std::vector<int> optr2, tid2;
        std::vector<double> tobs2, var2;
        if (np > 0 && !ctx->tune.locate_no_align) {
            host::align_event_blocks(gs::kEventsPerBlock, nevents, optr.data(), table_id + obs_ptr[0], tobs_cor + obs_ptr[0], varobs + obs_ptr[0], optr2, tid2,
                               tobs2, var2);
            if (tid2.size() != (size_t)np || optr2 != optr) {  // some block was re-laid out: continue with the aligned picks
                optr.swap(optr2);
                obs_ptr = optr.data();
                table_id = tid2.data(); tobs_cor = tobs2.data(); varobs = var2.data();
                maxp = 0;
                for (int e = 0; e < nevents; ++e) maxp = std::max(maxp, optr[e + 1] - optr[e]);
                return locate_host_arrays(ctx, job, nevents, optr, (int)tid2.size(), maxp, table_id, tobs_cor, varobs, tori, iopt,
                                          t0opt, objopt);
            }
        }
        return locate_host_arrays(ctx, job, nevents, optr, np, maxp, table_id + obs_ptr[0], tobs_cor + obs_ptr[0],
                                  varobs + obs_ptr[0], tori, iopt, t0opt, objopt);
    });
}

// picks [0, np) with optr[0] == 0 on the host -> device -> located events back
static int locate_host_arrays(mceik_ctx *ctx, int job, int nevents, const std::vector<int> &optr, int np, int maxp,
                              const int *table_id, const double *tobs_cor, const double *varobs, const double *tori, int *iopt,
                              double *t0opt, double *objopt) {
    {
        DeviceGuard dg(ctx->device);
        cudaStream_t st = ctx->stream;
        const size_t npp = std::max(np, 1);
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
        const size_t o_ptr = take(sizeof(int) * (nevents + 1)), o_tid = take(sizeof(int) * npp);
        const size_t o_tobs = take(sizeof(double) * npp), o_var = take(sizeof(double) * npp), o_tori = take(sizeof(double) * nevents);
        char *in = static_cast<char *>(ctx->ws_gs_in.ensure(off));
        MCEIK_CUDA(cudaMemcpyAsync(in + o_ptr, optr.data(), sizeof(int) * (nevents + 1), cudaMemcpyHostToDevice, st));
        if (np > 0) {
            MCEIK_CUDA(cudaMemcpyAsync(in + o_tid, table_id, sizeof(int) * np, cudaMemcpyHostToDevice, st));
            MCEIK_CUDA(cudaMemcpyAsync(in + o_tobs, tobs_cor, sizeof(double) * np, cudaMemcpyHostToDevice, st));
            MCEIK_CUDA(cudaMemcpyAsync(in + o_var, varobs, sizeof(double) * np, cudaMemcpyHostToDevice, st));
        }
        if (tori) MCEIK_CUDA(cudaMemcpyAsync(in + o_tori, tori, sizeof(double) * nevents, cudaMemcpyHostToDevice, st));
        else if (job == 1) { set_error("mceik_locate_batched_host: job 1 needs tori"); return -1; }
        const size_t q_t0 = align_up(sizeof(int) * nevents), q_obj = q_t0 + align_up(sizeof(double) * nevents);
        char *out = static_cast<char *>(ctx->ws_gs_out.ensure(q_obj + sizeof(double) * nevents));
        const int rc = locate_dev(ctx, job, nevents, np, maxp, reinterpret_cast<int *>(in + o_ptr),
                                  reinterpret_cast<int *>(in + o_tid), reinterpret_cast<double *>(in + o_tobs),
                                  reinterpret_cast<double *>(in + o_var), reinterpret_cast<double *>(in + o_tori),
                                  reinterpret_cast<int *>(out), reinterpret_cast<double *>(out + q_t0),
                                  reinterpret_cast<double *>(out + q_obj));
        if (rc != 0) return rc;
        MCEIK_CUDA(cudaMemcpyAsync(iopt, out, sizeof(int) * nevents, cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaMemcpyAsync(t0opt, out + q_t0, sizeof(double) * nevents, cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaMemcpyAsync(objopt, out + q_obj, sizeof(double) * nevents, cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
}

int mceik_locate_event_logpdf_host(mceik_ctx *ctx, int job, int npicks, const int *table_id, const double *tobs_cor,
                                   const double *varobs, double tori, double *logpdf, float *logpdf4, double *t0grid) {
    return guarded([&]() -> int {
        if (!ctx || !ctx->d_tables) { set_error("mceik_locate_event_logpdf_host: no travel-time tables set"); return -1; }
        if (job != 1 && job != 2) { set_error("mceik_locate_event_logpdf_host: job %d not supported", job); return 1; }
        if (npicks < 1 || !table_id || !tobs_cor || !varobs) { set_error("mceik_locate_event_logpdf_host: bad argument"); return -1; }
        for (int p = 0; p < npicks; ++p)
            if (table_id[p] >= ctx->ntables) { set_error("mceik_locate_event_logpdf_host: table id out of range"); return -1; }
        DeviceGuard dg(ctx->device);
        cudaStream_t st = ctx->stream;
        const size_t ng = (size_t)ctx->ngrd;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
        const size_t o_ptr = take(sizeof(int) * 2), o_tid = take(sizeof(int) * npicks), o_tobs = take(sizeof(double) * npicks);
        const size_t o_var = take(sizeof(double) * npicks), o_w0 = take(sizeof(double) * npicks), o_w1 = take(sizeof(double) * npicks);
        const size_t o_nuse = take(sizeof(int)), o_pdf = take(sizeof(double) * ng), o_pdf4 = take(sizeof(float) * ng), o_t0 = take(sizeof(double) * ng);
        char *b = static_cast<char *>(ctx->ws_gs_misc.ensure(off));
        const int ptr[2] = {0, npicks};
        MCEIK_CUDA(cudaMemcpyAsync(b + o_ptr, ptr, sizeof(ptr), cudaMemcpyHostToDevice, st));
        MCEIK_CUDA(cudaMemcpyAsync(b + o_tid, table_id, sizeof(int) * npicks, cudaMemcpyHostToDevice, st));
        MCEIK_CUDA(cudaMemcpyAsync(b + o_tobs, tobs_cor, sizeof(double) * npicks, cudaMemcpyHostToDevice, st));
        MCEIK_CUDA(cudaMemcpyAsync(b + o_var, varobs, sizeof(double) * npicks, cudaMemcpyHostToDevice, st));
        gs::launch_prepare(1, reinterpret_cast<int *>(b + o_ptr), reinterpret_cast<int *>(b + o_tid),
                           reinterpret_cast<double *>(b + o_var), reinterpret_cast<double *>(b + o_w0),
                           reinterpret_cast<double *>(b + o_w1), reinterpret_cast<int *>(b + o_nuse), st);
        gs::launch_event_grid(ctx->ngrd, ctx->ldgrd, npicks, reinterpret_cast<int *>(b + o_tid), reinterpret_cast<double *>(b + o_tobs),
                              reinterpret_cast<double *>(b + o_w0), reinterpret_cast<double *>(b + o_w1), job == 2, tori,
                              ctx->d_tables, logpdf ? reinterpret_cast<double *>(b + o_pdf) : nullptr,
                              logpdf4 ? reinterpret_cast<float *>(b + o_pdf4) : nullptr,
                              t0grid ? reinterpret_cast<double *>(b + o_t0) : nullptr, st);
        if (logpdf) MCEIK_CUDA(cudaMemcpyAsync(logpdf, b + o_pdf, sizeof(double) * ng, cudaMemcpyDeviceToHost, st));
        if (logpdf4) MCEIK_CUDA(cudaMemcpyAsync(logpdf4, b + o_pdf4, sizeof(float) * ng, cudaMemcpyDeviceToHost, st));
        if (t0grid) MCEIK_CUDA(cudaMemcpyAsync(t0grid, b + o_t0, sizeof(double) * ng, cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaStreamSynchronize(st));
        return 0;
    });
}

int mceik_locate_optnode_host(mceik_ctx *ctx, int ngrd, const double *pdf, int *node) {
    return guarded([&]() -> int {
        if (!ctx || ngrd < 1 || !pdf || !node) { set_error("mceik_locate_optnode_host: bad argument"); return -1; }
        DeviceGuard dg(ctx->device);
        cudaStream_t st = ctx->stream;
        const size_t o_s = align_up(sizeof(double) * (size_t)ngrd), o_out = o_s + align_up(gs::minloc_scratch_bytes());
        char *b = static_cast<char *>(ctx->ws_gs_misc.ensure(o_out + 256));
        MCEIK_CUDA(cudaMemcpyAsync(b, pdf, sizeof(double) * (size_t)ngrd, cudaMemcpyHostToDevice, st));
        gs::launch_minloc<double>(ngrd, reinterpret_cast<double *>(b), reinterpret_cast<int *>(b + o_out), b + o_s,
                                  gs::minloc_scratch_bytes(), st, /*maxloc=*/true);
        MCEIK_CUDA(cudaMemcpyAsync(node, b + o_out, sizeof(int), cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaStreamSynchronize(st));
        return 0;
    });
}

int mceik_locate_normalize_pdf_host(mceik_ctx *ctx, int ngrd, double *pdf, double *sum) {
    return guarded([&]() -> int {
        if (!ctx || ngrd < 1 || !pdf) { set_error("mceik_locate_normalize_pdf_host: bad argument"); return -1; }
        DeviceGuard dg(ctx->device);
        cudaStream_t st = ctx->stream;
        const size_t o_s = align_up(sizeof(double) * (size_t)ngrd), o_out = o_s + align_up(gs::minloc_scratch_bytes());
        char *b = static_cast<char *>(ctx->ws_gs_misc.ensure(o_out + 256));
        double xsum = 0.0;
        MCEIK_CUDA(cudaMemcpyAsync(b, pdf, sizeof(double) * (size_t)ngrd, cudaMemcpyHostToDevice, st));
        gs::launch_sum(ngrd, reinterpret_cast<double *>(b), reinterpret_cast<double *>(b + o_out), b + o_s,
                       gs::minloc_scratch_bytes(), st);
        MCEIK_CUDA(cudaMemcpyAsync(&xsum, b + o_out, sizeof(double), cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaStreamSynchronize(st));
        if (sum) *sum = xsum;
        if (xsum == 0.0) {  // locate.f90:56-60
            printf(" locate_optloc: Division by zero\n");
            set_error("mceik_locate_normalize_pdf_host: the PDF sums to zero");
            return 1;
        }
        gs::launch_scale(ngrd, 1.0 / xsum, reinterpret_cast<double *>(b), st);  // xsumi = one/xsum; DSCAL (locate.f90:61-62)
        MCEIK_CUDA(cudaMemcpyAsync(pdf, b, sizeof(double) * (size_t)ngrd, cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaStreamSynchronize(st));
        return 0;
    });
}

static void fill_hypo(const mceik_ctx *ctx, int nevents, const int *iopt, const double *t0, double *hypo) {
    for (int e = 0; e < nevents; ++e) {
        const int i = iopt[e];
        if (i < 0 || (size_t)i >= ctx->xlocs.size()) {
            hypo[4 * e] = hypo[4 * e + 1] = hypo[4 * e + 2] = hypo[4 * e + 3] = 0.0;
        } else {  // hypo4 = (xlocs, ylocs, zlocs, t0)(iopt), locate.f90:478-481
            hypo[4 * e] = (double)ctx->xlocs[i]; hypo[4 * e + 1] = (double)ctx->ylocs[i];
            hypo[4 * e + 2] = (double)ctx->zlocs[i]; hypo[4 * e + 3] = t0[e];
        }
    }
}

int mceik_locate_catalog(mceik_ctx *ctx, const struct mceik_catalog_struct *cat, const struct mceik_stations_struct *sta,
                         int job, double *hypo, int *iopt_out, double *obj_out) {
    return guarded([&]() -> int {
        if (!ctx || !cat || !sta || !hypo || cat->nevents < 0 || !cat->obsPtr) {
            set_error("mceik_locate_catalog: bad argument");
            return -1;
        }
        if ((int)ctx->xlocs.size() != ctx->ngrd) { set_error("mceik_locate_catalog: node coordinates not set"); return -1; }
        const int ne = cat->nevents, np = cat->obsPtr[ne] - cat->obsPtr[0];
        std::vector<int> tid(std::max(np, 1));
        std::vector<double> tc(std::max(np, 1));
        for (int p = cat->obsPtr[0]; p < cat->obsPtr[ne]; ++p) {
            const int q = p - cat->obsPtr[0], stn = cat->statPtr[p], ph = cat->pickType[p];
            if (cat->luseObs[p] == 0) { tid[q] = -1; tc[q] = 0.0; continue; }
            if (stn < 1 || stn > sta->nstat || (ph != P_PRIMARY_PICK && ph != S_PRIMARY_PICK)) {
                set_error("mceik_locate_catalog: pick %d has station %d / phase %d", p, stn, ph);
                return -1;
            }
            tid[q] = 2 * (stn - 1) + (ph - 1);
            const double cor = ph == P_PRIMARY_PICK ? (sta->pcorr ? sta->pcorr[stn - 1] : 0.0) : (sta->scorr ? sta->scorr[stn - 1] : 0.0);
            tc[q] = cat->tobs[p] - cor;
        }
        std::vector<int> optr(ne + 1), iopt(std::max(ne, 1));
        for (int e = 0; e <= ne; ++e) optr[e] = cat->obsPtr[e] - cat->obsPtr[0];
        std::vector<double> t0(std::max(ne, 1)), obj(std::max(ne, 1));
        const int rc = mceik_locate_batched_host(ctx, job, ne, optr.data(), tid.data(), tc.data(),
                                                 cat->varObs + cat->obsPtr[0], cat->tori, iopt.data(), t0.data(), obj.data());
        if (rc != 0) return rc;
        fill_hypo(ctx, ne, iopt.data(), t0.data(), hypo);
        for (int e = 0; e < ne; ++e) {
            if (iopt_out) iopt_out[e] = iopt[e];
            if (obj_out) obj_out[e] = obj[e];
        }
        return 0;
    });
}

// ============================================================================================
// DROP-IN SYMBOLS
// ============================================================================================
void eikonal3d_serial_driver(const int *job, const int *iverb, const int *maxit, const int *nsrc, const int *nx,
                             const int *ny, const int *nz, const double *tol, const double *h, const double *x0,
                             const double *y0, const double *z0, const double *ts, const double *xs, const double *ys,
                             const double *zs, const double *slow, double *u, int *ierr) {
    *ierr = 0;
    if (*job == 1) {
        if (g_serial.init) {
            printf(" eikonal3d_serial_driver: Already initialized!\n");
            *ierr = 1;
            return;
        }
        mceik_ctx *c = default_ctx();
        if (!c) { printf(" eikonal3d_serial_driver: %s\n", mceik_last_error()); *ierr = 1; return; }
        if (*iverb > 0) printf(" eikonal3d_serial_driver: Generating levels...\n");
        g_serial.init = true; g_serial.nx = *nx; g_serial.ny = *ny; g_serial.nz = *nz;
    } else if (*job == 2) {
        if (!g_serial.init) {
            printf(" eikonal3d_serial_driver: Solver not initalized!\n");
            *ierr = 1;
            return;
        }
        mceik_ctx *c = default_ctx();
        if (!c) { *ierr = 1; return; }
        if (*iverb > 0) printf(" eikonal3d_serial_driver: Solving...\n");
        mceik_fsm_grid g{*nx, *ny, *nz, *h, *x0, *y0, *z0, *tol, *maxit};
        const int fm = 0, sp[2] = {0, *nsrc};
        int ferr = 0;
        const int rc = mceik_fsm_solve_batched_host(c, &g, 1, slow, 1, &fm, sp, ts, xs, ys, zs, u, nullptr, 0, nullptr, &ferr);
        if (rc != 0) {
            if (ferr) printf(" eikonal3d_serial_driver: Error setting boundary conditions\n");
            else printf(" eikonal3d_serial_driver: %s\n", mceik_last_error());
            *ierr = 1;
        }
    } else {
        if (!g_serial.init) printf(" eikonal3d_serial_driver: Never initialized!\n");
        g_serial = SerialState();
    }
}

void eikonal3d_initialize(const int *comm, const int *iverb, const int *nx, const int *ny, const int *nz, const int *ndivx,
                          const int *ndivy, const int *ndivz, const int *noverlap, const int *maxit, const double *x0,
                          const double *y0, const double *z0, const double *h, const double *tol, int *ierr) {
    (void)comm; (void)iverb; (void)ndivx; (void)ndivy; (void)ndivz; (void)noverlap;
    *ierr = 0;
    if (!default_ctx()) { printf(" eikonal3d_initialize: %s\n", mceik_last_error()); *ierr = 1; return; }
    g_solve.grid = mceik_fsm_grid{*nx, *ny, *nz, *h, *x0, *y0, *z0, *tol, *maxit};
    g_solve.init = true;
}

void eikonal3d_solve(const int *comm, const int *nsrc, const int *n, const double *ts, const double *xs, const double *ys,
                     const double *zs, const double *slow, double *u, int *ierr) {
    (void)comm;
    *ierr = 0;
    mceik_ctx *c = g_solve.init ? default_ctx() : nullptr;
    const mceik_fsm_grid &g = g_solve.grid;
    if (!c || (long long)*n != (long long)g.nx * g.ny * g.nz) {
        printf(" eikonal3d_solve: solver not initialized or n /= nx*ny*nz\n");
        *ierr = 1;
        return;
    }
    // slowness must be positive everywhere (the reference rejects MINVAL(slow) == 0, fsm3d.f90:1802)
    for (long long i = 0; i < *n; ++i)
        if (slow[i] == 0.0) { printf(" eikonal3d_model: Error scattering model!\n"); *ierr = 1; return; }
    const int fm = 0, sp[2] = {0, *nsrc};
    int ferr = 0;
    if (mceik_fsm_solve_batched_host(c, &g, 1, slow, 1, &fm, sp, ts, xs, ys, zs, u, nullptr, 0, nullptr, &ferr) != 0) {
        printf(" eikonal3d_solve: %s\n", ferr ? "Error setting bcs" : mceik_last_error());
        *ierr = 1;
    }
}

void eikonal3d_finalize(const int *comm, int *ierr) {
    (void)comm;
    *ierr = 0;
    if (!g_solve.init) {
        printf(" eikonal3d_finalize: Solver was never initialized\n");
        *ierr = 1;
    }
    g_solve = SolveState();
}

}  // extern "C"

// ---- single-event full-grid searches --------------------------------------------------------
namespace {
template <typename T>
int full_grid_host(int ldgrd, int ngrd, const std::vector<int> &rows, const std::vector<T> &tobs, const std::vector<T> &w0,
                   const std::vector<T> &w1, int want_ot, T t0use, const T *test, T *t0, T *obj) {
    mceik_ctx *c = default_ctx();
    if (!c) return 1;
    return guarded([&]() -> int {
        DeviceGuard dg(c->device);
        cudaStream_t st = c->stream;
        const int nuse = (int)rows.size();
        const size_t row_bytes = sizeof(T) * (size_t)ngrd, ld = (size_t)ldgrd;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
        const size_t o_test = take(sizeof(T) * ld * std::max(nuse, 1)), o_t0 = take(row_bytes), o_obj = take(row_bytes);
        const size_t o_rows = take(sizeof(int) * std::max(nuse, 1)), o_tobs = take(sizeof(T) * std::max(nuse, 1));
        const size_t o_w0 = take(sizeof(T) * std::max(nuse, 1)), o_w1 = take(sizeof(T) * std::max(nuse, 1));
        char *b = static_cast<char *>(c->ws_gs_misc.ensure(off));
        std::vector<int> compact(nuse);
        for (int j = 0; j < nuse; ++j) {  // only the used rows travel, packed in pick order
            compact[j] = j;
            MCEIK_CUDA(cudaMemcpyAsync(b + o_test + sizeof(T) * ld * j, test + ld * (size_t)rows[j], row_bytes,
                                       cudaMemcpyHostToDevice, st));
        }
        if (nuse) {
            MCEIK_CUDA(cudaMemcpyAsync(b + o_rows, compact.data(), sizeof(int) * nuse, cudaMemcpyHostToDevice, st));
            MCEIK_CUDA(cudaMemcpyAsync(b + o_tobs, tobs.data(), sizeof(T) * nuse, cudaMemcpyHostToDevice, st));
            MCEIK_CUDA(cudaMemcpyAsync(b + o_w0, w0.data(), sizeof(T) * nuse, cudaMemcpyHostToDevice, st));
            MCEIK_CUDA(cudaMemcpyAsync(b + o_w1, w1.data(), sizeof(T) * nuse, cudaMemcpyHostToDevice, st));
        }
        gs::launch_full_grid<T>(ngrd, ld, nuse, reinterpret_cast<int *>(b + o_rows), reinterpret_cast<T *>(b + o_tobs),
                                reinterpret_cast<T *>(b + o_w0), reinterpret_cast<T *>(b + o_w1), want_ot, t0use,
                                reinterpret_cast<T *>(b + o_test), reinterpret_cast<T *>(b + o_t0),
                                reinterpret_cast<T *>(b + o_obj), st);
        if (t0) MCEIK_CUDA(cudaMemcpyAsync(t0, b + o_t0, row_bytes, cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaMemcpyAsync(obj, b + o_obj, row_bytes, cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaStreamSynchronize(st));
        return 0;
    }) == 0 ? 0 : 1;
}

template <typename T>
int l2_gridsearch_c(const char *fcnm, int ldgrd, int ngrd, int nobs, int iwantOT, T t0use, const int *mask, const T *tobs,
                    const T *tcorr, const T *varobs, const T *test, T *t0, T *objfn, T sqrt2i) {
    // argument checks of locate.c:948-974 (and :1104-1130 for float)
    if ((sizeof(T) * (size_t)ldgrd) % 64 != 0 || ldgrd < ngrd || nobs < 1 || !mask || !tobs || !varobs || !test || !t0 || !objfn) {
        if ((sizeof(T) * (size_t)ldgrd) % 64 != 0) printf("%s: Error ldgrd must be divisible by 64\n", fcnm);
        if (ldgrd < ngrd) printf("%s: Error ldgrd < ngrd\n", fcnm);
        if (!mask) printf("%s: mask is null\n", fcnm);
        if (!tobs) printf("%s: tobs is null\n", fcnm);
        if (!varobs) printf("%s: varobs is null\n", fcnm);
        if (!test) printf("%s: test is null\n", fcnm);
        if (!t0) printf("%s: t0 is null\n", fcnm);
        if (!objfn) printf("%s: objfn is null\n", fcnm);
        return 1;
    }
    if ((uintptr_t)t0 % 64 || (uintptr_t)test % 64 || (uintptr_t)objfn % 64) {
        printf("%s: Input arrays are not 64 bit aligned\n", fcnm);
        return 1;
    }
    // compress the unmasked picks (locate.c:981-1013); scalar work, same fp operations in T
    std::vector<int> rows;
    std::vector<T> tc, wt;
    T xnorm = (T)0;
    for (int i = 0; i < nobs; ++i) {
        if (mask[i] != 0) continue;
        tc.push_back(tcorr ? tobs[i] - tcorr[i] : tobs[i]);
        wt.push_back((T)1 / varobs[i]);
        xnorm = xnorm + wt.back();
        rows.push_back(i);
    }
    std::vector<T> w0(wt.size()), w1(wt.size());
    for (size_t j = 0; j < wt.size(); ++j) {
        w0[j] = wt[j] / xnorm;   // locate.c:399
        w1[j] = wt[j] * sqrt2i;  // locate.c:500
    }
    if (ngrd == 0) return 0;
    if (full_grid_host<T>(ldgrd, ngrd, rows, tc, w0, w1, iwantOT == 1, t0use, test, t0, objfn) != 0) {
        printf("%s: %s\n", fcnm, mceik_last_error());
        return 1;
    }
    return 0;
}

// Host-side weighted median in the form locate.c:73 declares (the reference defines it nowhere):
// perm is a candidate ordering kept by the caller between calls; *lsort tells whether it had to be redone.
double weighted_median_host(int n, const double *x, const double *w, int *perm, bool *lsort, int *ierr) {
    if (ierr) *ierr = 0;
    if (lsort) *lsort = false;
    if (n < 1 || !x || !w) {
        if (ierr) *ierr = 1;
        return 0.0;
    }
    std::vector<int> own;
    if (!perm) {
        own.resize(n);
        for (int i = 0; i < n; ++i) own[i] = i;
        perm = own.data();
    }
    bool ok = true;
    std::vector<char> seen(n, 0);
    for (int i = 0; i < n && ok; ++i) {
        if (perm[i] < 0 || perm[i] >= n || seen[perm[i]]) ok = false;
        else seen[perm[i]] = 1;
    }
    auto before = [&](int a, int b) { return x[a] < x[b] || (x[a] == x[b] && a < b); };
    for (int i = 1; i < n && ok; ++i)
        if (!before(perm[i - 1], perm[i])) ok = false;
    if (!ok) {
        for (int i = 0; i < n; ++i) perm[i] = i;
        std::sort(perm, perm + n, before);
        if (lsort) *lsort = true;
    }
    double W = 0.0;
    for (int k = 0; k < n; ++k) W = W + w[perm[k]];
    const double half = 0.5 * W;
    double cum = 0.0;
    for (int k = 0; k < n; ++k) {
        cum = cum + w[perm[k]];
        if (cum > half) return x[perm[k]];
        if (cum == half) return k + 1 < n ? 0.5 * (x[perm[k]] + x[perm[k + 1]]) : x[perm[k]];
    }
    return x[perm[n - 1]];
}

int l1_gridsearch_c(int ldgrd, int ngrd, int nobs, int iwantOT, double t0use, const int *mask, const double *tobs,
                    const double *varobs, const double *test, double *t0, double *objfn) {
    const char *fcnm = "locate_l1_gridSearch__double64";
    if (ldgrd < ngrd || ngrd < 0 || nobs < 1 || !mask || !tobs || !varobs || !test || !t0 || !objfn) {
        printf("%s: Error invalid argument\n", fcnm);
        return 1;
    }
    // locate.c:1236-1246 (weights of the unmasked picks, packed) and :1263-1273 (normalisation, analytic-t0 branch only)
    std::vector<int> rows;
    std::vector<double> tc, wt;
    double wtsum = 0.0;
    for (int i = 0; i < nobs; ++i) {
        if (mask[i] != 0) continue;
        wt.push_back(1.0 / varobs[i]);
        tc.push_back(tobs[i]);
        rows.push_back(i);
        wtsum = wtsum + wt.back();
    }
    if (iwantOT == 1 && fabs(wtsum - 1.0) > 1.e-14) {
        const double wtsumi = 1.0 / wtsum;
        for (double &w : wt) w = w * wtsumi;
    }
    if (ngrd == 0) return 0;
    if ((int)rows.size() > gs::kL1MaxObs) {
        printf("%s: Error more than %d used observations\n", fcnm, gs::kL1MaxObs);
        return 1;
    }
    mceik_ctx *c = default_ctx();
    if (!c) return 1;
    const int rc = guarded([&]() -> int {
        DeviceGuard dg(c->device);
        cudaStream_t st = c->stream;
        const int nuse = (int)rows.size();
        const size_t row_bytes = sizeof(double) * (size_t)ngrd, ld = (size_t)ldgrd;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
        const size_t o_test = take(sizeof(double) * ld * std::max(nuse, 1)), o_t0 = take(row_bytes), o_obj = take(row_bytes);
        const size_t o_tobs = take(sizeof(double) * std::max(nuse, 1)), o_wt = take(sizeof(double) * std::max(nuse, 1));
        char *b = static_cast<char *>(c->ws_gs_misc.ensure(off));
        for (int j = 0; j < nuse; ++j)
            MCEIK_CUDA(cudaMemcpyAsync(b + o_test + sizeof(double) * ld * j, test + ld * (size_t)rows[j], row_bytes,
                                       cudaMemcpyHostToDevice, st));
        if (nuse) {
            MCEIK_CUDA(cudaMemcpyAsync(b + o_tobs, tc.data(), sizeof(double) * nuse, cudaMemcpyHostToDevice, st));
            MCEIK_CUDA(cudaMemcpyAsync(b + o_wt, wt.data(), sizeof(double) * nuse, cudaMemcpyHostToDevice, st));
        }
        gs::launch_l1_grid(ngrd, ld, nuse, reinterpret_cast<double *>(b + o_tobs), reinterpret_cast<double *>(b + o_wt),
                           iwantOT == 1, t0use, reinterpret_cast<double *>(b + o_test), reinterpret_cast<double *>(b + o_t0),
                           reinterpret_cast<double *>(b + o_obj), st);
        MCEIK_CUDA(cudaMemcpyAsync(t0, b + o_t0, row_bytes, cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaMemcpyAsync(objfn, b + o_obj, row_bytes, cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaStreamSynchronize(st));
        return 0;
    });
    if (rc != 0) printf("%s: %s\n", fcnm, mceik_last_error());
    return rc == 0 ? 0 : 1;
}

template <typename T>
void gridsearch_f90(const char *fcnm, int ldgrd, int ngrd, int nobs, int iwantOT, const int *mask, const T *tobs,
                    const T *varobs, const T *test, T *logPDF, int *ierr, T eps) {
    *ierr = 0;
    if (ldgrd % 64 != 0) { printf(" %s: Require arrays be 64 byte aligned\n", fcnm); *ierr = 1; return; }
    if (ngrd > ldgrd) { printf(" %s: ngrd cannot be greater than ldgrd\n", fcnm); *ierr = 1; return; }
    int msum = 0;
    T vsum = (T)0;
    for (int i = 0; i < nobs; ++i) { msum += mask[i]; vsum = vsum + varobs[i]; }
    if (msum == nobs) { printf(" %s: No observations\n", fcnm); *ierr = 1; return; }
    if (std::fabs(vsum - (T)0) < eps) { printf(" %s: Will be division by zero\n", fcnm); *ierr = 1; return; }
    T xnorm = (T)0;  // sum of the variances of the unmasked picks (gridsearch.f90:431-434)
    for (int i = 0; i < nobs; ++i)
        if (mask[i] != 1) xnorm = xnorm + varobs[i];
    const T sqrt2i = (T)1 / std::sqrt((T)2);
    std::vector<int> rows;
    std::vector<T> tc, w0, w1;
    for (int i = 0; i < nobs; ++i) {
        if (mask[i] == 1) continue;
        rows.push_back(i);
        tc.push_back(tobs[i]);
        w0.push_back((T)1 / (varobs[i] * xnorm));  // gridsearch.f90:185
        w1.push_back(sqrt2i / varobs[i]);          // gridsearch.f90:273
    }
    if (ngrd == 0) return;
    if (full_grid_host<T>(ldgrd, ngrd, rows, tc, w0, w1, iwantOT == 1, (T)0, test, (T *)nullptr, logPDF) != 0) {
        printf(" %s: %s\n", fcnm, mceik_last_error());
        *ierr = 1;
    }
}

template <typename T>
int minloc_host(int n, const T *x) {
    mceik_ctx *c = default_ctx();
    if (!c || n < 1 || !x) return 0;
    int result = 0;
    guarded([&]() -> int {
        DeviceGuard dg(c->device);
        cudaStream_t st = c->stream;
        const size_t o_s = align_up(sizeof(T) * (size_t)n), o_out = o_s + align_up(gs::minloc_scratch_bytes());
        char *b = static_cast<char *>(c->ws_gs_misc.ensure(o_out + 256));
        MCEIK_CUDA(cudaMemcpyAsync(b, x, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, st));
        gs::launch_minloc<T>(n, reinterpret_cast<T *>(b), reinterpret_cast<int *>(b + o_out), b + o_s,
                             gs::minloc_scratch_bytes(), st);
        MCEIK_CUDA(cudaMemcpyAsync(&result, b + o_out, sizeof(int), cudaMemcpyDeviceToHost, st));
        MCEIK_CUDA(cudaStreamSynchronize(st));
        return 0;
    });
    return result;
}
}  // namespace

extern "C" {

int locate_l2_gridSearch__double64(const int ldgrd, const int ngrd, const int nobs, const int iwantOT, const double t0use,
                                   const int *mask, const double *tobs, const double *tcorr, const double *varobs,
                                   const double *test, double *t0, double *objfn) {
    return l2_gridsearch_c<double>("locate_l2_gridSearch__double64", ldgrd, ngrd, nobs, iwantOT, t0use, mask, tobs, tcorr,
                                   varobs, test, t0, objfn, 0.7071067811865475);
}
int locate_l2_gridSearch__float64(const int ldgrd, const int ngrd, const int nobs, const int iwantOT, const float t0use,
                                  const int *mask, const float *tobs, const float *tcorr, const float *varobs,
                                  const float *test, float *t0, float *objfn) {
    return l2_gridsearch_c<float>("locate_l2_gridSearch__float64", ldgrd, ngrd, nobs, iwantOT, t0use, mask, tobs, tcorr,
                                  varobs, test, t0, objfn, 0.7071067811865475f);
}
int locate_l1_gridSearch__double64(const int ldgrd, const int ngrd, const int nobs, const int iwantOT, const double t0use,
                                   const int *mask, const double *tobs, const double *varobs, const double *test,
                                   double *t0, double *objfn) {
    return l1_gridsearch_c(ldgrd, ngrd, nobs, iwantOT, t0use, mask, tobs, varobs, test, t0, objfn);
}
double weightedMedian__double(const int n, const double *x, const double *w, int *perm, bool *lsort, int *ierr) {
    return weighted_median_host(n, x, w, perm, lsort, ierr);
}
int locate_minLocDouble64(const int n, const double *x) { return minloc_host<double>(n, x); }
int locate_minLocFloat64(const int n, const float *x) { return minloc_host<float>(n, x); }

void locate3d_gridsearch__double64(const int *ldgrd, const int *ngrd, const int *nobs, const int *iwantOT, const int *mask,
                                   const double *tobs, const double *varobs, const double *test, double *logPDF, int *ierr) {
    gridsearch_f90<double>("locate3d_gridsearch_double64", *ldgrd, *ngrd, *nobs, *iwantOT, mask, tobs, varobs, test, logPDF,
                           ierr, 2.220446049250313e-16);
}
void locate3d_gridsearch__float64(const int *ldgrd, const int *ngrd, const int *nobs, const int *iwantOT, const int *mask,
                                  const float *tobs, const float *varobs, const float *test, float *logPDF, int *ierr) {
    gridsearch_f90<float>("locate3d_gridsearch_float64", *ldgrd, *ngrd, *nobs, *iwantOT, mask, tobs, varobs, test, logPDF,
                          ierr, 1.1920929e-07f);
}

// ---- catalogue locator ----------------------------------------------------------------------
void locate3d_initialize(const int *comm, const int *iverb, const long *tttFileID, const long *locFileID, const int *ndivx,
                         const int *ndivy, const int *ndivz, int *ierr) {
    (void)comm; (void)tttFileID; (void)locFileID; (void)ndivx; (void)ndivy; (void)ndivz;
    *ierr = 0;
    if (!default_ctx()) { printf(" locate3d_initialize: %s\n", mceik_last_error()); *ierr = 1; return; }
    g_loc.init = true;
    g_loc.iverb = *iverb;
}

void locate3d_gridsearch(const int *model, const int *job, const int *nobs, const int *nevents, const int *luseObs,
                         const int *statPtr, const int *pickType, const double *statCor, const double *tori,
                         const double *varobs, const double *tobs, double *test, double *hypo, int *ierr) {
    (void)model; (void)test;
    *ierr = 0;
    mceik_ctx *c = g_loc.init ? default_ctx() : nullptr;
    if (!c || !c->d_tables || (int)c->xlocs.size() != c->ngrd) {
        printf(" locate3d_gridsearch: locator not initialized (tables / node coordinates missing)\n");
        *ierr = 1;
        return;
    }
    if (*job != 1 && *job != 2) {  // locate.f90:501-515
        if (*job == 3 || *job == 5) printf(" Not yet done\n");
        else printf(" locate_gridsearch: Invalid job\n");
        *ierr = 1;
        return;
    }
    const int ne = *nevents, no = *nobs;
    std::vector<int> optr(ne + 1), tid((size_t)std::max(ne * no, 1)), iopt(std::max(ne, 1));
    std::vector<double> tc((size_t)std::max(ne * no, 1)), t0(std::max(ne, 1)), obj(std::max(ne, 1));
    for (int e = 0; e <= ne; ++e) optr[e] = e * no;
    for (int e = 0; e < ne; ++e)
        for (int i = 0; i < no; ++i) {
            const size_t p = (size_t)e * no + i;  // myobs = (isrc-1)*nobs + iobs, locate.f90:394
            tid[p] = luseObs[p] == 0 ? -1 : 2 * (statPtr[p] - 1) + (pickType[p] - 1);
            tc[p] = tobs[p] - statCor[i];         // locate.f90:409, 453
        }
    if (mceik_locate_batched_host(c, *job, ne, optr.data(), tid.data(), tc.data(), varobs, tori, iopt.data(), t0.data(),
                                  obj.data()) != 0) {
        printf(" locate3d_gridsearch: %s\n", mceik_last_error());
        *ierr = 1;
        return;
    }
    fill_hypo(c, ne, iopt.data(), t0.data(), hypo);
}

void locate3d_finalize(void) { g_loc = LocState(); }

// ---- analytic homogeneous table --------------------------------------------------------------
int computeHomogeneousTraveltimes(const int nx, const int ny, const int nz, double x0, double y0, double z0, const double dx,
                                  double dy, const double dz, const double xs, const double ys, double zs, const double vel,
                                  double *ttimes) {
    // distance * (1/vel) per node in fp64 on the device, as homog.c:605-619
    mceik_ctx *c = default_ctx();
    if (!c || !ttimes || nx < 1 || ny < 1 || nz < 1) return 1;
    return guarded([&]() -> int {
        DeviceGuard dg(c->device);
        const size_t N = (size_t)nx * ny * nz;
        const std::vector<double> xyzv = {xs, ys, zs, 1.0 / vel};
        c->ws_xyzv.ensure(sizeof(double) * 4);
        const double *d_xyzv = upload(c->ws_xyzv, 0, xyzv, c->stream);
        double *d_t = static_cast<double *>(c->ws_u.ensure(sizeof(double) * N));
        fsm::launch_homog_tables<double>(nx, ny, nz, x0, y0, z0, dx, dy, dz, 1, d_xyzv, d_t, N, c->stream);
        MCEIK_CUDA(cudaMemcpyAsync(ttimes, d_t, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
        MCEIK_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    }) == 0 ? 0 : 1;
}

}  // extern "C"
