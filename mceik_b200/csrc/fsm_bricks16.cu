// fsm_bricks16.cu -- the streaming brick sweep kernel for grids with nx % 8 == 0 (sm_100a): the default sweep path.
//
// Same algorithm, skewed ring and scheduling as sweep_bricks_kernel (fsm_bricks.cu -- read its header first: a warp
// walks one 8 x 8 x zc brick of one field for one sweep, one brick hyperplane per step); what differs is the data path:
//
//   * the fields live in a BLOCKED layout while they are solved (BrickArgs::blocked, fsm.cuh): [brick column][z][80]
//     = the 64 nodes of a brick plane (512 contiguous bytes) + copies of its columns 0 and 7.  A brick walk of the
//     caller's [z][y][x] layout reads 64-byte pieces at a 2 KB stride, which the DRAM serves at 4.4 TB/s at most, and
//     pays 8 sectors per side and plane for its x halo; here the walk is sequential and the x halo is the neighbour's
//     face copy.  The kernel keeps the copies up to date itself (face staging ring FB; a side's plane is stored in
//     halves, 32 bytes = one sector as soon as its four rows are final, so the downwind x neighbour follows four
//     steps closer).  The [z][y][x] layout remains available (blocked = 0; tuning key NATURAL).
//   * a step has ONE __syncwarp, at its top; the neighbour reads, the write-back of the slot that became final in
//     the previous step, the copies six slots ahead, the face store and the two update chains are one region the
//     compiler interleaves (tests/brick_pipeline_emulation.c models the schedule, blocked layout included).
//   * every global <-> shared transfer of nodes moves a PAIR of x-adjacent nodes (16 bytes): cp.async.cg (L2 only)
//     and 128-bit stores -- 2 copies + 1 halo copy (24 lanes) + 1 store per lane and step.
//   * a ring row keeps the brick's x order of MEMORY (column q = x - x_lo at cell q + 2 of a 10-cell row), whatever
//     the sweep direction, so pairs land in order; the sweep direction only decides which of the two x-neighbours is
//     "upwind" and which slot it lives in.  Rows (y) and planes (z) stay in sweep order.
//   * steps in which every transfer, update and store touches in-brick nodes run in whole publication chunks with all
//     global addresses formed as (per-task byte base) + (one running plane offset) and predicated halo copies;
//   * a warp remembers the progress of its upwind neighbours it has already observed and looks one chunk ahead;
//   * tickets interleave two groups of fields half a sweep apart (decode_ticket, fsm.cuh);
//   * with few active fields the kernel runs in its publisher flavour (template parameter): the last warp of
//     the CTA executes the gpu-scope fences and progress stores for the others.
//
// nx % 8 == 0 makes every brick full in x and every pair 16-byte aligned; other grids use sweep_bricks_kernel.
// What limits this kernel and everything that was tried: DESIGN.md section 4.1, profiles/kernel_evolution_r2.md.
#include <algorithm>
#include <type_traits>
#include <vector>
#include "fsm.cuh"
#include "fsm_solve.cuh"

namespace mceik {
namespace fsm {

namespace {

// timing experiments only (wrong results): -DMCEIK_DBG=bits, 1 no solver, 2 no progress waits / publications,
// 4 no global loads, 8 no global stores (profiles/kernel_evolution_r2.md)
#ifndef MCEIK_DBG
#define MCEIK_DBG 0
#endif
constexpr int kDbg = MCEIK_DBG;

constexpr int kBx = 8, kBy = 8, kNC = 2;
constexpr int kPrefetch = 2;
constexpr int kRing = 9 + kPrefetch;
// ring cell of node (row j, memory column q), j in [-1, By], q in [-1, Bx]: (j + 1) * kURow + q + 2 -- a row is 10 cells
// wide and its column -1 shares the stride with the previous row's unused 10th cell; pairs of columns start on even
// cells (16-byte transfers), and the stride keeps the lanes of a 64-bit access on different banks
constexpr int kURow = 10;
constexpr int kUCells = kURow * (kBy + 2) + 2;     // 102
constexpr int kFRow = 10;                          // slowness tile: row stride (8 used)
constexpr int kFCells = kFRow * kBy;
constexpr int kMaxZc = 256;
constexpr int kProgShift = 12;
constexpr int kLead = kBy + 4;                     // a slot is stored 4 steps after it was entered: slot m + By is in memory once m + By + 5 steps are done
constexpr int kRec = kBx * kBy + 2 * kBy;          // blocked layout: doubles per brick plane (64 nodes + the two x-face copies)
constexpr int kFacePlanes = 16;                    // planes of the face staging ring (a plane is written over 9 steps, its halves stored 11..16 steps after it was entered)
constexpr int kBcMax = 16;                         // planes of one brick that hold boundary-condition nodes of one field
// Fields walked per task.  A flavour with two fields of one slowness model per task (slots side by side, slowness tile
// loaded once, 4 update chains per lane, 8 warps per SM) was built and measured slower at every field count
// (profiles/kernel_evolution_r2.md); the per-field loops below are what is left of it.
constexpr int kNF = 1;
#ifndef MCEIK_B16_WARPS
#define MCEIK_B16_WARPS 12
#endif
constexpr int kWarps = MCEIK_B16_WARPS;
constexpr int kSlot = kNF * kUCells + kFCells;      // 182 doubles, 16-byte aligned
constexpr int kBcBytes = kNF * kBcMax * 16;         // [kNF][kBcMax] masks (8 B), then [kNF][kBcMax] planes (4 B, padded)
constexpr int kFaceCells = kNF * 2 * kFacePlanes * kBy;  // blocked layout: [field][x side][plane & 15][row] face values on
                                                         // their way to the brick records (see BrickArgs::blocked)
constexpr size_t kWarpSmem = sizeof(double) * (size_t)(kRing * kSlot + kFaceCells) + kBcBytes;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
// predicated forms: no branch around the copy (the halo lanes of a steady step differ only in a predicate)
__device__ __forceinline__ void cp_async16_if(bool p, void *smem_dst, const void *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q cp.async.cg.shared.global [%0], [%1], 16; }" ::"r"(s), "l"(gsrc), "r"((int)p) : "memory");
}
__device__ __forceinline__ void cp_async8_if(bool p, void *smem_dst, const void *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q cp.async.ca.shared.global [%0], [%1], 8; }" ::"r"(s), "l"(gsrc), "r"((int)p) : "memory");
}
// TMA-family experiment (MCEIK_FSM_L2PF = planes ahead; 0 = off, the default): one lane asks the L2 for whole brick
// records ahead of the ring's own copies with a bulk prefetch
__device__ __forceinline__ void bulk_prefetch_l2(const void *g, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int xgroup(int i) { return (i + 4) >> 2; }

// Publisher variant (few active fields: the release fence of a publication, not the arithmetic, is what a sweeping
// warp waits for).  The last warp of the CTA sweeps nothing: it takes (progress word, value) requests from the other
// warps through shared memory, executes one gpu-scope fence for all pending requests and stores the values.  The
// hand-off is a cta-scope release / acquire, so the sweeping warp's stores happen-before the publisher's fence and
// its relaxed store completes a (cumulative) gpu-scope release pattern -- the same guarantee as st.release.gpu by
// the sweeping warp itself, minus the stall.
struct Mail {
    unsigned long long req;  // (sequence << 32) | value, written with st.release.cta
    unsigned int ack;        // last sequence the publisher has stored, written with st.release.cta
    int *ptr;                // progress word of the request; changes only when ack == sequence
};
__device__ __forceinline__ unsigned long long ld_acquire_cta_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.cta.shared.u64 %0, [%1];" : "=l"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.cta.shared.u64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_cta_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu(int *p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace

template <bool kPub>
__global__ void __launch_bounds__(kWarps * 32, 1) sweep_bricks16_kernel(const BrickArgs a) {
    constexpr int kNQ = kNF * kNC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ Mail mail[kWarps];
    __shared__ unsigned int workers_done;
    constexpr int kWorkers = kPub ? kWarps - 1 : kWarps;
    unsigned int my_seq = 0;  // lane 0 of a sweeping warp: sequence of its last request
    if (kPub) {
        if (threadIdx.x < kWarps) { mail[threadIdx.x].req = 0ULL; mail[threadIdx.x].ack = 0u; mail[threadIdx.x].ptr = nullptr; }
        if (threadIdx.x == 0) workers_done = 0u;
        __syncthreads();
        if (warp == kWarps - 1) {  // ---- the publisher
            unsigned int last = 0;  // lane w: last sequence stored for sweeping warp w
            while (true) {
                const unsigned int fin = ld_acquire_cta_u32(&workers_done);  // read BEFORE the requests: a warp posts, then finishes
                unsigned long long r = 0ULL;
                bool pend = false;
                if (lane < kWorkers) {
                    r = ld_acquire_cta_u64(&mail[lane].req);
                    pend = (unsigned int)(r >> 32) != last;
                }
                if (__any_sync(0xffffffffu, pend)) {
                    __threadfence();  // fence.sc.gpu >= acq_rel: everything the acquired requests cover is performed gpu-wide
                    if (pend) {
                        st_relaxed_gpu(mail[lane].ptr, (int)(unsigned int)r);
                        last = (unsigned int)(r >> 32);
                        st_release_cta_u32(&mail[lane].ack, last);
                    }
                } else {
                    if (fin == (unsigned int)kWorkers) break;
                    __nanosleep(64);
                }
            }
            return;
        }
    }
    // progress publication of a sweeping warp (lane 0 only, after __syncwarp)
    auto publish_progress = [&](int *ptr, int val) {
        if (kDbg & 2) return;
        if (kPub) {
            Mail &m = mail[warp];
            if (m.ptr != ptr) {  // a request for another brick may still be on its way: let it go out first
                while (ld_acquire_cta_u32(&m.ack) != my_seq) __nanosleep(32);
                m.ptr = ptr;
            }
            ++my_seq;
            st_release_cta_u64(&m.req, ((unsigned long long)my_seq << 32) | (unsigned int)val);
        } else {
            st_release_gpu(ptr, val);
        }
    };
    double *U = reinterpret_cast<double *>(smem_raw + (size_t)warp * kWarpSmem);  // [kRing][kSlot]
    double *FB = U + kRing * kSlot;                                                            // [kNF][2][kFacePlanes][8]
    unsigned long long *bc_mask = reinterpret_cast<unsigned long long *>(FB + kFaceCells);  // [kNF][kBcMax]
    int *bc_plane = reinterpret_cast<int *>(bc_mask + kNF * kBcMax);                           // [kNF][kBcMax], -1 = free

    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const size_t nxy = (size_t)nx * ny, N = nxy * nz;
    const int nf = a.nfields_active;
    const long long ntasks = 8LL * a.nbricks * nf;
    const int nl = a.nblevels, nf0 = a.nf0, nf1 = nf - a.nf0;
    // compute map: lane -> sweep column li and rows j0, j0 + 1
    const int li = lane >> 2, j0 = kNC * (lane & 3);
    const int ig = xgroup(li);
    const bool xm_prev = (li & 3) == 0, xp_next = (li & 3) == 3;
    // transfer map: lane -> pair (memory columns 2p, 2p + 1) of row jt
    const int tp = lane & 3, jt = lane >> 2;
    const int publish = a.publish;

    while (true) {
        long long t = 0;
        if (lane == 0) t = (long long)atomicAdd(a.queue, 1ULL);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntasks) break;
        const long long t_start = a.stats ? clock64() : 0;

        // ---- ticket -> virtual level -> (group, sweep, brick level, brick, field)
        TicketTask task;
        if (a.batch > 0) {  // fields in batches of a.batch, one batch after the other (vptr is the one-field table)
            const long long per = 8LL * a.nbricks * a.batch;
            const int b = (int)(t / per);
            const int nfb = min(a.batch, nf - b * a.batch);
            const long long r = t - (long long)b * per;
            task = decode_ticket(r / nfb, a.vptr, a.blevel_ptr, nl, 0, 1, 0);
            const long long r2 = r - a.vptr[task.sweep * nl + task.level] * nfb;
            task.bidx = (int)(r2 / nfb);
            task.fidx = b * a.batch + (int)(r2 - (long long)task.bidx * nfb);
        } else {
            task = decode_ticket(t, a.vptr, a.blevel_ptr, nl, a.stagger, nf0, nf1);
        }
        const int s = task.sweep;
        int fld[kNF];
        fld[0] = __ldg(a.active + task.fidx);
        const int f = fld[0];
        const int packed = __ldg(a.brick_order + __ldg(a.blevel_ptr + task.level) + task.bidx);
        const bool revx = (s & 1) != 0, revy = (s & 2) != 0, revz = (s & 4) != 0;  // fsm3d.f90:46-53
        int I = packed & 1023, J = (packed >> 10) & 1023, K = packed >> 20;
        if (revx) I = a.nbx - 1 - I;
        if (revy) J = a.nby - 1 - J;
        if (revz) K = a.nbz - 1 - K;
        const int brick = (K * a.nby + J) * a.nbx + I;
        int *done_f = a.done + (size_t)f * a.nbricks;

        // ---- dependencies (see sweep_bricks_kernel)
        const int *up_ptr = nullptr;
        if (lane < 7) {
            int di = 0, dj = 0, dk = 0;
            if (lane == 1) di = -1; else if (lane == 2) di = 1;
            else if (lane == 3) dj = -1; else if (lane == 4) dj = 1;
            else if (lane == 5) dk = -1; else if (lane == 6) dk = 1;
            const int NI = I + di, NJ = J + dj, NK = K + dk;
            if (NI >= 0 && NI < a.nbx && NJ >= 0 && NJ < a.nby && NK >= 0 && NK < a.nbz) {
                int need = s << kProgShift;
                if (dk != 0 && dk == (revz ? 1 : -1)) need = (s + 1) << kProgShift;
                const int *p = done_f + ((NK * a.nby + NJ) * a.nbx + NI);
                if (!(kDbg & 2))
                    while (ld_acquire_gpu(p) < need) __nanosleep(400);
            }
        }
        if (lane < 2) {
            const int NI = I + (lane == 0 ? (revx ? 1 : -1) : 0), NJ = J + (lane == 1 ? (revy ? 1 : -1) : 0);
            if (NI >= 0 && NI < a.nbx && NJ >= 0 && NJ < a.nby) up_ptr = done_f + ((K * a.nby + NJ) * a.nbx + NI);
        }
        // rows of this brick column inside the grid: bricks that are full in y hand their x-face planes on in halves
        const int x_adj = min(J * kBy + kBy, ny) - J * kBy == kBy ? 2 : -2;
        long long t_upwind = 0;
        // progress of the watched neighbour: `seen` = highest value observed so far (it only grows), `ahead` = an
        // observation issued one chunk earlier whose latency is hidden behind that chunk's arithmetic
        int seen = 0, ahead = 0;
        auto wait_upwind = [&](int steps_needed) {
            const long long t0 = a.stats ? clock64() : 0;
            if (up_ptr && !(kDbg & 2)) {
                // lane 0 watches the upwind x neighbour, lane 1 the upwind y neighbour.  A slot's y-halo row is the
                // neighbour's row By-1 (its slot index is By larger), its x-halo column the neighbour's last
                // column (slot index only xgroup(7) - xgroup(-1) = 2 larger): x needs By - 2 steps less lead.
                // Blocked layout: the x-halo comes from the neighbour's face copies, written in half planes (rows 0..3
                // 12, rows 4..7 16 steps after the plane was entered): x needs 2 steps less lead than y (2 more
                // for brick columns cut by the grid's y face, whose planes are written whole).
                const int need = (s << kProgShift) + steps_needed - (lane == 0 ? (a.blocked ? x_adj : kBy - 2) : 0);
                seen = max(seen, ahead);
                if (seen < need)
                    while ((seen = ld_acquire_gpu(up_ptr)) < need) __nanosleep(200);
                if (seen < ((s + 1) << kProgShift)) ahead = ld_acquire_gpu(up_ptr);  // consumed at the next chunk
            }
            __syncwarp();
            if (a.stats) t_upwind += clock64() - t0;
        };
        const long long t_deps = a.stats ? clock64() : 0;
        wait_upwind(4 + kPrefetch + kLead);
        __syncwarp();

        // brick frame: x in memory order (q = x - x_lo), y and z in sweep order
        const int x_lo = I * kBx, y_lo = J * kBy, z_lo = K * a.zc;
        const int y_hi = min(y_lo + kBy, ny) - 1, z_hi = min(z_lo + a.zc, nz) - 1;
        const int ey = y_hi - y_lo + 1, ez = z_hi - z_lo + 1;
        const int yb = revy ? y_hi : y_lo, zb = revz ? z_hi : z_lo;
        const int sx = revx ? -1 : 1, sy = revy ? -1 : 1, sz = revz ? -1 : 1;
        // blocked layout (BrickArgs::blocked): a field is [brick column][z][80], the slowness [brick column][z][64]
        const bool blocked = a.blocked != 0;
        const size_t ncol = (size_t)a.nbx * a.nby;
        const size_t fstride = blocked ? ncol * nz * kRec : N, sstride = blocked ? ncol * nz * (kBx * kBy) : N;
        double *ufs[kNF];
#pragma unroll
        for (int fi = 0; fi < kNF; ++fi) ufs[fi] = a.u + (size_t)fld[fi] * fstride;
        const double *sl = a.slow + (size_t)__ldg(a.field_model + f) * sstride;

        // ---- boundary-condition nodes inside the brick: per field a short list of (plane, 64-bit mask) with
        // bit = j * 8 + i in sweep coordinates (at most kBcMax planes per field and brick; the host checks)
        bool hasbc = false;
#pragma unroll
        for (int fi = 0; fi < kNF; ++fi) {
            const int b0 = __ldg(a.bc_ptr + fld[fi]), b1 = __ldg(a.bc_ptr + fld[fi] + 1);
            bool mine = false;
            for (int n = b0 + lane; n < b1; n += 32) {
                const int node = __ldg(a.bc_node + n);
                const int gx = node % nx, gy = (node / nx) % ny, gz = node / (nx * ny);
                mine |= gx >= x_lo && gx < x_lo + kBx && gy >= y_lo && gy <= y_hi && gz >= z_lo && gz <= z_hi;
            }
            hasbc = __any_sync(0xffffffffu, mine) || hasbc;
        }
        if (hasbc) {
            for (int e = lane; e < kNF * kBcMax; e += 32) { bc_mask[e] = 0ULL; bc_plane[e] = -1; }
            __syncwarp();
#pragma unroll
            for (int fi = 0; fi < kNF; ++fi) {
                const int b0 = __ldg(a.bc_ptr + fld[fi]), b1 = __ldg(a.bc_ptr + fld[fi] + 1);
                for (int n = b0 + lane; n < b1; n += 32) {
                    const int node = __ldg(a.bc_node + n);
                    const int gx = node % nx, gy = (node / nx) % ny, gz = node / (nx * ny);
                    if (gx >= x_lo && gx < x_lo + kBx && gy >= y_lo && gy <= y_hi && gz >= z_lo && gz <= z_hi) {
                        const int i = revx ? x_lo + kBx - 1 - gx : gx - x_lo, j = (gy - yb) * sy, k = (gz - zb) * sz;
                        for (int e = 0; e < kBcMax; ++e) {  // claim or find the entry of plane k
                            const int old = atomicCAS(bc_plane + fi * kBcMax + e, -1, k);
                            if (old == -1 || old == k) { atomicOr(bc_mask + fi * kBcMax + e, 1ULL << (j * kBx + i)); break; }
                        }
                    }
                }
            }
            __syncwarp();
        }
        auto is_bc = [&](int fi, int k, int bit) {  // rare path (bricks holding a source)
            unsigned long long m = 0ULL;
            for (int e = 0; e < kBcMax; ++e)
                if (bc_plane[fi * kBcMax + e] == k) m = bc_mask[fi * kBcMax + e];
            return ((m >> bit) & 1ULL) != 0ULL;
        };

        // ---- transfer descriptors of this lane (all per task): interior pair, slowness pair, halo pair
        // plane strides of a field and of the slowness (blocked: one record each)
        const long long zstride = blocked ? (long long)sz * kRec : (long long)sz * (long long)nxy;
        const long long szstride = blocked ? (long long)sz * (kBx * kBy) : zstride;
        const int klo = (zb - sz < 0 || zb - sz > nz - 1) ? 0 : -1;
        const int khi = (zb + sz * ez < 0 || zb + sz * ez > nz - 1) ? ez - 1 : ez;
        const size_t col = (size_t)J * a.nbx + I;  // brick column (blocked records are indexed by column and plane)
        // interior pair (2tp, 2tp+1) of row jt: sweep index of its first node decides the x-group
        const int gi = xgroup(revx ? kBx - 1 - 2 * tp : 2 * tp);
        const int kofs_t = gi + jt + 2;
        // rows beyond ey of a partial brick are transferred too: row ey is the clamped boundary row or, in a sweep
        // against y, the nearest row of the next brick (it serves as the last row's halo)
        const int yt = min(max(yb + sy * jt, 0), ny - 1);
        const size_t colt = (size_t)(yt / kBy) * a.nbx + I;
        const size_t rowt = blocked ? (colt * nz + zb) * kRec + (size_t)(yt % kBy) * kBx + 2 * tp
                                    : (size_t)zb * nxy + (size_t)yt * nx + x_lo + 2 * tp;
        const double *pf_t = sl + (blocked ? (colt * nz + zb) * (kBx * kBy) + (size_t)(yt % kBy) * kBx + 2 * tp : rowt);
        // ring cells: field fi's rows start at fi * kUCells, the slowness tile follows the last field
        const int cu_t = (jt + 1) * kURow + 2 * tp + 2, cf_t = kNF * kUCells + jt * kFRow + 2 * tp;
        const bool act_t = jt < ey;
        // halo transfer of lanes 0..23: 0-7 left x halo of row h; 8-15 right x halo of row h-8; 16-19 row j = -1, pair
        // h-16; 20-23 row j = By, pair h-20.  hmode: 0 none, 1 pair copy (y rows), 2 one value (x columns): the
        // neighbour's column -- blocked: the entry of its face copy -- or, for a brick on the grid's x face, the
        // boundary column itself (the halo repeats the boundary node, fsm3d.f90:495-499)
        int hmode = 0, kofs_h = 0, cu_h = 0;
        size_t offh = 0;  // offset of the halo transfer's source inside a field
        if (lane < 16) {
            const bool left = lane < 8;
            const int hj = lane & 7;
            const int ih = left ? (revx ? kBx : -1) : (revx ? -1 : kBx);  // sweep index of the needed halo column
            kofs_h = xgroup(ih) + hj + 2;
            const int yh = min(max(yb + sy * hj, 0), ny - 1);
            const size_t row = (size_t)zb * nxy + (size_t)yh * nx;
            const bool in_f = left ? I > 0 : I < a.nbx - 1;  // else the halo is the clamped boundary column itself
            if (hj < ey) {
                hmode = 2; cu_h = (hj + 1) * kURow + (left ? 1 : 10);
                if (blocked) {
                    const int side = left ? (in_f ? 1 : 0) : (in_f ? 0 : 1);  // 0 = copy of column 0, 1 = of column 7
                    const size_t cn = left ? (in_f ? col - 1 : col) : (in_f ? col + 1 : col);
                    offh = (cn * nz + zb) * kRec + kBx * kBy + side * kBy + (yh - y_lo);
                } else {
                    offh = row + (left ? (in_f ? x_lo - 1 : x_lo) : (in_f ? x_lo + kBx : x_lo + kBx - 1));
                }
            }
        } else if (lane < 24) {
            const bool low = lane < 20;
            const int hp = lane & 3;
            const int gh = xgroup(revx ? kBx - 1 - 2 * hp : 2 * hp);
            const int hj = low ? -1 : kBy;
            kofs_h = gh + hj + 2;
            hmode = 1;
            cu_h = (hj + 1) * kURow + 2 * hp + 2;
            const int yh = min(max(yb + sy * hj, 0), ny - 1);  // the neighbour brick's nearest row, or the clamped boundary row
            offh = blocked ? (((size_t)(yh / kBy) * a.nbx + I) * nz + zb) * kRec + (size_t)(yh % kBy) * kBx + 2 * hp
                           : (size_t)zb * nxy + (size_t)yh * nx + x_lo + 2 * hp;
        }
        // blocked: values of memory columns 0 and 7 go to the face staging FB[field][side][plane & 15][row] when their
        // pair is written back, and half a plane of a side (4 rows, 32 contiguous bytes of the record) is stored the
        // step after its rows are in (so the step's own __syncwarp orders it): rows 4..7 15 steps after the plane was
        // entered for the side of sweep column 0, 16 for sweep column 7; rows 0..3 four steps earlier
        const bool face_lane = blocked && (tp == 0 || tp == 3) && act_t;
        const int fb_t = (tp == 0 ? 0 : 1) * (kFacePlanes * kBy) + (yt - y_lo);  // + (plane & 15) * kBy + fi * 2 * kFacePlanes * kBy
        const int fl_side = (lane >> 2) & 1, fl_pair = lane & 3;                 // flush lanes 0..7: side, pair of rows
        // rows 0..3 (in sweep order) of a face plane are complete 4 steps before rows 4..7: each half (32 bytes, one
        // sector) is stored as soon as it is, so the downwind x neighbour can follow 4 steps closer
        // (bricks that are full in y; x neighbours share J, hence ey, so both sides of the hand-off agree)
        const int fl_half = (revy ? 3 - fl_pair : fl_pair) >> 1;                 // 0 = the rows the sweep enters first
        const int fl_lag = (((fl_side == 0) == !revx) ? 15 : 16) - (fl_half == 0 && ey == kBy ? 4 : 0);  // memory column 0 is sweep column 0 unless revx
        const size_t fl_off = (col * nz + zb) * kRec + kBx * kBy + fl_side * kBy + 2 * fl_pair;
        if (blocked && ey < kBy) {  // rows outside the grid are never written: keep their face entries at u_nan
            for (int e = lane; e < kFaceCells; e += 32) FB[e] = DBL_MAX;
            __syncwarp();
        }
        // per-field bases of the three transfers of this lane: interior pair (loaded and written back), halo source
        double *pu_t[kNF];
        const double *pu_h[kNF];
        double *pu_fl[kNF];  // blocked, lanes 0..7: this lane's 16 bytes of the record's face copies
#pragma unroll
        for (int fi = 0; fi < kNF; ++fi) {
            pu_t[fi] = ufs[fi] + rowt;
            pu_h[fi] = ufs[fi] + offh;
            pu_fl[fi] = ufs[fi] + fl_off;
        }

        int ld_slot = 0, ld_m = 0;
        // steady steps form every global address as (per-task byte base) + zo, one running plane offset in bytes
        const long long zsb = zstride * (long long)sizeof(double);
        long long zo = -(long long)kofs_t * zsb;                            // (ld_m - kofs_t) planes
        // the slowness has its own plane stride when blocked (running offset zos)
        const long long szsb = szstride * (long long)sizeof(double);
        long long zos = -(long long)kofs_t * szsb;
        const char *bf_t = reinterpret_cast<const char *>(pf_t);
        const char *bu_t[kNF], *bu_h[kNF];
        char *bu_st[kNF], *bu_fl[kNF];
#pragma unroll
        for (int fi = 0; fi < kNF; ++fi) {
            bu_t[fi] = reinterpret_cast<const char *>(pu_t[fi]);
            bu_h[fi] = reinterpret_cast<const char *>(pu_h[fi]) + (long long)(kofs_t - kofs_h) * zsb;  // halo transfer of the same issue
            bu_st[fi] = reinterpret_cast<char *>(pu_t[fi]) - (9 + kPrefetch) * zsb;  // pair written back in the same step (slot l - 4)
            // face plane stored in step l: plane l - fl_lag; zo stands at (l + 7 - kofs_t) planes when the stores are issued
            bu_fl[fi] = reinterpret_cast<char *>(pu_fl[fi]) + (long long)(kofs_t - 7 - fl_lag) * zsb;
        }
#ifdef MCEIK_B16_L2PF  // measured 4 % slower (profiles/kernel_evolution_r2.md): compiled in only for the experiment
        const int pf_planes = blocked ? a.l2_prefetch : 0;
#else
        constexpr int pf_planes = 0;
#endif
        const double *rec0 = ufs[0] + (col * nz + zb) * kRec, *srec0 = sl + (col * nz + zb) * (kBx * kBy);
        auto issue_slot = [&](auto steady_tag) {
            constexpr bool kSteady = decltype(steady_tag)::value;
            double *sp = U + ld_slot;
            const int k = ld_m - kofs_t;
            if (kDbg & 4) {
            } else if (kSteady) {
#pragma unroll
                for (int fi = 0; fi < kNF; ++fi) cp_async16(sp + fi * kUCells + cu_t, bu_t[fi] + zo);
                cp_async16(sp + cf_t, bf_t + zos);
                if (pf_planes > 0 && lane == 0) {
                    const int kp = ld_m + pf_planes;  // plane (in sweep order) the prefetch asks for
                    if (kp < ez) {
                        bulk_prefetch_l2(rec0 + (long long)kp * zstride, kRec * 8);
                        bulk_prefetch_l2(srec0 + (long long)kp * szstride, kBx * kBy * 8);
                    }
                }
            } else if (k >= -1 && k <= ez) {  // rows beyond ey are loaded too: row ey is the clamped / downwind halo
                const int kc = min(max(k, klo), khi);
#pragma unroll
                for (int fi = 0; fi < kNF; ++fi) cp_async16(sp + fi * kUCells + cu_t, pu_t[fi] + (long long)kc * zstride);
                if (k >= 0 && k < ez) cp_async16(sp + cf_t, pf_t + (long long)kc * szstride);
            }
            const int kh = ld_m - kofs_h;
#pragma unroll
            for (int fi = 0; fi < kNF; ++fi) {
                double *sh = sp + fi * kUCells + cu_h;
                if (kDbg & 4) {
                } else if (kSteady) {  // full brick: every halo lane copies a pair (the x lanes one value)
                    cp_async8_if(hmode == 2, sh, bu_h[fi] + zo);
                    cp_async16_if(hmode == 1, sh, bu_h[fi] + zo);
                } else if (hmode == 1) {
                    if (kh >= 0 && kh < ez) cp_async16(sh, pu_h[fi] + (long long)kh * zstride);
                } else if (hmode == 2) {
                    if (kh >= 0 && kh < ez) cp_async8(sh, pu_h[fi] + (long long)kh * zstride);
                }
            }
            cp_async_commit();
            ++ld_m;
            zo += zsb;
            zos += szsb;
            ld_slot = (ld_slot + kSlot == kRing * kSlot) ? 0 : ld_slot + kSlot;
        };
        for (int m = 0; m < 4 + kPrefetch; ++m) issue_slot(std::false_type());

        // ---- compute descriptors
        const int qx = revx ? kBx - 1 - li : li;            // memory column of this lane's sweep column
        const int cu0 = (j0 + 1) * kURow + qx + 2;          // ring cell of compute column c: cu0 + c * kURow (+ fi * kUCells)
        const int cf0 = kNF * kUCells + j0 * kFRow + qx;
        bool act[kNC];
#pragma unroll
        for (int c = 0; c < kNC; ++c) act[c] = j0 + c < ey;
        const int mc = ig - li + 2;
        int oc = ((mc % kRing) + kRing) % kRing * kSlot;
        int om = (oc == 0) ? (kRing - 1) * kSlot : oc - kSlot;
        int op = (oc + kSlot == kRing * kSlot) ? 0 : oc + kSlot;
        int st_slot = ((-4 % kRing) + kRing) % kRing * kSlot;

        cp_async_wait<kPrefetch - 1>();
        __syncwarp();
        // update chain q = fi * kNC + c: node (li, j0 + c) of field fi
        double self[kNQ], zm[kNQ];
#pragma unroll
        for (int fi = 0; fi < kNF; ++fi) {
#pragma unroll
            for (int c = 0; c < kNC; ++c) {
                const int k = -li - j0 - c, q = fi * kNC + c, cell = fi * kUCells + cu0 + c * kURow;
                self[q] = 0.0; zm[q] = 0.0;
                if (k == 0) {
                    self[q] = U[(ig + j0 + c + 2) * kSlot + cell];
                    zm[q] = U[(ig + j0 + c + 1) * kSlot + cell];
                } else if (k == -1) {
                    self[q] = U[(ig + j0 + c + 1) * kSlot + cell];
                }
            }
        }

        const int nsteps = ez + kBy + 7 + (blocked ? 1 : 0);  // blocked: one more step stores the last face plane
#ifdef MCEIK_B16_PROFILE  // step-phase cycle counters (lane 0): issue | copy wait + syncwarp | reads + solve + writes | write-back
        long long ph[4] = {0, 0, 0, 0};
#define MCEIK_PH(i) do { const long long t_ = clock64(); ph[i] += t_ - tph; tph = t_; } while (0)
        long long tph = clock64();
#else
#define MCEIK_PH(i) do { } while (0)
#endif
        // One step = one brick hyperplane.  A single __syncwarp per step: after it the copies of slots <= l + 4 have landed
        // for every lane and the results of step l - 1 are visible.  Everything else of the step is one region the compiler
        // can interleave with the update chain: the neighbour reads, the write-back of slot l - 4 (final since step l - 1),
        // the copies of slot l + 6 (they overwrite slot l - 5, last read in step l - 1) and the face plane completed in
        // step l - 1.
        auto step_core = [&](int l, auto steady_tag) {
            constexpr bool kSteady = decltype(steady_tag)::value;
#ifdef MCEIK_B16_PROFILE
            tph = clock64();
#endif
            cp_async_wait<kPrefetch - 1>();
            __syncwarp();
            MCEIK_PH(0);

            const int k0 = l - li - j0;
            bool go[kNQ];
#pragma unroll
            for (int fi = 0; fi < kNF; ++fi) {
#pragma unroll
                for (int c = 0; c < kNC; ++c) go[fi * kNC + c] = kSteady || (act[c] && (unsigned)(k0 - c) < (unsigned)ez);
            }
            if (!kSteady && hasbc) {
#pragma unroll
                for (int fi = 0; fi < kNF; ++fi) {
#pragma unroll
                    for (int c = 0; c < kNC; ++c)
                        if (go[fi * kNC + c] && is_bc(fi, k0 - c, (j0 + c) * kBx + li)) go[fi * kNC + c] = false;
                }
            }
            double ux[kNQ], uy[kNQ], uz[kNQ], fh[kNQ], zp[kNQ], nv[kNQ];
#pragma unroll
            for (int fi = 0; fi < kNF; ++fi) {
                const double *pm = U + om + fi * kUCells + cu0, *pc = U + oc + fi * kUCells + cu0, *pp = U + op + fi * kUCells + cu0;
                const double *pxm = (xm_prev ? pm : pc) - sx, *pxp = (xp_next ? pp : pc) + sx;
#pragma unroll
                for (int c = 0; c < kNC; ++c) zp[fi * kNC + c] = pp[c * kURow];
#pragma unroll
                for (int c = 0; c < kNC; ++c) {
                    const int q = fi * kNC + c;
                    ux[q] = dmin2(pxm[c * kURow], pxp[c * kURow]);
                    // the lane's own column supplies two of the y neighbours from registers: row j0+c-1 one plane back is
                    // what the lane wrote (or kept) for its node c-1 in the previous step, row j0+c+1 is zp of node c+1
                    const double ym = c > 0 ? zm[q - 1] : pm[(c - 1) * kURow];
                    const double yp = c + 1 < kNC ? zp[q + 1] : pp[(c + 1) * kURow];
                    uy[q] = dmin2(ym, yp);
                    uz[q] = dmin2(zm[q], zp[q]);
                    fh[q] = U[oc + cf0 + c * kFRow];  // slow(ijk)*h (fsm3d.f90:470), multiplied once per solve (scale_slowness)
                }
            }

            // slot l - 4 is final: its pair of this lane goes back (one 16-byte store per field); lanes 0..7 also store the
            // face plane whose last row arrived in the previous step
            const int ks = l - 4 - kofs_t, kf = l - fl_lag;
            const bool st_pair = !(kDbg & 8) && (kSteady || (act_t && (unsigned)ks < (unsigned)ez));
            const bool st_face = blocked && lane < 8 && !(kDbg & 8) && (kSteady || (unsigned)kf < (unsigned)ez);
            double2 wb[kNF], wf[kNF];
#pragma unroll
            for (int fi = 0; fi < kNF; ++fi) {
                wb[fi] = wf[fi] = make_double2(0.0, 0.0);
                if (st_pair) wb[fi] = *reinterpret_cast<const double2 *>(U + st_slot + fi * kUCells + cu_t);
                if (st_face) wf[fi] = *reinterpret_cast<const double2 *>(FB + fi * 2 * kFacePlanes * kBy + fl_side * (kFacePlanes * kBy) +
                                                                          (kf & (kFacePlanes - 1)) * kBy + 2 * fl_pair);
            }
            issue_slot(steady_tag);  // zo now stands at (l + 7 - kofs_t) planes
#pragma unroll
            for (int fi = 0; fi < kNF; ++fi) {
                if (st_pair) {
                    if (kSteady) __stcg(reinterpret_cast<double2 *>(bu_st[fi] + zo), wb[fi]);
                    else __stcg(reinterpret_cast<double2 *>(pu_t[fi] + (long long)ks * zstride), wb[fi]);
                    if (face_lane) FB[fi * 2 * kFacePlanes * kBy + fb_t + (ks & (kFacePlanes - 1)) * kBy] = tp == 0 ? wb[fi].x : wb[fi].y;
                }
                if (st_face) {
                    if (kSteady) __stcg(reinterpret_cast<double2 *>(bu_fl[fi] + zo), wf[fi]);
                    else __stcg(reinterpret_cast<double2 *>(pu_fl[fi] + (long long)kf * zstride), wf[fi]);
                }
            }
            MCEIK_PH(1);

            if (kDbg & 1) {
#pragma unroll
                for (int q = 0; q < kNQ; ++q) nv[q] = __dadd_rn(__dadd_rn(ux[q], uy[q]), __dadd_rn(uz[q], fh[q]));
            } else {
                local_solve_xn<kNQ>(ux, uy, uz, fh, go, nv);
            }
#pragma unroll
            for (int fi = 0; fi < kNF; ++fi) {
#pragma unroll
                for (int c = 0; c < kNC; ++c) {
                    const int q = fi * kNC + c;
                    const bool upd = go[q] && nv[q] < self[q];  // u = MIN(u, ubar) (fsm3d.f90:477)
                    if (upd) U[oc + fi * kUCells + cu0 + c * kURow] = nv[q];
                    zm[q] = upd ? nv[q] : self[q];
                    self[q] = zp[q];
                }
            }
            MCEIK_PH(2);
            om = oc; oc = op;
            op = (op + kSlot == kRing * kSlot) ? 0 : op + kSlot;
            st_slot = (st_slot + kSlot == kRing * kSlot) ? 0 : st_slot + kSlot;
            MCEIK_PH(3);
        };
        // progress of the brick is published (and the upwind bricks' progress awaited) every `publish` steps
        auto wait_chunk = [&](int l) { wait_upwind(l + publish + 4 + kPrefetch + kLead); };
        auto publish_chunk = [&](int l1) {
            if (l1 < nsteps) {
                __syncwarp();
                if (lane == 0) publish_progress(done_f + brick, (s << kProgShift) + l1);
            }
        };
        auto step = [&](int l) {
            if ((l & (publish - 1)) == 0) wait_chunk(l);
            step_core(l, std::false_type());
            if (((l + 1) & (publish - 1)) == 0) publish_chunk(l + 1);
        };
        // steady window: all transfers, updates and stores of the step touch in-brick nodes of a full
        // brick that is not on the grid's x faces and holds no boundary-condition node; it runs in whole
        // publication chunks with no per-step bookkeeping
        const bool full = ey == kBy && !hasbc;
        // (the pair stored in a steady step, plane l - 15 at least, and the face plane, l - 16, must exist)
        int s_lo = (kBy + 8 + publish - 1) & ~(publish - 1), s_hi = (ez - 3 - kPrefetch) & ~(publish - 1);
        if (!full || s_hi <= s_lo) s_lo = s_hi = nsteps;
        int l = 0;
        for (; l < s_lo; ++l) step(l);
        for (; l < s_hi; l += publish) {
            wait_chunk(l);
            for (int q = 0; q < publish; ++q) step_core(l + q, std::true_type());
            publish_chunk(l + publish);
        }
        for (; l < nsteps; ++l) step(l);

        cp_async_wait<0>();
        __syncwarp();
        if (lane == 0) publish_progress(done_f + brick, (s + 1) << kProgShift);
        __syncwarp();
#ifdef MCEIK_B16_PROFILE
        if (a.stats && lane == 0)
            for (int i = 0; i < 4; ++i) atomicAdd(a.stats + 4 + i, (unsigned long long)ph[i]);
#endif
        if (a.stats && lane == 0) {
            const long long t_end = clock64();
            atomicAdd(a.stats + 0, (unsigned long long)(t_deps - t_start));
            atomicAdd(a.stats + 1, (unsigned long long)t_upwind);
            atomicAdd(a.stats + 2, (unsigned long long)(t_end - t_deps));
            atomicAdd(a.stats + 3, 1ULL);
        }
    }
    if (kPub && lane == 0) {  // the last request is posted before the count goes up; the publisher reads them in the opposite order
        __threadfence_block();
        atomicAdd(&workers_done, 1u);
    }
}

namespace {
template <bool kPub>
void launch_flavour(const BrickArgs &a, int nsm, cudaStream_t st) {
    constexpr int kWorkers = kPub ? kWarps - 1 : kWarps;
    const size_t smem = kWarpSmem * kWarps;
    const long long ntasks = 8LL * a.nbricks * a.nfields_active;
    MCEIK_CUDA(cudaFuncSetAttribute(sweep_bricks16_kernel<kPub>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<long long>((ntasks + kWorkers - 1) / kWorkers, nsm);
    sweep_bricks16_kernel<kPub><<<grid, kWarps * 32, smem, st>>>(a);
}
}  // namespace

void launch_iteration_bricks16(const BrickArgs &a, cudaStream_t st) {
    if (a.nfields_active == 0) return;
    if (a.zc < 1 || a.zc > kMaxZc || a.by != kBy || a.nx % kBx != 0) throw CudaError("bricks16: unsupported geometry");
    if (!a.slow_is_fh) throw CudaError("bricks16: expects the slowness premultiplied by h");
    int dev = 0, nsm = 0;
    MCEIK_CUDA(cudaGetDevice(&dev));
    MCEIK_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    if (a.publisher) launch_flavour<true>(a, nsm, st);
    else launch_flavour<false>(a, nsm, st);
    MCEIK_LAUNCH_CHECK();
}

int bricks16_max_bc_planes() { return kBcMax; }

}  // namespace fsm
}  // namespace mceik
