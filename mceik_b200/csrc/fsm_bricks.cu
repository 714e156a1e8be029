// fsm_bricks.cu -- streaming "brick" sweep kernel of the batched fast-sweeping solver (sm_100a).
//
// Same mathematics and the same dependency argument as the tile kernel in fsm.cu (any order that
// updates upwind neighbours first reproduces the reference's hyperplane order bit for bit,
// fsm3d.f90:62-85, 419-456), different mapping to the machine:
//
//   * A task is one WARP walking one brick (8 x 8 nodes in cross-section -- 8 x 16 as an option --
//     zc nodes long) of one field for one sweep.  Lane (i, q) owns the columns (i, 2q), (i, 2q+1) of
//     the cross-section and marches along the brick axis: at step l it updates the nodes
//     i + j + k = l, so a full hyperplane of the brick (64 nodes, 2 independent updates per lane) is
//     in flight every step and there is no block-wide barrier anywhere -- only __syncwarp().
//   * The brick streams through a per-warp shared-memory ring of 11 skewed planes (10x10 travel-
//     time cells + 8x8 slowness cells each) filled by cp.async two slots ahead in 32-byte sectors,
//     updated in place, and written back as soon as a slot is final.  Nothing but the ring is ever
//     staged, so loads, the 64 Godunov updates per step and stores overlap continuously, and
//     12 warps (one brick each) are resident per SM.
//   * Bricks form the same DAG as tiles; a persistent grid of independent warps pulls tickets in a
//     topological order and spins on per-(field, brick) completion counters.
#include <algorithm>
#include <type_traits>
#include <vector>
#include "fsm.cuh"
#include "fsm_solve.cuh"

namespace mceik {
namespace fsm {

namespace {

constexpr int kBx = 8;                   // brick cross-section: kBx x (4 * NC) nodes, NC columns per lane
constexpr int kPrefetch = 2;             // ring slots loaded ahead of their first reader
constexpr int kRing = 9 + kPrefetch;     // ring depth in slots (see "ring" below)
constexpr int kURow = kBx + 2;           // cell row stride inside a slot (doubles), halo included
constexpr int kMaxZc = 256;              // longest brick (mask storage)
// Progress word of a (field, brick): (sweeps completed << kProgShift) while idle, and
// (sweep << kProgShift) + steps completed while the brick is being swept.
constexpr int kProgShift = 12;
// A brick may run behind its upwind x / y neighbours: the halo node it loads at step l is in ring
// slot M = l + 4 + prefetch, and the neighbour wrote it back by its step M + By + 3 (y face; M + 5
// for the x face): the neighbour must have completed M + By + 4 steps (BrickCfg::kLead = By + 3).

template <int NC> struct BrickCfg {
    static constexpr int kBy = 4 * NC;
    static constexpr int kUCells = kURow * (kBy + 2);
    static constexpr int kSlot = kUCells + kBx * kBy;   // one slot: travel-time cells (with halo) + slowness cells
    static constexpr int kHalo = 2 * kBy + 2 * kBx;     // halo cells per slot
    static constexpr int kMaskWords = (kBx * kBy + 63) / 64;
    static constexpr int kWarps = NC == 2 ? 12 : 7;  // NC = 2: 12 x 16.5 KB = 198 KB; the rest of the 228 KB stays L1 for cp.async.ca
    static constexpr int kLead = kBy + 3;
    static constexpr size_t kWarpSmem = sizeof(double) * (size_t)(kRing * kSlot) + sizeof(unsigned long long) * kMaxZc * kMaskWords;
};

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
// experiment only (MCEIK_FSM_DEBUG & 32): plain L2-only load of the same address, value folded into acc
__device__ __forceinline__ void dbg_ldg(double &acc, const void *gsrc) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(gsrc) : "memory");
    acc += v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// x-group of a cell column: i = -1 -> 0, 0..3 -> 1, 4..7 -> 2, 8 -> 3 (one 32-byte sector each)
__device__ __forceinline__ int xgroup(int i) { return (i + 4) >> 2; }

}  // namespace

// The ring.  Cell (i, j, k) of the brick (halo: i in [-1, 8], j in [-1, By], k in [-1, ez]) lives in
// ring slot
//     m = xgroup(i) + (j + 1) + (k + 1)        (mod kRing)
// at cell offset (j+1)*10 + (i+1): a slot is a *skewed* plane made of 4-node x-sectors, one per
// (x-group, j), so that global traffic stays sector-granular (32 B) while a slot only lives from the
// step its first node is read to the step its last node is final: the node (i, j, k) is updated at
// step l = i + j + k, i.e. in slot l + c with c = xgroup(i) - i + 2 in [-3, 3] (independent of j, so
// the cross-section can be widened in y for free), and reads slots m-1, m, m+1 only.  Slot m is
// first read at step m-4, last written at step m+3 and last read at step m+4, so kRing = 9 +
// prefetch slots suffice (By + 9 + prefetch planes would be needed without the skew).
//
// A lane owns NC columns (li, NC*q .. NC*q + NC-1), q = lane & 3, li = lane >> 2: NC independent
// Godunov updates per step whose instruction streams the compiler interleaves.  The z-neighbours of
// a column never leave the lane: the value read ahead as zp becomes `self` one step later and the
// lane's own result becomes zm.
template <int NC>
__global__ void __launch_bounds__(BrickCfg<NC>::kWarps * 32, 1) sweep_bricks_kernel(const BrickArgs a) {
    using Cfg = BrickCfg<NC>;
    constexpr int kBy = Cfg::kBy, kUCells = Cfg::kUCells, kSlot = Cfg::kSlot, kMaskWords = Cfg::kMaskWords;
    constexpr int kHaloPerLane = (Cfg::kHalo + 31) / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *U = reinterpret_cast<double *>(smem_raw + (size_t)warp * Cfg::kWarpSmem);  // [kRing][kSlot]
    unsigned long long *bcm = reinterpret_cast<unsigned long long *>(U + kRing * kSlot);  // [kMaxZc][kMaskWords]

    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const size_t nxy = (size_t)nx * ny, N = nxy * nz;
    const int nf = a.nfields_active;
    const long long per_sweep = (long long)a.nbricks * nf;
    const long long ntasks = 8 * per_sweep;
    const int li = lane >> 2, j0 = NC * (lane & 3);
    const int ig = xgroup(li);
    // halo cells of this lane: cell h = lane + 32*t: [0, By): (i=-1, j=h); [By, 2By): (i=8, j=h-By);
    // then (i = h', j = -1) for 8 cells and (i = h'', j = By) for 8 cells
    int cuh[kHaloPerLane], kofsh[kHaloPerLane], hi_i[kHaloPerLane], hi_j[kHaloPerLane];
#pragma unroll
    for (int t = 0; t < kHaloPerLane; ++t) {
        const int h = lane + 32 * t;
        int hi, hj;
        if (h < kBy) { hi = -1; hj = h; }
        else if (h < 2 * kBy) { hi = kBx; hj = h - kBy; }
        else if (h < 2 * kBy + kBx) { hi = h - 2 * kBy; hj = -1; }
        else { hi = h - 2 * kBy - kBx; hj = kBy; }
        if (h >= Cfg::kHalo) { hi = 0; hj = 0; }
        hi_i[t] = hi; hi_j[t] = hj;
        cuh[t] = (hj + 1) * kURow + (hi + 1);
        kofsh[t] = (h < Cfg::kHalo) ? xgroup(hi) + hj + 2 : (1 << 20);  // k = m - kofs (never valid when unused)
    }
    const int cu0 = (j0 + 1) * kURow + (li + 1);        // cell offset of column c: cu0 + c * kURow
    const int cf0 = kUCells + j0 * kBx + li;            // slowness cell of column c: cf0 + c * kBx
    const int kofs0 = ig + j0 + 2;                      // k of column c in slot m: m - kofs0 - c
    const bool xm_prev = (li & 3) == 0, xp_next = (li & 3) == 3;  // x-neighbour in the previous / next slot
    const int publish = a.publish;  // steps between progress publications (power of two)

    while (true) {
        long long t = 0;
        if (lane == 0) t = (long long)atomicAdd(a.queue, 1ULL);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntasks) break;
        const long long t_start = a.stats ? clock64() : 0;

        // ---- ticket -> (sweep, brick level, brick, field); order [sweep][brick level][brick][field]
        const int s = (int)(t / per_sweep);
        long long r = t - (long long)s * per_sweep;
        int lo = 0, hi = a.nblevels;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((long long)nf * __ldg(a.blevel_ptr + mid) <= r) lo = mid; else hi = mid;
        }
        r -= (long long)nf * __ldg(a.blevel_ptr + lo);
        const int bidx = (int)(r / nf);
        const int f = __ldg(a.active + (int)(r - (long long)bidx * nf));
        const int packed = __ldg(a.brick_order + __ldg(a.blevel_ptr + lo) + bidx);
        const bool revx = (s & 1) != 0, revy = (s & 2) != 0, revz = (s & 4) != 0;  // fsm3d.f90:46-53
        int I = packed & 1023, J = (packed >> 10) & 1023, K = packed >> 20;
        if (revx) I = a.nbx - 1 - I;
        if (revy) J = a.nby - 1 - J;
        if (revz) K = a.nbz - 1 - K;
        const int brick = (K * a.nby + J) * a.nbx + I;
        int *done_f = a.done + (size_t)f * a.nbricks;

        // ---- dependencies (coarse): this brick and its 6 neighbours finished sweep s-1; the upwind z
        //      neighbour finished sweep s.  The upwind x / y neighbours only need a kLead-step head
        //      start, which is checked every `publish` steps inside the sweep loop (fine-grained
        //      pipelining of the brick wavefront).
        const int *up_ptr = nullptr;  // lanes 0 / 1: progress word of the upwind x / y neighbour
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
                while (ld_acquire_gpu(p) < need) __nanosleep(400);
            }
        }
        if (lane < 2) {
            const int NI = I + (lane == 0 ? (revx ? 1 : -1) : 0), NJ = J + (lane == 1 ? (revy ? 1 : -1) : 0);
            if (NI >= 0 && NI < a.nbx && NJ >= 0 && NJ < a.nby) up_ptr = done_f + ((K * a.nby + NJ) * a.nbx + NI);
        }
        long long t_upwind = 0;
        auto wait_upwind = [&](int steps_needed) {  // upwind x / y neighbours have completed that many steps
            const long long t0 = a.stats ? clock64() : 0;
            if (up_ptr) {
                const int need = (s << kProgShift) + steps_needed;
                while (ld_acquire_gpu(up_ptr) < need) __nanosleep(200);
            }
            __syncwarp();
            if (a.stats) t_upwind += clock64() - t0;
        };
        const long long t_deps = a.stats ? clock64() : 0;
        wait_upwind(4 + kPrefetch + Cfg::kLead);  // the prologue issues slots 0 .. 3 + kPrefetch
        __syncwarp();

        // brick extent in memory coordinates and the sweep-oriented local frame:
        //   local (i, j, k) -> global (revx ? x_hi - i : x_lo + i, ...), clamped to the grid.  A clamped
        //   halo cell repeats the boundary node itself, which is what GET_U?MIN3D substitutes at a
        //   face (fsm3d.f90:495-499, 517-521, 539-543).
        const int x_lo = I * kBx, y_lo = J * kBy, z_lo = K * a.zc;
        const int x_hi = min(x_lo + kBx, nx) - 1, y_hi = min(y_lo + kBy, ny) - 1, z_hi = min(z_lo + a.zc, nz) - 1;
        const int ex = x_hi - x_lo + 1, ey = y_hi - y_lo + 1, ez = z_hi - z_lo + 1;
        const int xb = revx ? x_hi : x_lo, yb = revy ? y_hi : y_lo, zb = revz ? z_hi : z_lo;
        const int sx = revx ? -1 : 1, sy = revy ? -1 : 1, sz = revz ? -1 : 1;
        double *uf = a.u + (size_t)f * N;
        const double *sl = a.slow + (size_t)__ldg(a.field_model + f) * N;

        // ---- boundary-condition nodes of this field inside the brick (never updated)
        bool hasbc = false;
        {
            const int b0 = __ldg(a.bc_ptr + f), b1 = __ldg(a.bc_ptr + f + 1);
            bool mine = false;
            for (int n = b0 + lane; n < b1; n += 32) {
                const int node = __ldg(a.bc_node + n);
                const int gx = node % nx, gy = (node / nx) % ny, gz = node / (nx * ny);
                mine |= gx >= x_lo && gx <= x_hi && gy >= y_lo && gy <= y_hi && gz >= z_lo && gz <= z_hi;
            }
            hasbc = __any_sync(0xffffffffu, mine);
            if (hasbc) {
                for (int k = lane; k < kMaxZc * kMaskWords; k += 32) bcm[k] = 0ULL;
                __syncwarp();
                for (int n = b0 + lane; n < b1; n += 32) {
                    const int node = __ldg(a.bc_node + n);
                    const int gx = node % nx, gy = (node / nx) % ny, gz = node / (nx * ny);
                    if (gx >= x_lo && gx <= x_hi && gy >= y_lo && gy <= y_hi && gz >= z_lo && gz <= z_hi) {
                        const int i = (gx - xb) * sx, j = (gy - yb) * sy, k = (gz - zb) * sz;
                        const int bit = j * kBx + i;
                        atomicOr(bcm + k * kMaskWords + (bit >> 6), 1ULL << (bit & 63));
                    }
                }
                __syncwarp();
            }
        }

        // per-lane base pointers (plane k = 0) of its interior columns and of its halo columns;
        // plane k sits at base + kclamp(k) * zstride, where kclamp folds the two z-halo planes
        // (k = -1, k = ez) back onto the boundary plane when the brick touches the grid boundary.
        const int gxi = min(max(xb + sx * li, 0), nx - 1);
        double *pu[NC];
        const double *pf[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const size_t col = (size_t)zb * nxy + (size_t)min(max(yb + sy * (j0 + c), 0), ny - 1) * nx + gxi;
            pu[c] = uf + col;
            pf[c] = sl + col;
        }
        const double *puh[kHaloPerLane];
#pragma unroll
        for (int t2 = 0; t2 < kHaloPerLane; ++t2)
            puh[t2] = uf + (size_t)zb * nxy + (size_t)min(max(yb + sy * hi_j[t2], 0), ny - 1) * nx +
                      min(max(xb + sx * hi_i[t2], 0), nx - 1);
        const long long zstride = (long long)sz * (long long)nxy;
        const int klo = (zb - sz < 0 || zb - sz > nz - 1) ? 0 : -1;             // plane index used for k = -1
        const int khi = (zb + sz * ez < 0 || zb + sz * ez > nz - 1) ? ez - 1 : ez;  // ... and for k = ez

        int ld_slot = 0;  // element offset of the ring slot the next issue_slot() fills
        int ld_m = 0;
        double dbg_acc = 0.0;
        const bool dbg_ld = (a.debug & 32) != 0;
        auto cpa = [&](double *dst, const double *src) {
            if (dbg_ld) dbg_ldg(dbg_acc, src); else cp_async8(dst, src);
        };
        // `steady` = every cell of the slot is inside the brick (no k-range predicates needed)
        auto issue_slot = [&](auto steady_tag) {
            constexpr bool kSteady = decltype(steady_tag)::value;
            double *sp = U + ld_slot;
            if (a.debug & 4) { cp_async_commit(); ++ld_m; ld_slot = (ld_slot + kSlot == kRing * kSlot) ? 0 : ld_slot + kSlot; return; }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int k = ld_m - kofs0 - c;
                if (kSteady) {
                    const long long z = (long long)k * zstride;
                    cpa(sp + cu0 + c * kURow, pu[c] + z);
                    if (!(a.debug & 16)) cpa(sp + cf0 + c * kBx, pf[c] + z);
                } else if (k >= -1 && k <= ez) {
                    const long long z = (long long)min(max(k, klo), khi) * zstride;
                    cpa(sp + cu0 + c * kURow, pu[c] + z);
                    if (k >= 0 && k < ez) cpa(sp + cf0 + c * kBx, pf[c] + z);
                }
            }
#pragma unroll
            for (int t2 = 0; t2 < kHaloPerLane; ++t2) {
                const int kh = ld_m - kofsh[t2];
                if (!(a.debug & 8) && (kSteady ? (kofsh[t2] < (1 << 20)) : (kh >= 0 && kh < ez))) cpa(sp + cuh[t2], puh[t2] + (long long)kh * zstride);
            }
            cp_async_commit();
            ++ld_m;
            ld_slot = (ld_slot + kSlot == kRing * kSlot) ? 0 : ld_slot + kSlot;
        };
        for (int m = 0; m < 4 + kPrefetch; ++m) issue_slot(std::false_type());

        bool act[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) act[c] = li < ex && j0 + c < ey;
        // slot offsets of the nodes this lane updates at step l: m = l + cofs, cofs = ig - li + 2
        const int mc = ig - li + 2;                 // slot number at l = 0 (may be negative)
        int oc = ((mc % kRing) + kRing) % kRing * kSlot;
        int om = (oc == 0) ? (kRing - 1) * kSlot : oc - kSlot;
        int op = (oc + kSlot == kRing * kSlot) ? 0 : oc + kSlot;
        int st_slot = ((-3 % kRing) + kRing) % kRing * kSlot;  // slot l - 3 at l = 0

        // z-neighbour registers of every column: self = u(k) before its update, zm = u(k-1) after its
        // update.  Columns that start at k = 0 or k = -1 at step 0 take them from the ring (slots <= 4
        // have landed); the others pick them up on their way (zp read at k = -2, -1).
        cp_async_wait<kPrefetch - 1>();
        __syncwarp();
        double self[NC], zm[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int k = -li - j0 - c;  // k of column c at step 0
            self[c] = 0.0; zm[c] = 0.0;
            if (k == 0) {
                self[c] = U[(ig + j0 + c + 2) * kSlot + cu0 + c * kURow];
                zm[c] = U[(ig + j0 + c + 1) * kSlot + cu0 + c * kURow];
            } else if (k == -1) {
                self[c] = U[(ig + j0 + c + 1) * kSlot + cu0 + c * kURow];
            }
        }

        const int nsteps = ez + kBy + 6;
        // One step of the march.  `steady` specialises the body for the steps in which every lane is
        // active for loads, updates and stores of a full brick without boundary-condition nodes: no
        // activity predicates at all (most steps of a long brick).
        auto step = [&](int l, auto steady_tag) {
            constexpr bool kSteady = decltype(steady_tag)::value;
            if ((l & (publish - 1)) == 0) wait_upwind(l + publish + 4 + kPrefetch + Cfg::kLead);
            issue_slot(steady_tag);
            cp_async_wait<kPrefetch>();  // slots <= l + 4 have landed (for this lane)
            __syncwarp();                // ... and for every lane of the warp

            const int k0 = l - li - j0;  // k of column c: k0 - c
            bool go[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) go[c] = kSteady || (act[c] && (unsigned)(k0 - c) < (unsigned)ez);
            if (!kSteady && hasbc) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int bit = (j0 + c) * kBx + li;
                    if (go[c] && ((bcm[(k0 - c) * kMaskWords + (bit >> 6)] >> (bit & 63)) & 1ULL)) go[c] = false;
                }
            }
            // All NC nodes of the lane are evaluated unconditionally and side by side; only the final
            // store is predicated.  An inactive node reads valid ring cells whose values are not used.
            const double *pm = U + om + cu0, *pc = U + oc + cu0, *pp = U + op + cu0;
            const double *pxm = (xm_prev ? pm : pc) - 1, *pxp = (xp_next ? pp : pc) + 1;
            double ux[NC], uy[NC], uz[NC], fh[NC], zp[NC], nv[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                zp[c] = pp[c * kURow];
                ux[c] = dmin2(pxm[c * kURow], pxp[c * kURow]);
                uy[c] = dmin2(pm[(c - 1) * kURow], pp[(c + 1) * kURow]);
                uz[c] = dmin2(zm[c], zp[c]);
                fh[c] = __dmul_rn(U[oc + cf0 + c * kBx], a.h);
            }
            if (a.debug & 1) {
#pragma unroll
                for (int c = 0; c < NC; ++c) nv[c] = __dadd_rn(dmin2(ux[c], dmin2(uy[c], uz[c])), fh[c]);
            } else {
                local_solve_xn<NC>(ux, uy, uz, fh, go, nv);
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const bool upd = go[c] && nv[c] < self[c];  // u = MIN(u, ubar) (fsm3d.f90:477)
                if (upd) U[oc + cu0 + c * kURow] = nv[c];
                zm[c] = upd ? nv[c] : self[c];
                self[c] = zp[c];
            }
            __syncwarp();

            // slot l - 3 is final now: write its nodes back (32-byte sectors, one per lane quad)
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int ks = l - 3 - kofs0 - c;
                if (!(a.debug & 2) && (kSteady || (act[c] && (unsigned)ks < (unsigned)ez)))
                    __stcg(pu[c] + (long long)ks * zstride, U[st_slot + cu0 + c * kURow]);
            }
            om = oc; oc = op;
            op = (op + kSlot == kRing * kSlot) ? 0 : op + kSlot;
            st_slot = (st_slot + kSlot == kRing * kSlot) ? 0 : st_slot + kSlot;
            if (((l + 1) & (publish - 1)) == 0 && l + 1 < nsteps) {  // publish progress
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    st_release_gpu(done_f + brick, (s << kProgShift) + l + 1);
                }
            }
        };
        // steady window: loads (slot l + 4 + prefetch), updates (step l) and stores (slot l - 3) all
        // touch in-brick nodes only for l in [By + 6, ez - 3 - prefetch)
        const bool full = ex == kBx && ey == kBy && !hasbc;
        const int s_lo = full ? kBy + 6 : nsteps, s_hi = full ? ez - 3 - kPrefetch : nsteps;
        for (int l = 0; l < nsteps; ++l) {
            if (l >= s_lo && l < s_hi) step(l, std::true_type());
            else step(l, std::false_type());
        }
        cp_async_wait<0>();
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            st_release_gpu(done_f + brick, (s + 1) << kProgShift);
        }
        __syncwarp();
        if (dbg_ld && dbg_acc == 1.2345e-300) uf[0] = dbg_acc;  // keeps the experiment's loads alive
        if (a.stats && lane == 0) {
            const long long t_end = clock64();
            atomicAdd(a.stats + 0, (unsigned long long)(t_deps - t_start));   // ticket + coarse dependency wait
            atomicAdd(a.stats + 1, (unsigned long long)t_upwind);             // fine-grained upwind waits
            atomicAdd(a.stats + 2, (unsigned long long)(t_end - t_deps));     // everything else (incl. upwind waits)
            atomicAdd(a.stats + 3, 1ULL);
        }
    }
}

template <int NC>
static void launch_bricks_impl(const BrickArgs &a, cudaStream_t st) {
    using Cfg = BrickCfg<NC>;
    const size_t smem = Cfg::kWarpSmem * Cfg::kWarps;
    MCEIK_CUDA(cudaFuncSetAttribute(sweep_bricks_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, nsm = 0;
    MCEIK_CUDA(cudaGetDevice(&dev));
    MCEIK_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const long long ntasks = 8LL * a.nbricks * a.nfields_active;
    const int grid = (int)std::min<long long>((ntasks + Cfg::kWarps - 1) / Cfg::kWarps, nsm);
    sweep_bricks_kernel<NC><<<grid, Cfg::kWarps * 32, smem, st>>>(a);
    MCEIK_LAUNCH_CHECK();
}

void launch_iteration_bricks(const BrickArgs &a, cudaStream_t st) {
    if (a.nfields_active == 0) return;
    if (a.zc < 1 || a.zc > kMaxZc) throw CudaError("brick length out of range");
    if (a.by == 16) launch_bricks_impl<4>(a, st);
    else if (a.by == 8) launch_bricks_impl<2>(a, st);
    else throw CudaError("brick width must be 8 or 16");
}

void BrickPlan::build(int nx_, int ny_, int nz_, int by_, int zc_, cudaStream_t st) {
    if (nx_ == nx && ny_ == ny && nz_ == nz && zc_ == zc && by_ == by && nbricks > 0) return;
    nx = nx_; ny = ny_; nz = nz_; zc = zc_; by = by_;
    nbx = (nx + kBx - 1) / kBx;
    nby = (ny + by - 1) / by;
    nbz = (nz + zc - 1) / zc;
    if (nbx > 1023 || nby > 1023 || nbz > 1023) throw CudaError("grid too large for the brick plan");
    nbricks = nbx * nby * nbz;
    nblevels = nbx + nby + nbz - 2;
    std::vector<int> order, ptr(nblevels + 1, 0);
    order.reserve(nbricks);
    for (int d = 0; d < nblevels; ++d) {
        ptr[d] = (int)order.size();
        for (int K = 0; K < nbz; ++K)
            for (int J = 0; J < nby; ++J) {
                const int I = d - K - J;
                if (I >= 0 && I < nbx) order.push_back(I | (J << 10) | (K << 20));
            }
    }
    ptr[nblevels] = (int)order.size();
    h_blevel_ptr = ptr;
    MCEIK_CUDA(cudaMemcpyAsync(brick_order.ensure(sizeof(int) * nbricks), order.data(), sizeof(int) * nbricks,
                               cudaMemcpyHostToDevice, st));
    MCEIK_CUDA(cudaMemcpyAsync(blevel_ptr.ensure(sizeof(int) * (nblevels + 1)), ptr.data(), sizeof(int) * (nblevels + 1),
                               cudaMemcpyHostToDevice, st));
    MCEIK_CUDA(cudaStreamSynchronize(st));
}

void BrickPlan::release() {
    brick_order.release();
    blevel_ptr.release();
    nbricks = 0;
    nx = ny = nz = zc = by = 0;
}

}  // namespace fsm
}  // namespace mceik
