// fsm_bricks.cu -- streaming "brick" sweep kernel of the batched fast-sweeping solver (sm_100a).
//
// Same mathematics and the same dependency argument as the tile kernel in fsm.cu (any order that
// updates upwind neighbours first reproduces the reference's hyperplane order bit for bit,
// fsm3d.f90:62-85, 419-456), different mapping to the machine:
//
//   * A task is one WARP walking one brick (8 x 8 nodes in cross-section, zc nodes long) of one
//     field for one sweep.  Lane (i, jq) owns the columns (i, jq) and (i, jq+4) of the cross-
//     section and marches along the brick axis: at step l it updates the nodes i + j + k = l, so a
//     full hyperplane of the brick (64 nodes, 2 independent updates per lane) is in flight every
//     step and there is no block-wide barrier anywhere -- only __syncwarp().
//   * Planes of the brick stream through a per-warp shared-memory ring (19 planes of 10x10
//     doubles: 8x8 nodes + halo) filled by cp.async two planes ahead with coalesced 64-byte rows,
//     updated in place, and written back as soon as the last lane is done with a plane.  The
//     slowness plane ring rides along.  Nothing but the ring is ever staged, so loads, the 2 x 64
//     Godunov updates per step and stores overlap continuously.
//   * Bricks form the same DAG as tiles; a persistent grid of independent warps pulls tickets in a
//     topological order and spins on per-(field, brick) completion counters.
#include <algorithm>
#include <vector>
#include "fsm.cuh"
#include "fsm_solve.cuh"

namespace mceik {
namespace fsm {

namespace {

constexpr int kBx = 8, kBy = 8;          // brick cross-section (nodes)
constexpr int kPrefetch = 2;             // planes loaded ahead of the first reader
constexpr int kRing = kPrefetch + 17;    // ring depth: (kBx-1)+(kBy-1) steps of life + store + prefetch
constexpr int kURow = kBx + 2;           // ring row stride (doubles) with halo
constexpr int kUPlane = kURow * (kBy + 2);
constexpr int kFPlane = kBx * kBy;
constexpr int kMaxZc = 64;               // longest brick (mask storage)
constexpr int kWarpsPerCta = 8;
constexpr size_t kWarpSmem = sizeof(double) * (size_t)(kRing * (kUPlane + kFPlane)) + sizeof(unsigned long long) * kMaxZc;

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace

__global__ void __launch_bounds__(kWarpsPerCta * 32, 1) sweep_bricks_kernel(const BrickArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *U = reinterpret_cast<double *>(smem_raw + (size_t)warp * kWarpSmem);  // [kRing][kBy+2][kBx+2]
    double *F = U + kRing * kUPlane;                                              // [kRing][kBy][kBx]
    unsigned long long *bcm = reinterpret_cast<unsigned long long *>(F + kRing * kFPlane);  // [kMaxZc]

    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const size_t nxy = (size_t)nx * ny, N = nxy * nz;
    const int nf = a.nfields_active;
    const long long per_sweep = (long long)a.nbricks * nf;
    const long long ntasks = 8 * per_sweep;
    const int li = lane & 7, jq = lane >> 3;  // lane -> columns (li, jq) and (li, jq + 4)

    while (true) {
        long long t = 0;
        if (lane == 0) t = (long long)atomicAdd(a.queue, 1ULL);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntasks) break;

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

        // ---- dependencies: this brick and its 6 neighbours finished sweep s-1, upwind ones sweep s
        if (lane < 7) {
            int di = 0, dj = 0, dk = 0;
            if (lane == 1) di = -1; else if (lane == 2) di = 1;
            else if (lane == 3) dj = -1; else if (lane == 4) dj = 1;
            else if (lane == 5) dk = -1; else if (lane == 6) dk = 1;
            const int NI = I + di, NJ = J + dj, NK = K + dk;
            if (NI >= 0 && NI < a.nbx && NJ >= 0 && NJ < a.nby && NK >= 0 && NK < a.nbz) {
                int need = s;
                if ((di != 0 && di == (revx ? 1 : -1)) || (dj != 0 && dj == (revy ? 1 : -1)) ||
                    (dk != 0 && dk == (revz ? 1 : -1)))
                    need = s + 1;
                const int *p = done_f + ((NK * a.nby + NJ) * a.nbx + NI);
                while (ld_acquire_gpu(p) < need) __nanosleep(64);
            }
        }
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
                for (int k = lane; k < kMaxZc; k += 32) bcm[k] = 0ULL;
                __syncwarp();
                for (int n = b0 + lane; n < b1; n += 32) {
                    const int node = __ldg(a.bc_node + n);
                    const int gx = node % nx, gy = (node / nx) % ny, gz = node / (nx * ny);
                    if (gx >= x_lo && gx <= x_hi && gy >= y_lo && gy <= y_hi && gz >= z_lo && gz <= z_hi) {
                        const int i = (gx - xb) * sx, j = (gy - yb) * sy, k = (gz - zb) * sz;
                        atomicOr(bcm + k, 1ULL << (j * kBx + i));
                    }
                }
                __syncwarp();
            }
        }

        // per-lane global offsets of its two interior columns and of its one halo cell, at plane k = 0
        const int gxi = min(max(xb + sx * li, 0), nx - 1);
        const int gy0 = min(max(yb + sy * jq, 0), ny - 1), gy1 = min(max(yb + sy * (jq + 4), 0), ny - 1);
        const size_t col0 = (size_t)gy0 * nx + gxi, col1 = (size_t)gy1 * nx + gxi;
        // halo cell of this lane: lanes 0-7: (i=-1, j=lane); 8-15: (i=8, j); 16-23: (i, j=-1); 24-31: (i, j=8)
        int hi_i, hi_j;
        if (lane < 8) { hi_i = -1; hi_j = lane; }
        else if (lane < 16) { hi_i = kBx; hi_j = lane - 8; }
        else if (lane < 24) { hi_i = lane - 16; hi_j = -1; }
        else { hi_i = lane - 24; hi_j = kBy; }
        const size_t colh = (size_t)min(max(yb + sy * hi_j, 0), ny - 1) * nx + min(max(xb + sx * hi_i, 0), nx - 1);
        const int soff0 = (jq + 1) * kURow + (li + 1), soff1 = (jq + 5) * kURow + (li + 1);
        const int soffh = (hi_j + 1) * kURow + (hi_i + 1);
        const int foff0 = jq * kBx + li, foff1 = (jq + 4) * kBx + li;

        // plane kk in [-1, ez] lives in ring slot (kk + 1) % kRing
        auto issue_plane = [&](int kk) {
            if (kk <= ez) {
                const int slot = (kk + 1) % kRing;
                const size_t zoff = (size_t)min(max(zb + sz * kk, 0), nz - 1) * nxy;
                double *up = U + slot * kUPlane;
                cp_async8(up + soff0, uf + zoff + col0);
                cp_async8(up + soff1, uf + zoff + col1);
                if (kk >= 0 && kk < ez) {
                    cp_async8(up + soffh, uf + zoff + colh);
                    double *fp = F + slot * kFPlane;
                    cp_async8(fp + foff0, sl + zoff + col0);
                    cp_async8(fp + foff1, sl + zoff + col1);
                }
            }
            cp_async_commit();
        };

        for (int kk = -1; kk <= kPrefetch; ++kk) issue_plane(kk);

        const bool act0 = li < ex && jq < ey, act1 = li < ex && jq + 4 < ey;
        const int span = (ex - 1) + (ey - 1);
        const int nsteps = ez + span;
        for (int l = 0; l < nsteps; ++l) {
            issue_plane(l + 1 + kPrefetch);
            cp_async_wait<kPrefetch>();  // planes <= l + 1 have landed (for this lane)
            __syncwarp();                // ... and for every lane of the warp

            const int k0 = l - li - jq, k1 = k0 - 4;
            const bool do0 = act0 && k0 >= 0 && k0 < ez && !(hasbc && ((bcm[max(k0, 0)] >> (jq * kBx + li)) & 1ULL));
            const bool do1 = act1 && k1 >= 0 && k1 < ez && !(hasbc && ((bcm[max(k1, 0)] >> ((jq + 4) * kBx + li)) & 1ULL));
            double n0 = 0.0, n1 = 0.0, c0 = 0.0, c1 = 0.0;
            double *p0 = nullptr, *p1 = nullptr;
            if (do0) {
                const int sm = k0 % kRing, sc = (k0 + 1) % kRing, sp = (k0 + 2) % kRing;
                p0 = U + sc * kUPlane + soff0;
                c0 = *p0;
                const double ux = fmin(p0[-1], p0[1]);
                const double uy = fmin(p0[-kURow], p0[kURow]);
                const double uz = fmin(U[sm * kUPlane + soff0], U[sp * kUPlane + soff0]);
                n0 = local_solve_sl(ux, uy, uz, __dmul_rn(F[sc * kFPlane + foff0], a.h));
            }
            if (do1) {
                const int sm = k1 % kRing, sc = (k1 + 1) % kRing, sp = (k1 + 2) % kRing;
                p1 = U + sc * kUPlane + soff1;
                c1 = *p1;
                const double ux = fmin(p1[-1], p1[1]);
                const double uy = fmin(p1[-kURow], p1[kURow]);
                const double uz = fmin(U[sm * kUPlane + soff1], U[sp * kUPlane + soff1]);
                n1 = local_solve_sl(ux, uy, uz, __dmul_rn(F[sc * kFPlane + foff1], a.h));
            }
            if (do0 && n0 < c0) *p0 = n0;  // u = MIN(u, ubar) (fsm3d.f90:477)
            if (do1 && n1 < c1) *p1 = n1;
            __syncwarp();

            // plane kd is final once the last lane (i = ex-1, j = ey-1) has passed it: write it back
            const int kd = l - span;
            if (kd >= 0) {
                const int slot = (kd + 1) % kRing;
                const size_t zoff = (size_t)(zb + sz * kd) * nxy;
                if (act0) __stcg(uf + zoff + col0, U[slot * kUPlane + soff0]);
                if (act1) __stcg(uf + zoff + col1, U[slot * kUPlane + soff1]);
            }
        }
        cp_async_wait<0>();
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            red_release_gpu_add(done_f + brick, 1);
        }
        __syncwarp();
    }
}

size_t bricks_smem_bytes() { return kWarpSmem * kWarpsPerCta; }

void launch_iteration_bricks(const BrickArgs &a, cudaStream_t st) {
    if (a.nfields_active == 0) return;
    if (a.zc < 1 || a.zc > kMaxZc) throw CudaError("brick length out of range");
    const size_t smem = bricks_smem_bytes();
    MCEIK_CUDA(cudaFuncSetAttribute(sweep_bricks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, nsm = 0;
    MCEIK_CUDA(cudaGetDevice(&dev));
    MCEIK_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const long long ntasks = 8LL * a.nbricks * a.nfields_active;
    const int grid = (int)std::min<long long>((ntasks + kWarpsPerCta - 1) / kWarpsPerCta, nsm);
    sweep_bricks_kernel<<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    MCEIK_LAUNCH_CHECK();
}

void BrickPlan::build(int nx_, int ny_, int nz_, int zc_, cudaStream_t st) {
    if (nx_ == nx && ny_ == ny && nz_ == nz && zc_ == zc && nbricks > 0) return;
    nx = nx_; ny = ny_; nz = nz_; zc = zc_;
    nbx = (nx + kBx - 1) / kBx;
    nby = (ny + kBy - 1) / kBy;
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
    nx = ny = nz = zc = 0;
}

}  // namespace fsm
}  // namespace mceik
