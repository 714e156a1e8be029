// fsm.cu -- batched fast-sweeping eikonal solver for sm_100a.
//
// What the reference does (serial path): for every iteration, 8 Gauss-Seidel sweeps over the
// whole grid, each sweep visiting the hyperplanes ix+iy+iz = const of the (possibly mirrored)
// grid in ascending order (EIKONAL3D_FSM fsm3d.f90:62-85, EVAL_UPDATE3D :419-456); a node on
// level L reads its six neighbours, which lie on levels L-1 (already updated in this sweep)
// and L+1 (not yet), so any execution order that respects "minus-side neighbour first,
// plus-side neighbour after" reproduces the reference bit for bit.
//
// How it is done here: the grid is cut into 16^3 tiles.  A tile task = (group of up to 4
// fields sharing a slowness model, sweep, tile): the CTA stages the tile + 1-node halo of each
// field and the slowness tile in shared memory, runs the tile's 46 local hyperplanes with one
// named barrier per hyperplane and field, and writes the interior back.  Tile tasks form a
// DAG (upwind tiles of the same sweep first; the same tile and its 6 neighbours of the
// previous sweep first) that a persistent kernel walks with a ticket counter and per-tile
// completion counters: tickets are handed out in a topological order, so a task only ever
// waits for tickets that are already running or done -- no co-residency assumption, no grid
// barrier, and consecutive sweeps pipeline across the grid.
#include <algorithm>
#include <vector>
#include "fsm.cuh"
#include "fsm_solve.cuh"

namespace mceik {
namespace fsm {

__constant__ int c_lvl_ptr[kTileLevels + 1];

// ------------------------------------------------------------------------------------------
// fill + boundary conditions (EIKONAL3D_SETBCS, fsm3d.f90:782-834)
// ------------------------------------------------------------------------------------------
__global__ void fill_kernel(double2 *__restrict__ p2, size_t n2, double *__restrict__ tail, int ntail, double v) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) p2[i] = make_double2(v, v);
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = v;
}

void launch_fill(double *d_u, size_t n, double value, cudaStream_t st) {
    if (n == 0) return;
    // d_u comes from cudaMalloc (256-byte aligned) or a torch tensor (>= 16-byte aligned)
    size_t n2 = n / 2;
    int ntail = (int)(n - 2 * n2);
    int blocks = (int)std::min<size_t>((n2 + 255) / 256 + 1, 148 * 16);
    fill_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<double2 *>(d_u), n2, d_u + 2 * n2, ntail, value);
    MCEIK_LAUNCH_CHECK();
}

__global__ void apply_bcs_kernel(int nfields, size_t n, const int *__restrict__ field_model,
                                 const int *__restrict__ rec_ptr, const BcRecord *__restrict__ recs,
                                 const double *__restrict__ slow, double *__restrict__ u) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfields) return;
    const double *sl = slow + (size_t)field_model[f] * n;
    double *uf = u + (size_t)f * n;
    for (int r = rec_ptr[f]; r < rec_ptr[f + 1]; ++r) {
        const BcRecord rec = recs[r];
        const double t = __dadd_rn(rec.ts, __dmul_rn(rec.d, sl[rec.node]));  // ts + d*slow (:826,828)
        if (rec.collocated)
            uf[rec.node] = t;
        else
            uf[rec.node] = fmin(uf[rec.node], t);
    }
}

void launch_apply_bcs(int nfields, size_t n, const int *d_field_model, const int *d_rec_ptr,
                      const BcRecord *d_recs, const double *d_slow, double *d_u, cudaStream_t st) {
    if (nfields == 0) return;
    apply_bcs_kernel<<<(nfields + 63) / 64, 64, 0, st>>>(nfields, n, d_field_model, d_rec_ptr, d_recs, d_slow, d_u);
    MCEIK_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// Tile-wavefront sweep kernel: one launch = one FSM iteration (8 sweeps) of all active groups.
// ------------------------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(B *kGroupThreads, 1) sweep_tiles_kernel(const SweepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *us = reinterpret_cast<double *>(smem_raw);               // [B][kHaloNodes]
    double *fs = us + (size_t)B * kHaloNodes;                        // [kTileNodes] slow*h
    uint16_t *lvl = reinterpret_cast<uint16_t *>(fs + kTileNodes);   // [kTileNodes]
    uint32_t *bcmask = reinterpret_cast<uint32_t *>(lvl + kTileNodes);  // [B][kTileNodes/32]
    __shared__ int s_task;
    __shared__ int s_hasbc[B];

    const int tid = threadIdx.x;
    const int slot = tid / kGroupThreads;
    const int gt = tid - slot * kGroupThreads;
    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const size_t nxy = (size_t)nx * ny;
    const size_t N = nxy * nz;
    const int tasks_per_sweep = a.ngroups * a.ntiles;
    const int ntasks = 8 * tasks_per_sweep;

    for (int i = tid; i < kTileNodes; i += B * kGroupThreads) lvl[i] = a.lvl_nodes[i];

    while (true) {
        __syncthreads();  // previous task fully retired (smem reusable, s_task consumed)
        if (tid == 0) s_task = atomicAdd(a.queue, 1);
        __syncthreads();
        const int t = s_task;
        if (t >= ntasks) break;

        // ---- decode ticket -> (sweep, tile level, group, tile); order [sweep][tile level][group][tile]
        const int s = t / tasks_per_sweep;
        int r = t - s * tasks_per_sweep;
        int lo = 0, hi = a.ntlevels;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (a.ngroups * __ldg(a.tlevel_ptr + mid) <= r) lo = mid; else hi = mid;
        }
        const int p0 = __ldg(a.tlevel_ptr + lo);
        const int nl = __ldg(a.tlevel_ptr + lo + 1) - p0;
        r -= a.ngroups * p0;
        const int g = r / nl;
        const int packed = __ldg(a.tile_order + p0 + (r - g * nl));
        const bool revx = (s & 1) != 0, revy = (s & 2) != 0, revz = (s & 4) != 0;  // sweep table fsm3d.f90:46-53
        int I = packed & 1023, J = (packed >> 10) & 1023, K = packed >> 20;
        if (revx) I = a.ntx - 1 - I;
        if (revy) J = a.nty - 1 - J;
        if (revz) K = a.ntz - 1 - K;
        const int tile = (K * a.nty + J) * a.ntx + I;
        int *done_g = a.done + (size_t)g * a.ntiles;

        // ---- wait for the dependencies: this tile and its 6 neighbours have finished sweep s-1,
        //      the (up to 3) upwind neighbours have finished sweep s.
        if (tid < 7) {
            int di = 0, dj = 0, dk = 0;
            if (tid == 1) di = -1; else if (tid == 2) di = 1;
            else if (tid == 3) dj = -1; else if (tid == 4) dj = 1;
            else if (tid == 5) dk = -1; else if (tid == 6) dk = 1;
            const int NI = I + di, NJ = J + dj, NK = K + dk;
            if (NI >= 0 && NI < a.ntx && NJ >= 0 && NJ < a.nty && NK >= 0 && NK < a.ntz) {
                int need = s;
                if ((di != 0 && di == (revx ? 1 : -1)) || (dj != 0 && dj == (revy ? 1 : -1)) ||
                    (dk != 0 && dk == (revz ? 1 : -1)))
                    need = s + 1;
                const int *p = done_g + ((NK * a.nty + NJ) * a.ntx + NI);
                while (ld_acquire_gpu(p) < need) __nanosleep(64);
            }
        }
        __syncthreads();

        const int x_lo = I * kTile, y_lo = J * kTile, z_lo = K * kTile;
        const int ex = min(kTile, nx - x_lo), ey = min(kTile, ny - y_lo), ez = min(kTile, nz - z_lo);

        // ---- stage f = slow*h for the tile (UPDATE3D: fijkh = slow(ijk)*h, fsm3d.f90:472)
        {
            const double *sl = a.slow + (size_t)__ldg(a.group_model + g) * N;
            for (int idx = tid; idx < kTileNodes; idx += B * kGroupThreads) {
                const int i = idx & 15, j = (idx >> 4) & 15, k = idx >> 8;
                const int gx = min(x_lo + i, nx - 1), gy = min(y_lo + j, ny - 1), gz = min(z_lo + k, nz - 1);
                fs[idx] = __dmul_rn(__ldg(sl + (size_t)gz * nxy + (size_t)gy * nx + gx), a.h);
            }
        }

        const int f = __ldg(a.group_fields + g * kMaxSlots + slot);
        double *myus = us + (size_t)slot * kHaloNodes;
        uint32_t *mymask = bcmask + slot * (kTileNodes / 32);
        if (f >= 0) {
            // ---- stage u: tile + halo.  A halo cell outside the grid takes the value of the
            //      adjacent boundary node itself, which is what GET_U?MIN3D substitutes at a
            //      face (fsm3d.f90:495-499, 517-521, 539-543).
            double *uf = a.u + (size_t)f * N;
            double *u0f = a.u0 + (size_t)f * N;
            for (int idx = gt; idx < kHaloNodes; idx += kGroupThreads) {
                const int kk = idx / (kHalo * kHalo);
                const int rem = idx - kk * (kHalo * kHalo);
                const int jj = rem / kHalo;
                const int ii = rem - jj * kHalo;
                const int i = ii - 1, j = jj - 1, k = kk - 1;
                const int gx = min(max(x_lo + i, 0), nx - 1);
                const int gy = min(max(y_lo + j, 0), ny - 1);
                const int gz = min(max(z_lo + k, 0), nz - 1);
                const size_t gi = (size_t)gz * nxy + (size_t)gy * nx + gx;
                const double v = __ldcg(uf + gi);
                myus[idx] = v;
                if (s == 0 && i >= 0 && i < ex && j >= 0 && j < ey && k >= 0 && k < ez) __stcg(u0f + gi, v);
            }
            // ---- boundary-condition nodes inside this tile are never updated (lupd, fsm3d.f90:2029-2032)
            if (gt < kTileNodes / 32) mymask[gt] = 0u;
            if (gt == 0) s_hasbc[slot] = 0;
            named_bar_sync(slot + 1, kGroupThreads);
            for (int n = __ldg(a.bc_ptr + f) + gt; n < __ldg(a.bc_ptr + f + 1); n += kGroupThreads) {
                if (__ldg(a.bc_tile + n) == tile) {
                    const int p = a.bc_local[n];
                    const int li = ((p >> 8) * kTile + ((p >> 4) & 15)) * kTile + (p & 15);
                    atomicOr(mymask + (li >> 5), 1u << (li & 31));
                    s_hasbc[slot] = 1;
                }
            }
            named_bar_sync(slot + 1, kGroupThreads);  // also publishes myus to the whole group
        }
        __syncthreads();  // fs complete (written by all groups)

        if (f >= 0) {
            const bool hasbc = s_hasbc[slot] != 0;
            const int nlev = ex + ey + ez - 2;
            for (int lev = 0; lev < nlev; ++lev) {
                const int end = c_lvl_ptr[lev + 1];
                for (int n = c_lvl_ptr[lev] + gt; n < end; n += kGroupThreads) {
                    const int p = lvl[n];
                    const int aa = p & 15, bb = (p >> 4) & 15, cc = p >> 8;
                    if (aa < ex && bb < ey && cc < ez) {
                        const int i = revx ? ex - 1 - aa : aa;
                        const int j = revy ? ey - 1 - bb : bb;
                        const int k = revz ? ez - 1 - cc : cc;
                        const int li = (k * kTile + j) * kTile + i;
                        if (hasbc && ((mymask[li >> 5] >> (li & 31)) & 1u)) continue;
                        const int si = ((k + 1) * kHalo + (j + 1)) * kHalo + (i + 1);
                        const double uc = myus[si];
                        const double ux = fmin(myus[si - 1], myus[si + 1]);
                        const double uy = fmin(myus[si - kHalo], myus[si + kHalo]);
                        const double uz = fmin(myus[si - kHalo * kHalo], myus[si + kHalo * kHalo]);
                        const double x = local_solve(ux, uy, uz, fs[li]);
                        if (x < uc) myus[si] = x;  // u = MIN(u, ubar) (fsm3d.f90:477)
                    }
                }
                named_bar_sync(slot + 1, kGroupThreads);
            }

            // ---- write the interior back; on the last sweep fold in the convergence test
            //      (fsm3d.f90:86-90): count nodes with NOT(|u0-u| < tol).
            double *uf = a.u + (size_t)f * N;
            const double *u0f = a.u0 + (size_t)f * N;
            unsigned int bad = 0;
            for (int idx = gt; idx < kTileNodes; idx += kGroupThreads) {
                const int i = idx & 15, j = (idx >> 4) & 15, k = idx >> 8;
                if (i < ex && j < ey && k < ez) {
                    const size_t gi = (size_t)(z_lo + k) * nxy + (size_t)(y_lo + j) * nx + (x_lo + i);
                    const double v = myus[((k + 1) * kHalo + (j + 1)) * kHalo + (i + 1)];
                    __stcg(uf + gi, v);
                    if (s == 7) {
                        const double o = __ldcg(u0f + gi);
                        if (!(fabs(__dsub_rn(o, v)) < a.tol)) ++bad;
                    }
                }
            }
            if (s == 7) {
                for (int off = 16; off > 0; off >>= 1) bad += __shfl_down_sync(0xffffffffu, bad, off);
                if ((tid & 31) == 0 && bad) atomicAdd(a.nonconv + f, (unsigned long long)bad);
            }
        }

        __syncthreads();
        if (tid == 0) {
            __threadfence();
            red_release_gpu_add(done_g + tile, 1);
        }
    }
}

// ------------------------------------------------------------------------------------------
// self-test: sqrt_fast(x) == __dsqrt_rn(x) bit for bit, and local_solve_sl == local_solve
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix(unsigned long long &z) {
    z += 0x9e3779b97f4a7c15ULL;
    unsigned long long x = z;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

__global__ void selftest_kernel(unsigned long long seed, int per_thread, unsigned long long *bad) {
    unsigned long long z = seed + 0x632be59bd9b4e019ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x);
    unsigned long long nbad_sqrt = 0, nbad_solve = 0;
    for (int it = 0; it < per_thread; ++it) {
        // sqrt: random mantissa, exponent uniform over the admitted range, plus near-square inputs
        const unsigned long long r = splitmix(z);
        const int ex = 64 + (int)(splitmix(z) % 1982);  // biased exponent in [64, 2045]
        double x = __longlong_as_double((r & 0x000fffffffffffffULL) | ((unsigned long long)ex << 52));
        if ((it & 3) == 3) { const double q = 1.0 + (double)(r >> 12) * 0x1p-52; x = q * q; }
        if (x >= MCEIK_SQRT_FAST_MIN && x < DBL_MAX) {
            if (__double_as_longlong(sqrt_fast(x)) != __double_as_longlong(__dsqrt_rn(x))) ++nbad_sqrt;
        }
        // solver: three neighbour times around a base value, f comparable to their spread
        const double base = (double)(splitmix(z) >> 40) * 1e-4;
        const double f = 1e-3 + (double)(splitmix(z) >> 44) * 1e-6;
        double n[3];
        for (int q = 0; q < 3; ++q) {
            const unsigned long long w = splitmix(z);
            n[q] = ((w & 15) == 0) ? DBL_MAX : base + f * 4.0 * ((double)(w >> 11) * 0x1p-53);
            if ((w & 0xf0) == 0x10) n[q] = n[(q + 2) % 3 < q ? (q + 2) % 3 : 0];  // exact ties
        }
        bool rare;
        double a = local_solve_sl(n[0], n[1], n[2], f, rare);
        if (rare) a = local_solve(n[0], n[1], n[2], f);
        if (__double_as_longlong(a) != __double_as_longlong(local_solve(n[0], n[1], n[2], f))) ++nbad_solve;
    }
    if (nbad_sqrt) atomicAdd(bad, nbad_sqrt);
    if (nbad_solve) atomicAdd(bad + 1, nbad_solve);
}

void launch_selftest(unsigned long long seed, int blocks, int per_thread, unsigned long long *d_bad, cudaStream_t st) {
    selftest_kernel<<<blocks, 256, 0, st>>>(seed, per_thread, d_bad);
    MCEIK_LAUNCH_CHECK();
}

size_t tiles_smem_bytes(int nslots) {
    return (size_t)nslots * kHaloNodes * sizeof(double) + kTileNodes * sizeof(double) +
           kTileNodes * sizeof(uint16_t) + (size_t)nslots * (kTileNodes / 32) * sizeof(uint32_t);
}

template <int B>
static void launch_tiles_impl(const SweepArgs &a, cudaStream_t st) {
    const size_t smem = tiles_smem_bytes(B);
    MCEIK_CUDA(cudaFuncSetAttribute(sweep_tiles_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, nsm = 0, occ = 0;
    MCEIK_CUDA(cudaGetDevice(&dev));
    MCEIK_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    MCEIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_tiles_kernel<B>, B * kGroupThreads, smem));
    if (occ < 1) throw CudaError("sweep_tiles_kernel does not fit on an SM");
    const long long ntasks = 8LL * a.ngroups * a.ntiles;
    const int grid = (int)std::min<long long>(ntasks, (long long)nsm * occ);
    sweep_tiles_kernel<B><<<grid, B * kGroupThreads, smem, st>>>(a);
    MCEIK_LAUNCH_CHECK();
}

void launch_iteration_tiles(const SweepArgs &a, cudaStream_t st) {
    if (a.ngroups == 0) return;
    switch (a.nslots) {
        case 1: launch_tiles_impl<1>(a, st); break;
        case 2: launch_tiles_impl<2>(a, st); break;
        case 4: launch_tiles_impl<4>(a, st); break;
        default: throw CudaError("nslots must be 1, 2 or 4");
    }
}

// ------------------------------------------------------------------------------------------
// Tile plan
// ------------------------------------------------------------------------------------------
void TilePlan::build(int nx_, int ny_, int nz_, cudaStream_t st) {
    if (nx_ == nx && ny_ == ny && nz_ == nz && ntiles > 0) return;
    nx = nx_; ny = ny_; nz = nz_;
    ntx = (nx + kTile - 1) / kTile;
    nty = (ny + kTile - 1) / kTile;
    ntz = (nz + kTile - 1) / kTile;
    if (ntx > 1023 || nty > 1023 || ntz > 1023) throw CudaError("grid too large for the tile plan");
    ntiles = ntx * nty * ntz;
    ntlevels = ntx + nty + ntz - 2;
    std::vector<int> order;
    order.reserve(ntiles);
    h_tlevel_ptr.assign(ntlevels + 1, 0);
    for (int d = 0; d < ntlevels; ++d) {
        h_tlevel_ptr[d] = (int)order.size();
        for (int K = 0; K < ntz; ++K)
            for (int J = 0; J < nty; ++J) {
                const int I = d - K - J;
                if (I >= 0 && I < ntx) order.push_back(I | (J << 10) | (K << 20));
            }
    }
    h_tlevel_ptr[ntlevels] = (int)order.size();
    MCEIK_CUDA(cudaMemcpyAsync(tile_order.ensure(sizeof(int) * ntiles), order.data(), sizeof(int) * ntiles,
                               cudaMemcpyHostToDevice, st));
    MCEIK_CUDA(cudaMemcpyAsync(tlevel_ptr.ensure(sizeof(int) * (ntlevels + 1)), h_tlevel_ptr.data(),
                               sizeof(int) * (ntlevels + 1), cudaMemcpyHostToDevice, st));
    // intra-tile hyperplanes: nodes (a,b,c) sorted by a+b+c, then c, then b
    std::vector<uint16_t> nodes;
    std::vector<int> lptr(kTileLevels + 1, 0);
    nodes.reserve(kTileNodes);
    for (int l = 0; l < kTileLevels; ++l) {
        lptr[l] = (int)nodes.size();
        for (int c = 0; c < kTile; ++c)
            for (int b = 0; b < kTile; ++b) {
                const int aa = l - b - c;
                if (aa >= 0 && aa < kTile) nodes.push_back((uint16_t)(aa | (b << 4) | (c << 8)));
            }
    }
    lptr[kTileLevels] = (int)nodes.size();
    MCEIK_CUDA(cudaMemcpyAsync(lvl_nodes.ensure(sizeof(uint16_t) * kTileNodes), nodes.data(),
                               sizeof(uint16_t) * kTileNodes, cudaMemcpyHostToDevice, st));
    // constant memory is per device and a plan is per context (= per device): upload with every (re)build
    MCEIK_CUDA(cudaMemcpyToSymbolAsync(c_lvl_ptr, lptr.data(), sizeof(int) * (kTileLevels + 1), 0,
                                       cudaMemcpyHostToDevice, st));
    MCEIK_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
}

void TilePlan::release() {
    tile_order.release();
    tlevel_ptr.release();
    lvl_nodes.release();
    ntiles = 0;
    nx = ny = nz = 0;
}

// ------------------------------------------------------------------------------------------
// Cross-check path: one launch per hyperplane, global memory only (same arithmetic).
// Thread = (iy, iz) pair of the mirrored grid; ix follows from the level.
// ------------------------------------------------------------------------------------------
__global__ void sweep_level_kernel(int nx, int ny, int nz, double h, int level, int revx, int revy, int revz,
                                   int kz1, const int *__restrict__ active_fields,
                                   const int *__restrict__ field_model, const double *__restrict__ slow,
                                   const uint8_t *__restrict__ lupd, double *__restrict__ u) {
    int iy = blockIdx.x * blockDim.x + threadIdx.x;
    int iz = kz1 + blockIdx.y;
    int ix = level - iy - iz;
    if (iy >= ny || ix < 0 || ix >= nx) return;
    if (revx) ix = nx - 1 - ix;
    if (revy) iy = ny - 1 - iy;
    if (revz) iz = nz - 1 - iz;
    const int f = active_fields[blockIdx.z];
    const size_t nxy = (size_t)nx * ny, N = nxy * nz;
    const size_t ijk = (size_t)iz * nxy + (size_t)iy * nx + ix;
    if (!lupd[(size_t)f * N + ijk]) return;
    double *uf = u + (size_t)f * N;
    const double uc = uf[ijk];
    const double ux = fmin(ix > 0 ? uf[ijk - 1] : uc, ix < nx - 1 ? uf[ijk + 1] : uc);
    const double uy = fmin(iy > 0 ? uf[ijk - nx] : uc, iy < ny - 1 ? uf[ijk + nx] : uc);
    const double uz = fmin(iz > 0 ? uf[ijk - nxy] : uc, iz < nz - 1 ? uf[ijk + nxy] : uc);
    const double fh = __dmul_rn(slow[(size_t)field_model[f] * N + ijk], h);
    const double x = local_solve(ux, uy, uz, fh);
    if (x < uc) uf[ijk] = x;
}

void launch_iteration_levels(int nx, int ny, int nz, double h, int nfields, const int *d_active_fields,
                             const int *d_field_model, const double *d_slow, const uint8_t *d_lupd,
                             double *d_u, cudaStream_t st) {
    if (nfields == 0) return;
    const int nlevels = nx + ny + nz - 2;
    for (int s = 0; s < 8; ++s)
        for (int level = 0; level < nlevels; ++level) {
            const int kz1 = std::max(0, level - (nx - 1) - (ny - 1));
            const int kz2 = std::min(nz - 1, level);
            dim3 grid((ny + 127) / 128, kz2 - kz1 + 1, nfields);
            sweep_level_kernel<<<grid, 128, 0, st>>>(nx, ny, nz, h, level, s & 1, (s >> 1) & 1, (s >> 2) & 1, kz1,
                                                     d_active_fields, d_field_model, d_slow, d_lupd, d_u);
            MCEIK_LAUNCH_CHECK();
        }
}

__global__ void convergence_kernel(size_t n, const int *__restrict__ active_fields, double tol,
                                   const double *__restrict__ u, double *__restrict__ u0,
                                   unsigned long long *__restrict__ nonconv) {
    const int f = active_fields[blockIdx.y];
    const double *uf = u + (size_t)f * n;
    double *u0f = u0 + (size_t)f * n;
    unsigned int bad = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double v[4], o[4];  // four independent streams per thread keep enough bytes in flight
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const size_t i = i0 + q * stride;
            if (i < n) { v[q] = __ldcs(uf + i); o[q] = __ldcs(u0f + i); }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const size_t i = i0 + q * stride;
            if (i < n) {
                if (!(fabs(__dsub_rn(o[q], v[q])) < tol)) ++bad;
                __stcs(u0f + i, v[q]);
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) bad += __shfl_down_sync(0xffffffffu, bad, off);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(nonconv + f, (unsigned long long)bad);
}

void launch_convergence(size_t n, int nfields, const int *d_active_fields, double tol, const double *d_u,
                        double *d_u0, unsigned long long *d_nonconv, cudaStream_t st) {
    if (nfields == 0) return;
    dim3 grid((unsigned)std::min<size_t>((n + 1023) / 1024, 148 * 4), nfields);
    convergence_kernel<<<grid, 256, 0, st>>>(n, d_active_fields, tol, d_u, d_u0, d_nonconv);
    MCEIK_LAUNCH_CHECK();
}

__global__ void scale_slowness_kernel(size_t n, double h, const double *__restrict__ slow, double *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __dmul_rn(slow[i], h);
}

void launch_scale_slowness(size_t n, double h, const double *d_slow, double *d_out, cudaStream_t st) {
    if (n == 0) return;
    scale_slowness_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(n, h, d_slow, d_out);
    MCEIK_LAUNCH_CHECK();
}

// ---- blocked layout of the bricks16 kernel: [field][brick column][z][80] (fsm.cuh, BrickArgs::blocked)
constexpr int kRecU = 80, kRecS = 64;
size_t blocked_field_doubles(int nx, int ny, int nz) { return (size_t)(nx / 8) * ((ny + 7) / 8) * nz * kRecU; }
size_t blocked_slowness_doubles(int nx, int ny, int nz) { return (size_t)(nx / 8) * ((ny + 7) / 8) * nz * kRecS; }

// one thread per record entry e in [0, 80): 0..63 node (j, i), 64..71 copy of column 0, 72..79 copy of column 7
__global__ void block_fields_kernel(int nx, int ny, int nz, const double *__restrict__ u, double *__restrict__ ub) {
    const int nbx = nx / 8, nby = (ny + 7) / 8;
    const size_t nxy = (size_t)nx * ny, N = nxy * nz, nrec = (size_t)nbx * nby * nz;
    const size_t total = nrec * kRecU, stride = (size_t)gridDim.x * blockDim.x;
    const double *uf = u + (size_t)blockIdx.y * N;
    double *ubf = ub + (size_t)blockIdx.y * total;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int e = (int)(t % kRecU);
        const size_t rec = t / kRecU;
        const int z = (int)(rec % nz);
        const size_t colid = rec / nz;
        const int I = (int)(colid % nbx), J = (int)(colid / nbx);
        const int j = e < 64 ? e >> 3 : (e - 64) & 7, i = e < 64 ? e & 7 : (e < 72 ? 0 : 7);
        const int y = J * 8 + j;
        ubf[t] = y < ny ? uf[(size_t)z * nxy + (size_t)y * nx + I * 8 + i] : DBL_MAX;
    }
}

// EIKONAL3D_SETBCS on the blocked layout: one thread per field applies its records in source order (fsm3d.f90:810-833)
// to the node's record entry and, for nodes of columns 0 / 7 of a brick, to the face copy
__global__ void apply_bcs_blocked_kernel(int nfields, int nx, int ny, int nz, const int *__restrict__ field_model,
                                         const int *__restrict__ rec_ptr, const BcRecord *__restrict__ recs,
                                         const double *__restrict__ slow, double *__restrict__ ub) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfields) return;
    const int nbx = nx / 8, nby = (ny + 7) / 8;
    const size_t nxy = (size_t)nx * ny, n = nxy * nz;
    const double *sl = slow + (size_t)field_model[f] * n;
    double *uf = ub + (size_t)f * ((size_t)nbx * nby * nz * kRecU);
    for (int r = rec_ptr[f]; r < rec_ptr[f + 1]; ++r) {
        const BcRecord rec = recs[r];
        const int x = (int)(rec.node % nx), y = (int)((rec.node / nx) % ny), z = (int)(rec.node / nxy);
        double *p = uf + (((size_t)(y >> 3) * nbx + (x >> 3)) * nz + z) * kRecU;
        const int e = (y & 7) * 8 + (x & 7);
        const double t = __dadd_rn(rec.ts, __dmul_rn(rec.d, sl[rec.node]));  // ts + d*slow (:826,828)
        const double v = rec.collocated ? t : fmin(p[e], t);
        p[e] = v;
        if ((x & 7) == 0) p[64 + (y & 7)] = v;
        if ((x & 7) == 7) p[72 + (y & 7)] = v;
    }
}

void launch_apply_bcs_blocked(int nfields, int nx, int ny, int nz, const int *d_field_model, const int *d_rec_ptr,
                              const BcRecord *d_recs, const double *d_slow, double *d_ub, cudaStream_t st) {
    if (nfields == 0) return;
    apply_bcs_blocked_kernel<<<(nfields + 63) / 64, 64, 0, st>>>(nfields, nx, ny, nz, d_field_model, d_rec_ptr, d_recs, d_slow, d_ub);
    MCEIK_LAUNCH_CHECK();
}

// convergence test on the blocked layout: the 64 nodes of every record (the face copies are duplicates; rows
// beyond ny hold u_nan in both arrays and compare equal), u0 = u for those entries (fsm3d.f90:86-90)
__global__ void convergence_blocked_kernel(size_t nrec, const int *__restrict__ active_fields, double tol,
                                           const double *__restrict__ u, double *__restrict__ u0,
                                           unsigned long long *__restrict__ nonconv) {
    const int f = active_fields[blockIdx.y];
    const double *uf = u + (size_t)f * nrec * kRecU;
    double *u0f = u0 + (size_t)f * nrec * kRecU;
    const size_t n2 = nrec * 32;  // pairs of nodes
    unsigned int bad = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n2; i0 += 4 * stride) {
        double2 v[4], o[4];  // four independent streams per thread keep enough bytes in flight
        size_t off[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const size_t i = i0 + q * stride;
            off[q] = (i >> 5) * kRecU + (i & 31) * 2;
            if (i < n2) { v[q] = __ldcs(reinterpret_cast<const double2 *>(uf + off[q])); o[q] = __ldcs(reinterpret_cast<const double2 *>(u0f + off[q])); }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (i0 + q * stride < n2) {
                if (!(fabs(__dsub_rn(o[q].x, v[q].x)) < tol)) ++bad;
                if (!(fabs(__dsub_rn(o[q].y, v[q].y)) < tol)) ++bad;
                __stcs(reinterpret_cast<double2 *>(u0f + off[q]), v[q]);
            }
        }
    }
    for (int off2 = 16; off2 > 0; off2 >>= 1) bad += __shfl_down_sync(0xffffffffu, bad, off2);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(nonconv + f, (unsigned long long)bad);
}

void launch_convergence_blocked(int nx, int ny, int nz, int nfields, const int *d_active_fields, double tol, const double *d_ub,
                                double *d_u0b, unsigned long long *d_nonconv, cudaStream_t st) {
    if (nfields == 0) return;
    const size_t nrec = (size_t)(nx / 8) * ((ny + 7) / 8) * nz;
    dim3 grid((unsigned)std::min<size_t>((nrec * 32 + 1023) / 1024, 148 * 4), nfields);
    convergence_blocked_kernel<<<grid, 256, 0, st>>>(nrec, d_active_fields, tol, d_ub, d_u0b, d_nonconv);
    MCEIK_LAUNCH_CHECK();
}

void launch_block_fields(int nx, int ny, int nz, int nfields, const double *d_u, double *d_ub, cudaStream_t st) {
    if (nfields == 0) return;
    const size_t total = blocked_field_doubles(nx, ny, nz);
    dim3 grid((unsigned)std::min<size_t>((total + 255) / 256, 148 * 8), nfields);
    block_fields_kernel<<<grid, 256, 0, st>>>(nx, ny, nz, d_u, d_ub);
    MCEIK_LAUNCH_CHECK();
}

// blocked records -> the caller's layout: fp64 field (u != nullptr) and / or fp32 table (tab != nullptr; SNGL(u),
// fsm3d.f90:1870-1872) of the listed fields.  One thread per node: the writes are coalesced.
__global__ void unblock_fields_kernel(int nx, int ny, int nz, const int *__restrict__ fields, const double *__restrict__ ub,
                                      double *__restrict__ u, float *__restrict__ tab, size_t ldtab) {
    const int f = fields ? fields[blockIdx.y] : blockIdx.y;
    const int nbx = nx / 8, nby = (ny + 7) / 8;
    const size_t nxy = (size_t)nx * ny, N = nxy * nz, stride = (size_t)gridDim.x * blockDim.x;
    const double *ubf = ub + (size_t)f * ((size_t)nbx * nby * nz * kRecU);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < N; t += stride) {
        const int x = (int)(t % nx), y = (int)((t / nx) % ny), z = (int)(t / nxy);
        const double v = __ldcs(ubf + (((size_t)(y >> 3) * nbx + (x >> 3)) * nz + z) * kRecU + (y & 7) * 8 + (x & 7));
        if (u) u[(size_t)f * N + t] = v;
        if (tab) tab[(size_t)f * ldtab + t] = __double2float_rn(v);
    }
}

void launch_unblock_fields(int nx, int ny, int nz, int nfields, const int *d_fields, const double *d_ub, double *d_u,
                           float *d_tables, size_t ldtab, cudaStream_t st) {
    if (nfields == 0 || (!d_u && !d_tables)) return;
    const size_t N = (size_t)nx * ny * nz;
    dim3 grid((unsigned)std::min<size_t>((N + 255) / 256, 148 * 8), nfields);
    unblock_fields_kernel<<<grid, 256, 0, st>>>(nx, ny, nz, d_fields, d_ub, d_u, d_tables, ldtab);
    MCEIK_LAUNCH_CHECK();
}

__global__ void scale_slowness_blocked_kernel(int nx, int ny, int nz, double h, const double *__restrict__ slow,
                                              double *__restrict__ out) {
    const int nbx = nx / 8, nby = (ny + 7) / 8;
    const size_t nxy = (size_t)nx * ny, N = nxy * nz, total = (size_t)nbx * nby * nz * kRecS;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const double *sf = slow + (size_t)blockIdx.y * N;
    double *of = out + (size_t)blockIdx.y * total;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int e = (int)(t % kRecS);
        const size_t rec = t / kRecS;
        const int z = (int)(rec % nz);
        const size_t colid = rec / nz;
        const int I = (int)(colid % nbx), J = (int)(colid / nbx), y = J * 8 + (e >> 3);
        of[t] = y < ny ? __dmul_rn(sf[(size_t)z * nxy + (size_t)y * nx + I * 8 + (e & 7)], h) : 1.0;
    }
}

void launch_scale_slowness_blocked(int nx, int ny, int nz, int nmodels, double h, const double *d_slow, double *d_out,
                                   cudaStream_t st) {
    if (nmodels == 0) return;
    const size_t total = blocked_slowness_doubles(nx, ny, nz);
    dim3 grid((unsigned)std::min<size_t>((total + 255) / 256, 148 * 8), nmodels);
    scale_slowness_blocked_kernel<<<grid, 256, 0, st>>>(nx, ny, nz, h, d_slow, d_out);
    MCEIK_LAUNCH_CHECK();
}

__global__ void mark_bcs_kernel(int nrec, const int *__restrict__ rec_field, const int *__restrict__ rec_node,
                                size_t n, uint8_t *__restrict__ lupd) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrec) lupd[(size_t)rec_field[r] * n + rec_node[r]] = 0;
}

void launch_mark_bcs(int nrec, const int *d_rec_field, const int *d_rec_node, size_t n, uint8_t *d_lupd,
                     cudaStream_t st) {
    if (nrec == 0) return;
    mark_bcs_kernel<<<(nrec + 127) / 128, 128, 0, st>>>(nrec, d_rec_field, d_rec_node, n, d_lupd);
    MCEIK_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// fp64 field -> fp32 table (SNGL(u), fsm3d.f90:1870-1872)
// ------------------------------------------------------------------------------------------
__global__ void pack_tables_kernel(size_t n, size_t ldtab, const double *__restrict__ u, float *__restrict__ tab) {
    const double *uf = u + (size_t)blockIdx.y * n;
    float *tf = tab + (size_t)blockIdx.y * ldtab;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        tf[i] = __double2float_rn(__ldcs(uf + i));
}

void launch_pack_tables(int nfields, size_t n, size_t ldtab, const double *d_u, float *d_tables, cudaStream_t st) {
    if (nfields == 0 || n == 0) return;
    dim3 grid((unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), nfields);
    pack_tables_kernel<<<grid, 256, 0, st>>>(n, ldtab, d_u, d_tables);
    MCEIK_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// analytic homogeneous tables (computeHomogeneousTraveltimes, homog.c:594-621, then
// double2FloatArray, homog.c:624-635).  xyzv = [nstations][4] = (xs, ys, zs, 1/vel).
// ------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void homog_tables_kernel(int nx, int ny, int nz, double x0, double y0, double z0, double dx, double dy,
                                    double dz, const double *__restrict__ xyzv, OutT *__restrict__ tab, size_t ldtab) {
    const int s = blockIdx.y;
    const double xs = xyzv[4 * s], ys = xyzv[4 * s + 1], zs = xyzv[4 * s + 2], slow = xyzv[4 * s + 3];
    const int iy = blockIdx.x % ny, iz = blockIdx.x / ny;
    const double ey = __dsub_rn(ys, __dadd_rn(y0, __dmul_rn((double)iy, dy)));
    const double ez = __dsub_rn(zs, __dadd_rn(z0, __dmul_rn((double)iz, dz)));
    OutT *row = tab + (size_t)s * ldtab + ((size_t)iz * ny + iy) * nx;
    for (int ix = threadIdx.x; ix < nx; ix += blockDim.x) {
        const double ex = __dsub_rn(xs, __dadd_rn(x0, __dmul_rn((double)ix, dx)));
        const double d = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez)));
        row[ix] = (OutT)__dmul_rn(d, slow);  // double -> float conversion is round-to-nearest
    }
}

template <typename OutT>
void launch_homog_tables(int nx, int ny, int nz, double x0, double y0, double z0, double dx, double dy, double dz,
                         int nstations, const double *d_xyzv, OutT *d_tables, size_t ldtab, cudaStream_t st) {
    if (nstations == 0) return;
    dim3 grid(ny * nz, nstations);
    homog_tables_kernel<OutT><<<grid, 128, 0, st>>>(nx, ny, nz, x0, y0, z0, dx, dy, dz, d_xyzv, d_tables, ldtab);
    MCEIK_LAUNCH_CHECK();
}
template void launch_homog_tables<float>(int, int, int, double, double, double, double, double, double, int,
                                         const double *, float *, size_t, cudaStream_t);
template void launch_homog_tables<double>(int, int, int, double, double, double, double, double, double, int,
                                          const double *, double *, size_t, cudaStream_t);

}  // namespace fsm
}  // namespace mceik
