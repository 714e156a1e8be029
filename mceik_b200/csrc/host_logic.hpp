// host_logic.hpp -- scalar host-side logic of the eikonal boundary conditions.
//
// The per-source stencil selection of the reference is integer / scalar work on at most 27
// nodes per source; it runs on the host and only its result (node, distance, source time) is
// shipped to the device, where the slowness lives and ts + d*slow(node) is evaluated.
// Compiled with -ffp-contract=off so the fp64 arithmetic is the reference's (no FMA).
#pragma once
#include <cmath>
#include <algorithm>
#include <vector>
#include "fsm.cuh"

namespace mceik {
namespace host {

// EIKONAL_SOURCE_INDEX (fsm3d.f90:697-711): nearest node, 1-based.
inline int source_index(int n, double x0, double dx, double xs) {
    if (xs <= x0) return 1;
    if (xs >= x0 + (double)(float)(n - 1) * dx) return n;
    return (int)((xs - x0) / dx + 0.5) + 1;
}

// EIKONAL_INIT_GRID (fsm3d.f90:716-755): 2- or 3-node stencil along one axis (1-based, -1 =
// unused).  Quirks kept on purpose: the `isx > 0` test is always true, so a source exactly on
// node 1 asks for node 0 and fails; `isx < nx-1` drops the upper neighbour at isx = nx-1.
inline int init_grid(int n, int is, double x0, double dx, double xs, int loc[3]) {
    int np = 0, ierr = 0;
    loc[0] = loc[1] = loc[2] = -1;
    const double est = x0 + (double)(float)(is - 1) * dx;
    if (est > xs) {
        loc[0] = is - 1; loc[1] = is; np = 2;
    } else if (est < xs) {
        loc[0] = is; loc[1] = is + 1; np = 2;
    } else {
        if (is > 0) loc[np++] = is - 1;
        loc[np++] = is;
        if (is < n - 1) loc[np++] = is + 1;
    }
    for (int i = 0; i < np; ++i)
        if (loc[i] < 1 || loc[i] > n) ierr = 1;
    return ierr;
}

// Stencil records of all sources of one field, in the loop order of EIKONAL3D_SETBCS
// (fsm3d.f90:791-834: sources, then x, y, z stencil positions).  Returns the Fortran ierr.
inline int build_bc_records(int nx, int ny, int nz, double h, double x0, double y0, double z0, int nsrc,
                            const double *ts, const double *xs, const double *ys, const double *zs,
                            std::vector<fsm::BcRecord> &out) {
    const long nxy = (long)nx * ny;
    for (int s = 0; s < nsrc; ++s) {
        int lx[3], ly[3], lz[3];
        if (init_grid(nx, source_index(nx, x0, h, xs[s]), x0, h, xs[s], lx)) return 1;
        if (init_grid(ny, source_index(ny, y0, h, ys[s]), y0, h, ys[s], ly)) return 1;
        if (init_grid(nz, source_index(nz, z0, h, zs[s]), z0, h, zs[s], lz)) return 1;
        for (int i = 0; i < 3; ++i) {
            if (lx[i] == -1) continue;
            for (int j = 0; j < 3; ++j) {
                if (ly[j] == -1) continue;
                for (int k = 0; k < 3; ++k) {
                    if (lz[k] == -1) continue;
                    const int ix = lx[i], iy = ly[j], iz = lz[k];
                    const double x = x0 + (double)(float)(ix - 1) * h;
                    const double y = y0 + (double)(float)(iy - 1) * h;
                    const double z = z0 + (double)(float)(iz - 1) * h;
                    const double ex = xs[s] - x, ey = ys[s] - y, ez = zs[s] - z;
                    fsm::BcRecord r;
                    r.d = std::sqrt(ex * ex + ey * ey + ez * ez);
                    r.ts = ts[s];
                    r.node = (int)((long)(iz - 1) * nxy + (long)(iy - 1) * nx + (ix - 1));
                    r.collocated = std::fabs(r.d) < 1.e-10 ? 1 : 0;
                    out.push_back(r);
                }
            }
        }
    }
    return 0;
}


// Ragged catalogues on the fast kernel.  locate_uniform_kernel needs the 8 events of a block to use the same table
// in every pick slot (unused picks are wildcards).  A block whose events list their used picks in strictly
// increasing table order is re-laid out over the union of its tables: every event gets one slot per table of the
// union, "unused" (-1) where it has no pick.  Each event's used picks keep their own order, and unused picks never
// touch an accumulator, so the results are the same bits.  Other blocks are copied unchanged (general kernel).
constexpr int kMaxAlignedPicks = 392;  // = gs::kUniformMaxPicks: a wider union would not fit the fast kernel
inline void align_event_blocks(int EB, int nevents, const int *optr, const int *tid, const double *tobs, const double *var,
                        std::vector<int> &optr2, std::vector<int> &tid2, std::vector<double> &tobs2,
                        std::vector<double> &var2) {
    optr2.assign(1, 0);
    tid2.clear(); tobs2.clear(); var2.clear();
    std::vector<int> uni;
    for (int e0 = 0; e0 < nevents; e0 += EB) {
        const int e1 = std::min(e0 + EB, nevents);
        int maxp = 0;
        for (int e = e0; e < e1; ++e) maxp = std::max(maxp, optr[e + 1] - optr[e]);
        bool uniform = true, sorted = true;
        for (int j = 0; j < maxp && uniform; ++j) {
            int id = -1;
            for (int e = e0; e < e1; ++e)
                if (j < optr[e + 1] - optr[e]) {
                    const int t = tid[optr[e] + j];
                    if (t >= 0) { if (id < 0) id = t; else if (id != t) uniform = false; }
                }
        }
        uni.clear();
        if (!uniform) {
            for (int e = e0; e < e1 && sorted; ++e) {
                int last = -1;
                for (int p = optr[e]; p < optr[e + 1]; ++p) {
                    if (tid[p] < 0) continue;
                    if (tid[p] <= last) { sorted = false; break; }
                    last = tid[p];
                    uni.push_back(tid[p]);
                }
            }
            std::sort(uni.begin(), uni.end());
            uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
        }
        if (uniform || !sorted || (int)uni.size() > kMaxAlignedPicks) {
            for (int e = e0; e < e1; ++e) {
                for (int p = optr[e]; p < optr[e + 1]; ++p) { tid2.push_back(tid[p]); tobs2.push_back(tobs[p]); var2.push_back(var[p]); }
                optr2.push_back((int)tid2.size());
            }
            continue;
        }
        for (int e = e0; e < e1; ++e) {
            int p = optr[e];
            for (int u : uni) {
                while (p < optr[e + 1] && tid[p] < 0) ++p;
                if (p < optr[e + 1] && tid[p] == u) { tid2.push_back(u); tobs2.push_back(tobs[p]); var2.push_back(var[p]); ++p; }
                else { tid2.push_back(-1); tobs2.push_back(0.0); var2.push_back(1.0); }
            }
            optr2.push_back((int)tid2.size());
        }
    }
}


// Ticket table of the bricks16 queue (BrickArgs::vptr): the active fields form two groups, group 1 runs `stagger`
// brick levels behind group 0; virtual level V holds level V % nblevels of sweep V / nblevels for group 0 and the
// same of V - stagger for group 1.  vptr[V] = number of tickets before V; vptr.back() = 8 * nbricks * nfields.
inline std::vector<long long> build_ticket_table(int nblevels, const int *blevel_ptr, int nfields, int nf0, int stagger) {
    const int nvl = 8 * nblevels + stagger, nf1 = nfields - nf0;
    std::vector<long long> vptr(nvl + 1, 0);
    for (int v = 0; v < nvl; ++v) {
        long long cnt = 0;
        if (v < 8 * nblevels) cnt += (long long)nf0 * (blevel_ptr[v % nblevels + 1] - blevel_ptr[v % nblevels]);
        const int v1 = v - stagger;
        if (nf1 > 0 && v1 >= 0 && v1 < 8 * nblevels) cnt += (long long)nf1 * (blevel_ptr[v1 % nblevels + 1] - blevel_ptr[v1 % nblevels]);
        vptr[v + 1] = vptr[v] + cnt;
    }
    return vptr;
}

}  // namespace host
}  // namespace mceik
