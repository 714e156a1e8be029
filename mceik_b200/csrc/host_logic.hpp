// host_logic.hpp -- scalar host-side logic of the eikonal boundary conditions.
//
// The per-source stencil selection of the reference is integer / scalar work on at most 27
// nodes per source; it runs on the host and only its result (node, distance, source time) is
// shipped to the device, where the slowness lives and ts + d*slow(node) is evaluated.
// Compiled with -ffp-contract=off so the fp64 arithmetic is the reference's (no FMA).
#pragma once
#include <cmath>
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

}  // namespace host
}  // namespace mceik
