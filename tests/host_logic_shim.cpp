// tests/host_logic_shim.cpp -- exposes mceik_b200/csrc/host_logic.hpp (host-only scalar logic of the
// boundary conditions) to the CPU tests.  Built on the fly by tests/test_host_logic.py with g++.
#include "host_logic.hpp"
extern "C" int shim_bc_records(int nx, int ny, int nz, double h, double x0, double y0, double z0, int nsrc,
                               const double *ts, const double *xs, const double *ys, const double *zs, int cap,
                               int *node, double *d, double *t, int *colloc) {
    std::vector<mceik::fsm::BcRecord> r;
    const int ierr = mceik::host::build_bc_records(nx, ny, nz, h, x0, y0, z0, nsrc, ts, xs, ys, zs, r);
    if (ierr) return -1;
    for (size_t i = 0; i < r.size() && (int)i < cap; ++i) { node[i] = r[i].node; d[i] = r[i].d; t[i] = r[i].ts; colloc[i] = r[i].collocated; }
    return (int)r.size();
}
