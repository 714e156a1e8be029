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

// align_event_blocks: returns the number of picks after alignment; arrays sized by the caller (cap picks)
extern "C" int shim_align_event_blocks(int eb, int nevents, const int *optr, const int *tid, const double *tobs, const double *var,
                                       int cap, int *optr2, int *tid2, double *tobs2, double *var2) {
    std::vector<int> o, t;
    std::vector<double> a, b;
    mceik::host::align_event_blocks(eb, nevents, optr, tid, tobs, var, o, t, a, b);
    if ((int)t.size() > cap) return -1;
    for (int e = 0; e <= nevents; ++e) optr2[e] = o[e];
    for (size_t i = 0; i < t.size(); ++i) { tid2[i] = t[i]; tobs2[i] = a[i]; var2[i] = b[i]; }
    return (int)t.size();
}

// bricks16 ticket queue: table + decode (the decode is the same __host__ __device__ function the kernel calls)
extern "C" long long shim_ticket_table(int nblevels, const int *blevel_ptr, int nfields, int nf0, int stagger, long long *vptr) {
    const std::vector<long long> v = mceik::host::build_ticket_table(nblevels, blevel_ptr, nfields, nf0, stagger);
    for (size_t i = 0; i < v.size(); ++i) vptr[i] = v[i];
    return v.back();
}
extern "C" void shim_decode_ticket(long long t, const long long *vptr, const int *blevel_ptr, int nblevels, int stagger, int nf0,
                                   int nf1, int *out4) {
    const mceik::fsm::TicketTask k = mceik::fsm::decode_ticket(t, vptr, blevel_ptr, nblevels, stagger, nf0, nf1);
    out4[0] = k.sweep; out4[1] = k.level; out4[2] = k.bidx; out4[3] = k.fidx;
}
