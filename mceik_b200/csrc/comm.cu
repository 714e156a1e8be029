// comm.cu -- the multi-GPU part of the C ABI: one process per GPU, sources sharded over the ranks, tables replicated.
//
// The reference splits ONE field over MPI ranks and exchanges ghost layers every sweep (fsm3d.f90:971-1045), then
// gathers through rank 0 (fsm3d.f90:1488-1553); its callers hand in a communicator and nothing else
// (mpiutils.f90:99-264).  Here a field never leaves its GPU: the ranks take whole fields, solve them without any
// communication and write their fp32 tables straight into their rows of the replicated table buffer.  Replication is
// one-sided when the buffer is the library's own (alloc_replicated: CUDA IPC mappings of every peer's buffer): a table
// is put into the peers' copies by the copy engines over NVLink as soon as its field has converged, under the sweeps
// of the remaining fields.  A caller-owned buffer is completed by one in-place ncclAllGather after the solve.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 -- the copy a host such as PyTorch already loaded, else the
// system one), so libmceik_b200.so has no link-time dependency on it and single-GPU hosts need no NCCL at all.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <numeric>
#include <vector>

#include "../../include/mceik_b200.h"
#include "comm.cuh"
#include "common.cuh"

namespace mceik {
namespace comm {

namespace {
// the slice of nccl.h this file needs (ABI stable over NCCL 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt32 = 2, ncclFloat32 = 7 };
struct Api {
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
Api g_api;
std::once_flag g_once;

void load_api() {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    auto sym = [&](const char *n) { return dlsym(h, n); };
    g_api.GetUniqueId = reinterpret_cast<decltype(g_api.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_api.CommInitRank = reinterpret_cast<decltype(g_api.CommInitRank)>(sym("ncclCommInitRank"));
    g_api.CommDestroy = reinterpret_cast<decltype(g_api.CommDestroy)>(sym("ncclCommDestroy"));
    g_api.AllGather = reinterpret_cast<decltype(g_api.AllGather)>(sym("ncclAllGather"));
    g_api.GetErrorString = reinterpret_cast<decltype(g_api.GetErrorString)>(sym("ncclGetErrorString"));
    g_api.ok = g_api.GetUniqueId && g_api.CommInitRank && g_api.CommDestroy && g_api.AllGather && g_api.GetErrorString;
}

const Api &api() {
    std::call_once(g_once, load_api);
    if (!g_api.ok) throw CudaError("NCCL (libnccl.so.2) could not be loaded: multi-GPU entry points are unavailable");
    return g_api;
}

void check(int rc, const char *what) {
    if (rc != ncclSuccess) throw CudaError(std::string(what) + " failed: " + api().GetErrorString(rc));
}
}  // namespace

struct Comm {
    ncclComm_t nccl = nullptr;
    int world = 1, rank = 0;
    // replicated table buffer (alloc_replicated): this rank's allocation and the peers' mappings of theirs
    void *rep_local = nullptr;
    size_t rep_bytes = 0;
    std::vector<void *> rep_peer;  // [world]; rep_peer[rank] == rep_local
};

void unique_id(void *out128) {
    ncclUniqueId id;
    check(api().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out128, &id, sizeof(id));
}

Comm *create(int world, int rank, const void *id128) {
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    Comm *c = new Comm();
    c->world = world;
    c->rank = rank;
    const int rc = api().CommInitRank(&c->nccl, world, id, rank);
    if (rc != ncclSuccess) {
        delete c;
        check(rc, "ncclCommInitRank");
    }
    return c;
}

void free_replicated(Comm *c) {
    if (!c || !c->rep_local) return;
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->rep_peer[r]) cudaIpcCloseMemHandle(c->rep_peer[r]);
    cudaFree(c->rep_local);
    c->rep_local = nullptr;
    c->rep_bytes = 0;
    c->rep_peer.clear();
}

void destroy(Comm *c) {
    if (!c) return;
    free_replicated(c);
    if (c->nccl) api().CommDestroy(c->nccl);
    delete c;
}

// One buffer of `bytes` per rank, every rank mapping every other rank's buffer (CUDA IPC; the 64-byte handles travel
// through an all-gather of the communicator): finished tables are then PUT into the peers' buffers by the copy engines
// over NVLink (cudaMemcpyAsync on a side stream) while the sweeps of the remaining fields keep the SMs -- no collective
// kernel, no rank waits for another during the solve.  Collective.
void *alloc_replicated(Comm *c, size_t bytes, cudaStream_t st) {
    if (!c) throw CudaError("alloc_replicated: the context has no communicator (mceik_comm_init)");
    free_replicated(c);
    MCEIK_CUDA(cudaMalloc(&c->rep_local, bytes));
    c->rep_bytes = bytes;
    c->rep_peer.assign(c->world, nullptr);
    c->rep_peer[c->rank] = c->rep_local;
    if (c->world == 1) return c->rep_local;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::vector<cudaIpcMemHandle_t> h(c->world);
    MCEIK_CUDA(cudaIpcGetMemHandle(&h[c->rank], c->rep_local));
    cudaIpcMemHandle_t *d_h = nullptr;
    MCEIK_CUDA(cudaMalloc(&d_h, sizeof(cudaIpcMemHandle_t) * c->world));
    MCEIK_CUDA(cudaMemcpyAsync(d_h + c->rank, &h[c->rank], sizeof(cudaIpcMemHandle_t), cudaMemcpyHostToDevice, st));
    all_gather_inplace(c, d_h, sizeof(cudaIpcMemHandle_t), st);
    MCEIK_CUDA(cudaMemcpyAsync(h.data(), d_h, sizeof(cudaIpcMemHandle_t) * c->world, cudaMemcpyDeviceToHost, st));
    MCEIK_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_h);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        MCEIK_CUDA(cudaIpcOpenMemHandle(&c->rep_peer[r], h[r], cudaIpcMemLazyEnablePeerAccess));
    }
    return c->rep_local;
}

void *replicated_local(const Comm *c) { return c ? c->rep_local : nullptr; }
size_t replicated_bytes(const Comm *c) { return c ? c->rep_bytes : 0; }

// rows [offset, offset + bytes) of this rank's buffer -> the same place in every peer's buffer (copy engines, `st`)
void put_to_peers(Comm *c, size_t offset, size_t bytes, cudaStream_t st) {
    if (!c || c->world == 1 || !c->rep_local) return;
    if (offset + bytes > c->rep_bytes) throw CudaError("put_to_peers: outside the replicated buffer");
    for (int k = 1; k < c->world; ++k) {  // start with the next rank: the puts of all ranks spread over all links
        const int r = (c->rank + k) % c->world;
        MCEIK_CUDA(cudaMemcpyAsync(static_cast<char *>(c->rep_peer[r]) + offset, static_cast<char *>(c->rep_local) + offset, bytes,
                                   cudaMemcpyDeviceToDevice, st));
    }
}

int world(const Comm *c) { return c ? c->world : 1; }
int rank(const Comm *c) { return c ? c->rank : 0; }

void all_gather_inplace(Comm *c, void *d_all, size_t bytes_per_rank, cudaStream_t st) {
    if (!c || c->world == 1 || bytes_per_rank == 0) return;
    if (bytes_per_rank % 4 != 0) throw CudaError("all_gather_inplace: slice size must be a multiple of 4 bytes");
    char *base = static_cast<char *>(d_all);
    check(api().AllGather(base + (size_t)c->rank * bytes_per_rank, base, bytes_per_rank / 4, ncclInt32, c->nccl, st), "ncclAllGather");
}

// Which rank solves which field.  Fields of one slowness model are dealt round-robin over the ranks that hold that
// model (with at least as many ranks as models every rank holds exactly one model, so its slowness stays hot in the
// L2); `cost` (may be NULL: all equal) is an estimate of the work per field -- the iteration counts of an earlier
// solve on nearly the same models, the common case inside an MCMC loop -- and turns the deal into a longest-
// processing-time-first assignment.  Every rank gets the same number of slots (ceil), so the table rows are
// rank-major: row = rank * slots + position.  Deterministic: every rank computes the same answer.
void assign_fields(int nfields, const int *field_model, const int *cost, int world, std::vector<int> &rank_of, std::vector<int> &row_of,
                   int &slots) {
    rank_of.assign(nfields, 0);
    row_of.assign(nfields, 0);
    slots = (nfields + world - 1) / world;
    std::vector<int> models(field_model, field_model + nfields);
    std::sort(models.begin(), models.end());
    models.erase(std::unique(models.begin(), models.end()), models.end());
    const int nm = (int)models.size();
    // ranks of model m: a contiguous group, sized in proportion to the model's share of the fields
    std::vector<int> count(nm, 0);
    for (int f = 0; f < nfields; ++f) count[std::lower_bound(models.begin(), models.end(), field_model[f]) - models.begin()]++;
    std::vector<int> g0(nm + 1, 0);
    if (nm <= world) {
        int used = 0, seen = 0;
        for (int m = 0; m < nm; ++m) {
            seen += count[m];
            int upto = (int)(((long long)seen * world + nfields - 1) / nfields);  // cumulative share, rounded up
            upto = std::max(upto, used + 1);
            upto = std::min(upto, world - (nm - 1 - m));
            g0[m + 1] = upto;
            used = upto;
        }
        g0[nm] = world;
    }
    std::vector<long long> load(world, 0);
    std::vector<int> fill(world, 0);
    std::vector<int> order(nfields);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (cost ? cost[a] : 1) > (cost ? cost[b] : 1); });
    for (int f : order) {
        const int m = (int)(std::lower_bound(models.begin(), models.end(), field_model[f]) - models.begin());
        int lo = 0, hi = world;
        if (nm <= world) { lo = g0[m]; hi = g0[m + 1]; }
        int best = -1;
        for (int pass = 0; pass < 2 && best < 0; ++pass) {  // second pass: the model's group is full, use any rank
            for (int r = (pass == 0 ? lo : 0); r < (pass == 0 ? hi : world); ++r)
                if (fill[r] < slots && (best < 0 || load[r] < load[best] || (load[r] == load[best] && fill[r] < fill[best]))) best = r;
        }
        rank_of[f] = best;
        load[best] += cost ? cost[f] : 1;
        fill[best]++;
    }
    // rows: a rank's fields in increasing field order
    std::fill(fill.begin(), fill.end(), 0);
    for (int f = 0; f < nfields; ++f) row_of[f] = rank_of[f] * slots + fill[rank_of[f]]++;
}

}  // namespace comm
}  // namespace mceik
