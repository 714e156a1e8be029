// gs.cu -- L2 travel-time-table grid search for sm_100a.
//
// Per event e and grid node g (reference: locate.c:923-1047, stack kernels :388-567):
//     t0[g]  = sum_i (w_i/sum w) * (tobs_i - T_i[g])                 w_i = 1/var_i
//     obj[g] = sum_i ( (w_i/sqrt2) * (tobs_i - (T_i[g] + t0[g])) )^2
//     iopt   = first index of the strict minimum of obj               (locate.c:811-830)
// accumulated over the used picks in catalogue order, with separate IEEE multiply and add (the
// reference build has no FMA), which is what makes the located index bit-exact.
//
// The reference streams every table twice per event and read-modify-writes t0/obj grids per
// pick.  Here nothing grid-sized is ever written: a CTA keeps t0 and obj of 8 events x 512
// nodes in registers, walks the picks once per pass with the table row loaded once for all 8
// events, reduces (obj, node, t0) with warp shuffles, and carries a running optimum across the
// chunks of the grid it owns.  fp32 tables are promoted to fp64 on load (locate.f90:414,459).
#include <algorithm>
#include <climits>
#include "gs.cuh"

namespace mceik {
namespace gs {

// literal of locate.c:496; note it rounds to 0x3FE6A09E667F3BCC, one ulp below M_SQRT1_2
#define MCEIK_SQRT2I 0.7071067811865475

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// ------------------------------------------------------------------------------------------
__global__ void prepare_kernel(int nevents, const int *__restrict__ obs_ptr, const int *__restrict__ table_id,
                               const double *__restrict__ varobs, double *__restrict__ w_t0,
                               double *__restrict__ w_obj, int *__restrict__ nuse) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nevents) return;
    const int beg = obs_ptr[e], end = obs_ptr[e + 1];
    double xnorm = 0.0;
    int n = 0;
    for (int p = beg; p < end; ++p)
        if (table_id[p] >= 0) {
            xnorm = __dadd_rn(xnorm, __ddiv_rn(1.0, varobs[p]));  // locate.c:993-994
            ++n;
        }
    for (int p = beg; p < end; ++p) {
        double a = 0.0, b = 0.0;
        if (table_id[p] >= 0) {
            const double wt = __ddiv_rn(1.0, varobs[p]);
            a = __ddiv_rn(wt, xnorm);         // locate.c:399
            b = __dmul_rn(wt, MCEIK_SQRT2I);  // locate.c:500
        }
        w_t0[p] = a;
        w_obj[p] = b;
    }
    nuse[e] = n;
}

void launch_prepare(int nevents, const int *d_obs_ptr, const int *d_table_id, const double *d_varobs,
                    double *d_w_t0, double *d_w_obj, int *d_nuse, cudaStream_t st) {
    if (nevents == 0) return;
    prepare_kernel<<<(nevents + 127) / 128, 128, 0, st>>>(nevents, d_obs_ptr, d_table_id, d_varobs, d_w_t0,
                                                          d_w_obj, d_nuse);
    MCEIK_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// (obj, node, t0) "less": smaller objective first, then smaller node index; NaN never wins.
__device__ __forceinline__ bool better(double v, int i, double bv, int bi) {
    return (v < bv) || (v == bv && i < bi);
}

template <int EB, int R>
__global__ void __launch_bounds__(kThreads, 2) locate_kernel(const LocateArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int P = a.maxpicks;
    double *s_tobs = reinterpret_cast<double *>(smem);  // [P][EB]
    double *s_w0 = s_tobs + (size_t)EB * P;
    double *s_w1 = s_w0 + (size_t)EB * P;
    int *s_id = reinterpret_cast<int *>(s_w1 + (size_t)EB * P);
    __shared__ double r_val[kThreads / 32][EB], r_t0[kThreads / 32][EB];
    __shared__ int r_idx[kThreads / 32][EB];
    __shared__ int s_nan0[EB];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int e0 = blockIdx.y * EB;
    if (a.blk_uniform && a.blk_uniform[blockIdx.y]) return;  // handled by locate_uniform_kernel

    for (int q = tid; q < EB * P; q += kThreads) {
        const int j = q / EB, e = q - j * EB, ev = e0 + e;
        int id = -1;
        double to = 0.0, w0 = 0.0, w1 = 0.0;
        if (ev < a.nevents) {
            const int beg = a.obs_ptr[ev];
            if (j < a.obs_ptr[ev + 1] - beg) {
                id = a.table_id[beg + j];
                to = a.tobs_cor[beg + j];
                w0 = a.w_t0[beg + j];
                w1 = a.w_obj[beg + j];
            }
        }
        s_id[q] = id; s_tobs[q] = to; s_w0[q] = w0; s_w1[q] = w1;
    }
    if (tid < EB) s_nan0[tid] = 0;
    __syncthreads();

    double tfix[EB];
#pragma unroll
    for (int e = 0; e < EB; ++e) tfix[e] = (a.job == 2 || e0 + e >= a.nevents) ? 0.0 : a.tori[e0 + e];

    // running optimum of event `tid` (threads 0..EB-1 only)
    double best_val = d_inf(), best_t0 = 0.0;
    int best_idx = INT_MAX;

    const int nchunks = (a.ngrd + kChunk - 1) / kChunk;
    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        int g[R], gl[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            g[r] = chunk * kChunk + r * kThreads + tid;
            gl[r] = min(g[r], a.ngrd - 1);
        }
        double t0[EB][R], obj[EB][R];
#pragma unroll
        for (int e = 0; e < EB; ++e)
#pragma unroll
            for (int r = 0; r < R; ++r) { t0[e][r] = tfix[e]; obj[e][r] = 0.0; }

#pragma unroll 1
        for (int pass = (a.job == 2 ? 0 : 1); pass < 2; ++pass) {
            const double *s_w = pass == 0 ? s_w0 : s_w1;
            // table row of the first used pick of slot j, loaded one slot ahead
            double Tn[R];
            int idn = -1;
#pragma unroll
            for (int e = EB - 1; e >= 0; --e) if (s_id[e] >= 0) idn = s_id[e];
            if (idn >= 0) {
                const float *row = a.tables + (size_t)idn * a.ldgrd;
#pragma unroll
                for (int r = 0; r < R; ++r) Tn[r] = (double)__ldg(row + gl[r]);
            }
#pragma unroll 1
            for (int j = 0; j < P; ++j) {
                double T[R];
                int cur = idn;
#pragma unroll
                for (int r = 0; r < R; ++r) T[r] = Tn[r];
                idn = -1;
                if (j + 1 < P) {
#pragma unroll
                    for (int e = EB - 1; e >= 0; --e) if (s_id[(j + 1) * EB + e] >= 0) idn = s_id[(j + 1) * EB + e];
                    if (idn >= 0) {
                        const float *row = a.tables + (size_t)idn * a.ldgrd;
#pragma unroll
                        for (int r = 0; r < R; ++r) Tn[r] = (double)__ldg(row + gl[r]);
                    }
                }
#pragma unroll
                for (int e = 0; e < EB; ++e) {
                    const int id = s_id[j * EB + e];
                    if (id >= 0) {  // uniform across the CTA
                        if (id != cur) {
                            cur = id;
                            const float *row = a.tables + (size_t)id * a.ldgrd;
#pragma unroll
                            for (int r = 0; r < R; ++r) T[r] = (double)__ldg(row + gl[r]);
                        }
                        const double to = s_tobs[j * EB + e], w = s_w[j * EB + e];
                        if (pass == 0) {
#pragma unroll
                            for (int r = 0; r < R; ++r)  // locate.c:409
                                t0[e][r] = __dadd_rn(t0[e][r], __dmul_rn(w, __dsub_rn(to, T[r])));
                        } else {
#pragma unroll
                            for (int r = 0; r < R; ++r) {  // locate.c:511-512
                                const double res = __dmul_rn(w, __dsub_rn(to, __dadd_rn(T[r], t0[e][r])));
                                obj[e][r] = __dadd_rn(obj[e][r], __dmul_rn(res, res));
                            }
                        }
                    }
                }
            }
        }

        // ---- optimum of this chunk per event
        if (chunk == 0 && tid == 0) {
#pragma unroll
            for (int e = 0; e < EB; ++e) if (obj[e][0] != obj[e][0]) s_nan0[e] = 1;
        }
#pragma unroll
        for (int e = 0; e < EB; ++e) {
            double bv = d_inf(), bt = 0.0;
            int bi = INT_MAX;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (g[r] < a.ngrd && better(obj[e][r], g[r], bv, bi)) { bv = obj[e][r]; bi = g[r]; bt = t0[e][r]; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, bv, off);
                const double ot = __shfl_down_sync(0xffffffffu, bt, off);
                const int oi = __shfl_down_sync(0xffffffffu, bi, off);
                if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; bt = ot; }
            }
            if (lane == 0) { r_val[warp][e] = bv; r_t0[warp][e] = bt; r_idx[warp][e] = bi; }
        }
        __syncthreads();
        if (tid < EB) {
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w)
                if (better(r_val[w][tid], r_idx[w][tid], best_val, best_idx)) {
                    best_val = r_val[w][tid]; best_idx = r_idx[w][tid]; best_t0 = r_t0[w][tid];
                }
        }
        __syncthreads();
    }

    if (tid < EB && e0 + tid < a.nevents) {
        Partial p;
        p.val = best_val; p.t0 = best_t0; p.idx = best_idx; p.nan0 = s_nan0[tid];
        a.partials[(size_t)(e0 + tid) * a.nlanes + blockIdx.x] = p;
    }
}


// ------------------------------------------------------------------------------------------
// Fast path: event blocks whose 8 events use the same table in every pick slot (the rectangular
// catalogues of locate3d_gridsearch / homog.c, where slot j is one (station, phase) for every
// event).  Unused picks enter with weight +0: acc + (+0 * d) == acc bit for bit for finite d, so
// the inner loop carries no branch, and the table value is converted once for all 8 events.  Table
// rows are fetched four slots ahead; the per-(slot, event) constants come from shared memory as
// one 16-byte broadcast.  ~85-90 % of the issued instructions are the 8 non-fused fp64 operations
// per (event, pick, node) that the reference arithmetic requires.
// ------------------------------------------------------------------------------------------
__global__ void classify_kernel(int nevents, int eb, const int *__restrict__ obs_ptr, const int *__restrict__ table_id,
                                int *__restrict__ blk_uniform) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nblk = (nevents + eb - 1) / eb;
    if (b >= nblk) return;
    const int e0 = b * eb, e1 = min(e0 + eb, nevents);
    int maxp = 0;
    for (int e = e0; e < e1; ++e) maxp = max(maxp, obs_ptr[e + 1] - obs_ptr[e]);
    bool ok = true;
    for (int j = lane; j < maxp; j += 32) {
        int id = -1;
        for (int e = e0; e < e1; ++e) {
            const int beg = obs_ptr[e];
            if (j < obs_ptr[e + 1] - beg) {
                const int t = table_id[beg + j];
                if (t >= 0) { if (id < 0) id = t; else if (id != t) ok = false; }
            }
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    // a block without any pick, or with more slots than the fast kernel's shared memory holds, takes the general kernel
    if (maxp == 0 || maxp > kUniformMaxPicks) ok = false;
    if (lane == 0) blk_uniform[b] = ok ? 1 : 0;
}

void launch_classify(int nevents, const int *d_obs_ptr, const int *d_table_id, int *d_blk_uniform, cudaStream_t st) {
    const int nblk = (nevents + kEventsPerBlock - 1) / kEventsPerBlock;
    if (nblk == 0) return;
    classify_kernel<<<(nblk * 32 + 127) / 128, 128, 0, st>>>(nevents, kEventsPerBlock, d_obs_ptr, d_table_id, d_blk_uniform);
    MCEIK_LAUNCH_CHECK();
}

template <int EB, int R>
__global__ void __launch_bounds__(kThreads, 2) locate_uniform_kernel(const LocateArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int e0 = blockIdx.y * EB;
    if (!a.blk_uniform[blockIdx.y]) return;  // mixed tables per slot: handled by locate_kernel
    // slots walked by this block: the longest pick list of its own events (shared memory is sized for a.maxpicks)
    int pblk = 0;
#pragma unroll
    for (int e = 0; e < EB; ++e)
        if (e0 + e < a.nevents) pblk = max(pblk, a.obs_ptr[e0 + e + 1] - a.obs_ptr[e0 + e]);
    const int P = (min(pblk, a.maxpicks) + 3) & ~3;
    double2 *s_a = reinterpret_cast<double2 *>(smem);  // [P][EB] (tobs, w_t0)
    double2 *s_b = s_a + (size_t)EB * P;               // [P][EB] (tobs, w_obj)
    int *s_id = reinterpret_cast<int *>(s_b + (size_t)EB * P);  // [P] table of the slot
    __shared__ double r_val[kThreads / 32][EB], r_t0[kThreads / 32][EB];
    __shared__ int r_idx[kThreads / 32][EB];
    __shared__ int s_nan0[EB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int j = tid; j < P; j += kThreads) s_id[j] = 0;
    __syncthreads();
    for (int q = tid; q < EB * P; q += kThreads) {
        const int j = q / EB, e = q - j * EB, ev = e0 + e;
        double to = 0.0, w0 = 0.0, w1 = 0.0;
        if (ev < a.nevents) {
            const int beg = a.obs_ptr[ev];
            if (j < a.obs_ptr[ev + 1] - beg) {
                const int id = a.table_id[beg + j];
                if (id >= 0) {
                    to = a.tobs_cor[beg + j]; w0 = a.w_t0[beg + j]; w1 = a.w_obj[beg + j];
                    s_id[j] = id;  // all used picks of the slot agree (uniform block)
                }
            }
        }
        s_a[q] = make_double2(to, w0);
        s_b[q] = make_double2(to, w1);
    }
    if (tid < EB) s_nan0[tid] = 0;
    __syncthreads();

    double tfix[EB];
#pragma unroll
    for (int e = 0; e < EB; ++e) tfix[e] = (a.job == 2 || e0 + e >= a.nevents) ? 0.0 : a.tori[e0 + e];
    double best_val = d_inf(), best_t0 = 0.0;
    int best_idx = INT_MAX;

    const int nchunks = (a.ngrd + kChunk - 1) / kChunk;
    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        int g[R], gl[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            g[r] = chunk * kChunk + r * kThreads + tid;
            gl[r] = min(g[r], a.ngrd - 1);
        }
        double t0[EB][R], obj[EB][R];
#pragma unroll
        for (int e = 0; e < EB; ++e)
#pragma unroll
            for (int r = 0; r < R; ++r) { t0[e][r] = tfix[e]; obj[e][r] = 0.0; }

#pragma unroll 1
        for (int pass = (a.job == 2 ? 0 : 1); pass < 2; ++pass) {
            const double2 *s_c = pass == 0 ? s_a : s_b;
            float Tn[4][R];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float *row = a.tables + (size_t)s_id[jj] * a.ldgrd;
#pragma unroll
                for (int r = 0; r < R; ++r) Tn[jj][r] = __ldg(row + gl[r]);
            }
#pragma unroll 1
            for (int j0 = 0; j0 < P; j0 += 4) {
                float Tc[4][R];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int r = 0; r < R; ++r) Tc[jj][r] = Tn[jj][r];
                if (j0 + 4 < P) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float *row = a.tables + (size_t)s_id[j0 + 4 + jj] * a.ldgrd;
#pragma unroll
                        for (int r = 0; r < R; ++r) Tn[jj][r] = __ldg(row + gl[r]);
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    double T[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) T[r] = (double)Tc[jj][r];  // DBLE(test4), locate.f90:414,459
                    // the (tobs, weight) constants of kCB events are fetched together and their kCB * R update
                    // chains are written stage by stage, so neither the shared-memory latency nor the fp64
                    // latency of one chain is exposed (each event/node chain keeps its own operation order)
                    const double2 *c = s_c + (size_t)(j0 + jj) * EB;
                    constexpr int kCB = 4;
                    static_assert(EB % kCB == 0, "event block must be a multiple of the constant batch");
                    if (pass == 0) {
#pragma unroll
                        for (int e0b = 0; e0b < EB; e0b += kCB) {
                            double2 cw[kCB];
                            double d[kCB][R];
#pragma unroll
                            for (int q = 0; q < kCB; ++q) cw[q] = c[e0b + q];
#pragma unroll
                            for (int q = 0; q < kCB; ++q)
#pragma unroll
                                for (int r = 0; r < R; ++r) d[q][r] = __dsub_rn(cw[q].x, T[r]);
#pragma unroll
                            for (int q = 0; q < kCB; ++q)
#pragma unroll
                                for (int r = 0; r < R; ++r) d[q][r] = __dmul_rn(cw[q].y, d[q][r]);
#pragma unroll
                            for (int q = 0; q < kCB; ++q)
#pragma unroll
                                for (int r = 0; r < R; ++r)  // locate.c:409
                                    t0[e0b + q][r] = __dadd_rn(t0[e0b + q][r], d[q][r]);
                        }
                    } else {
#pragma unroll
                        for (int e0b = 0; e0b < EB; e0b += kCB) {
                            double2 cw[kCB];
                            double d[kCB][R];
#pragma unroll
                            for (int q = 0; q < kCB; ++q) cw[q] = c[e0b + q];
#pragma unroll
                            for (int q = 0; q < kCB; ++q)
#pragma unroll
                                for (int r = 0; r < R; ++r) d[q][r] = __dadd_rn(T[r], t0[e0b + q][r]);
#pragma unroll
                            for (int q = 0; q < kCB; ++q)
#pragma unroll
                                for (int r = 0; r < R; ++r) d[q][r] = __dsub_rn(cw[q].x, d[q][r]);
#pragma unroll
                            for (int q = 0; q < kCB; ++q)
#pragma unroll
                                for (int r = 0; r < R; ++r) d[q][r] = __dmul_rn(cw[q].y, d[q][r]);  // locate.c:511
#pragma unroll
                            for (int q = 0; q < kCB; ++q)
#pragma unroll
                                for (int r = 0; r < R; ++r) d[q][r] = __dmul_rn(d[q][r], d[q][r]);
#pragma unroll
                            for (int q = 0; q < kCB; ++q)
#pragma unroll
                                for (int r = 0; r < R; ++r)  // locate.c:512
                                    obj[e0b + q][r] = __dadd_rn(obj[e0b + q][r], d[q][r]);
                        }
                    }
                }
            }
        }

        if (chunk == 0 && tid == 0) {
#pragma unroll
            for (int e = 0; e < EB; ++e) if (obj[e][0] != obj[e][0]) s_nan0[e] = 1;
        }
#pragma unroll
        for (int e = 0; e < EB; ++e) {
            double bv = d_inf(), bt = 0.0;
            int bi = INT_MAX;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (g[r] < a.ngrd && better(obj[e][r], g[r], bv, bi)) { bv = obj[e][r]; bi = g[r]; bt = t0[e][r]; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, bv, off);
                const double ot = __shfl_down_sync(0xffffffffu, bt, off);
                const int oi = __shfl_down_sync(0xffffffffu, bi, off);
                if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; bt = ot; }
            }
            if (lane == 0) { r_val[warp][e] = bv; r_t0[warp][e] = bt; r_idx[warp][e] = bi; }
        }
        __syncthreads();
        if (tid < EB) {
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w)
                if (better(r_val[w][tid], r_idx[w][tid], best_val, best_idx)) {
                    best_val = r_val[w][tid]; best_idx = r_idx[w][tid]; best_t0 = r_t0[w][tid];
                }
        }
        __syncthreads();
    }
    if (tid < EB && e0 + tid < a.nevents) {
        Partial p;
        p.val = best_val; p.t0 = best_t0; p.idx = best_idx; p.nan0 = s_nan0[tid];
        a.partials[(size_t)(e0 + tid) * a.nlanes + blockIdx.x] = p;
    }
}

size_t locate_uniform_smem_bytes(int maxpicks) {
    const size_t P = (size_t)((std::max(maxpicks, 1) + 3) & ~3);
    return P * kEventsPerBlock * 2 * sizeof(double2) + P * sizeof(int);
}

size_t locate_smem_bytes(int maxpicks) {
    return (size_t)kEventsPerBlock * (size_t)std::max(maxpicks, 1) * (3 * sizeof(double) + sizeof(int));
}

// Grid split of the search: every block of 8 events is searched by `lanes` CTAs that share out the grid chunks.  All
// CTAs do the same amount of work and 2 are resident per SM, so the launch runs in waves of 296 CTAs: the lane count is
// the one (up to 128) whose last wave is fullest -- 256 events x 19 lanes = 608 CTAs used to leave a third wave 4 %
// full (351 events/s against 425 at 64 events x 74 lanes = exactly two waves).
int locate_lanes(int nevents, int ngrd) {
    const int nblocks = std::max(1, (nevents + kEventsPerBlock - 1) / kEventsPerBlock);
    const int nchunks = (ngrd + kChunk - 1) / kChunk;
    const int wave = 148 * 2;
    int best = 1;
    double best_eff = -1.0;
    for (int lanes = 1; lanes <= std::min(nchunks, 128); ++lanes) {
        const long long ctas = (long long)nblocks * lanes;
        const long long waves = (ctas + wave - 1) / wave;
        double eff = (double)ctas / (double)(waves * wave);
        if (nchunks % lanes) eff *= (double)(nchunks / lanes) / (double)(nchunks / lanes + 1);  // uneven chunk counts
        if (eff > best_eff + 1e-9) { best_eff = eff; best = lanes; }
    }
    return best;
}

void launch_locate(const LocateArgs &a, cudaStream_t st) {
    if (a.nevents == 0) return;
    const size_t smem = locate_smem_bytes(a.maxpicks);
    if (smem > 200 * 1024) throw CudaError("too many picks per event for the locate kernel (limit ~900)");
    auto kern = locate_kernel<kEventsPerBlock, kPointsPerThread>;
    MCEIK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nblocks = (a.nevents + kEventsPerBlock - 1) / kEventsPerBlock;
    if (nblocks > 65535) throw CudaError("too many events per call (limit 524280)");
    dim3 grid(a.nlanes, nblocks);
    kern<<<grid, kThreads, smem, st>>>(a);
    MCEIK_LAUNCH_CHECK();
    if (a.blk_uniform) {  // blocks classify_kernel marked uniform (at most kUniformMaxPicks slots each)
        const size_t smem_u = locate_uniform_smem_bytes(std::min(a.maxpicks, kUniformMaxPicks));
        auto ku = locate_uniform_kernel<kEventsPerBlock, kPointsPerThread>;
        MCEIK_CUDA(cudaFuncSetAttribute(ku, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u));
        ku<<<grid, kThreads, smem_u, st>>>(a);
        MCEIK_LAUNCH_CHECK();
    }
}

// one warp per event merges the lane partials
__global__ void finalize_kernel(int nevents, int nlanes, const Partial *__restrict__ partials,
                                const int *__restrict__ nuse, int *__restrict__ iopt, double *__restrict__ t0opt,
                                double *__restrict__ objopt) {
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e >= nevents) return;
    double bv = d_inf(), bt = 0.0;
    int bi = INT_MAX, nan0 = 0;
    for (int l = lane; l < nlanes; l += 32) {
        const Partial p = partials[(size_t)e * nlanes + l];
        nan0 |= p.nan0;
        if (better(p.val, p.idx, bv, bi)) { bv = p.val; bi = p.idx; bt = p.t0; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, off);
        const double ot = __shfl_down_sync(0xffffffffu, bt, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        nan0 |= __shfl_down_sync(0xffffffffu, nan0, off);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; bt = ot; }
    }
    if (lane == 0) {
        if (nuse[e] == 0) {           // no usable pick: flagged, nothing located
            iopt[e] = -1; t0opt[e] = 0.0; objopt[e] = 0.0;
        } else if (nan0 || bi == INT_MAX) {  // NaN at node 0 is sticky in locate.c:821-828 -> index 0
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            iopt[e] = 0; t0opt[e] = qnan; objopt[e] = qnan;
        } else {
            iopt[e] = bi; t0opt[e] = bt; objopt[e] = bv;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Catalogue misfit of many proposals (BASELINE config 5; no reference code -- the definition is SURVEY.md 8d C5):
// for model m and event e at its catalogue node, with the model's own tables T_j = tables[m * ntab + j][node_e],
//   w_j = 1 / var_j,  t0 = sum_j (w_j / sum w) (tobs_j - T_j),  obj_e = sum_j (w_j * sqrt(1/2) * (tobs_j - (T_j + t0)))^2
// over the used picks in pick order (the arithmetic of locate.c:399-410, 500-513), misfit_m = sum_e obj_e.
// One block per model, one thread per event (strided), then a fixed-order tree: bit-reproducible run to run.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) catalog_misfit_kernel(const float *__restrict__ tables, size_t ldgrd, int ntab, int nevents,
                                                            const int *__restrict__ node, const double *__restrict__ tobs,
                                                            const double *__restrict__ var, const int *__restrict__ use,
                                                            double *__restrict__ out) {
    const int m = blockIdx.x;
    const float *tm = tables + (size_t)m * ntab * ldgrd;
    __shared__ double part[256];
    double acc = 0.0;
    for (int e = threadIdx.x; e < nevents; e += blockDim.x) {
        const int g = node[e];
        const double *to = tobs + (size_t)e * ntab, *va = var + (size_t)e * ntab;
        const int *us = use + (size_t)e * ntab;
        double xnorm = 0.0;
        for (int j = 0; j < ntab; ++j)
            if (us[j]) xnorm = __dadd_rn(xnorm, __ddiv_rn(1.0, va[j]));
        double t0 = 0.0;
        for (int j = 0; j < ntab; ++j)
            if (us[j]) {
                const double w = __ddiv_rn(__ddiv_rn(1.0, va[j]), xnorm);
                t0 = __dadd_rn(t0, __dmul_rn(w, __dsub_rn(to[j], (double)tm[(size_t)j * ldgrd + g])));
            }
        double obj = 0.0;
        for (int j = 0; j < ntab; ++j)
            if (us[j]) {
                const double w = __dmul_rn(__ddiv_rn(1.0, va[j]), 0.7071067811865475);
                const double r = __dmul_rn(w, __dsub_rn(to[j], __dadd_rn((double)tm[(size_t)j * ldgrd + g], t0)));
                obj = __dadd_rn(obj, __dmul_rn(r, r));
            }
        if (xnorm > 0.0) acc = __dadd_rn(acc, obj);  // an event without a usable pick contributes nothing
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) part[threadIdx.x] = __dadd_rn(part[threadIdx.x], part[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[m] = part[0];
}

void launch_catalog_misfit(const float *d_tables, size_t ldgrd, int nmodels, int ntab, int nevents, const int *d_node,
                           const double *d_tobs, const double *d_var, const int *d_use, double *d_out, cudaStream_t st) {
    if (nmodels == 0) return;
    catalog_misfit_kernel<<<nmodels, 256, 0, st>>>(d_tables, ldgrd, ntab, nevents, d_node, d_tobs, d_var, d_use, d_out);
    MCEIK_LAUNCH_CHECK();
}

void launch_finalize(int nevents, int nlanes, const Partial *d_partials, const int *d_nuse, int *d_iopt,
                     double *d_t0opt, double *d_objopt, cudaStream_t st) {
    if (nevents == 0) return;
    const int threads = 128;
    const int blocks = (int)(((size_t)nevents * 32 + threads - 1) / threads);
    finalize_kernel<<<blocks, threads, 0, st>>>(nevents, nlanes, d_partials, d_nuse, d_iopt, d_t0opt, d_objopt);
    MCEIK_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// Single event, full-grid outputs (the legacy per-call contract).
// ------------------------------------------------------------------------------------------
template <typename T> struct Ar;
template <> struct Ar<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
};
template <> struct Ar<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
};

template <typename T>
__global__ void full_grid_kernel(int ngrd, size_t ldgrd, int nuse, const int *__restrict__ row, const T *__restrict__ tobs,
                                 const T *__restrict__ w_t0, const T *__restrict__ w_obj, int want_ot, T t0use,
                                 const T *__restrict__ test, T *__restrict__ t0out, T *__restrict__ objout) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngrd) return;
    T t0 = want_ot ? (T)0 : t0use;
    if (want_ot)
        for (int j = 0; j < nuse; ++j)  // locate.c:407-410 / gridsearch.f90:189-191
            t0 = Ar<T>::add(t0, Ar<T>::mul(w_t0[j], Ar<T>::sub(tobs[j], __ldg(test + (size_t)row[j] * ldgrd + g))));
    T obj = (T)0;
    for (int j = 0; j < nuse; ++j) {    // locate.c:509-513 / gridsearch.f90:277-280
        const T res = Ar<T>::mul(w_obj[j], Ar<T>::sub(tobs[j], Ar<T>::add(__ldg(test + (size_t)row[j] * ldgrd + g), t0)));
        obj = Ar<T>::add(obj, Ar<T>::mul(res, res));
    }
    t0out[g] = t0;
    objout[g] = obj;
}

template <typename T>
void launch_full_grid(int ngrd, size_t ldgrd, int nuse, const int *d_row, const T *d_tobs, const T *d_w_t0,
                      const T *d_w_obj, int want_ot, T t0use, const T *d_test, T *d_t0, T *d_obj, cudaStream_t st) {
    if (ngrd == 0) return;
    full_grid_kernel<T><<<(ngrd + 255) / 256, 256, 0, st>>>(ngrd, ldgrd, nuse, d_row, d_tobs, d_w_t0, d_w_obj, want_ot,
                                                            t0use, d_test, d_t0, d_obj);
    MCEIK_LAUNCH_CHECK();
}
template void launch_full_grid<double>(int, size_t, int, const int *, const double *, const double *, const double *,
                                       int, double, const double *, double *, double *, cudaStream_t);
template void launch_full_grid<float>(int, size_t, int, const int *, const float *, const float *, const float *, int,
                                      float, const float *, float *, float *, cudaStream_t);

// ------------------------------------------------------------------------------------------
// Per-event posterior volume from the resident fp32 tables (the logPDF the reference accumulates
// per event, locate.f90:436-463: logPDF = -sum res^2, fp32 tables promoted to fp64), plus the
// origin-time grid; fp32 copy = what the location file stores (h5io.c:714-819).
// ------------------------------------------------------------------------------------------
__global__ void event_grid_kernel(int ngrd, size_t ldgrd, int npicks, const int *__restrict__ table_id,
                                  const double *__restrict__ tobs, const double *__restrict__ w_t0,
                                  const double *__restrict__ w_obj, int want_ot, double t0use,
                                  const float *__restrict__ tables, double *__restrict__ logpdf,
                                  float *__restrict__ logpdf4, double *__restrict__ t0out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngrd) return;
    double t0 = want_ot ? 0.0 : t0use;
    if (want_ot)
        for (int j = 0; j < npicks; ++j)
            if (table_id[j] >= 0)
                t0 = __dadd_rn(t0, __dmul_rn(w_t0[j], __dsub_rn(tobs[j], (double)__ldg(tables + (size_t)table_id[j] * ldgrd + g))));
    double obj = 0.0;
    for (int j = 0; j < npicks; ++j)
        if (table_id[j] >= 0) {
            const double res = __dmul_rn(w_obj[j], __dsub_rn(tobs[j], __dadd_rn((double)__ldg(tables + (size_t)table_id[j] * ldgrd + g), t0)));
            obj = __dadd_rn(obj, __dmul_rn(res, res));
        }
    if (logpdf) logpdf[g] = -obj;
    if (logpdf4) logpdf4[g] = __double2float_rn(-obj);
    if (t0out) t0out[g] = t0;
}

void launch_event_grid(int ngrd, size_t ldgrd, int npicks, const int *d_table_id, const double *d_tobs, const double *d_w_t0,
                       const double *d_w_obj, int want_ot, double t0use, const float *d_tables, double *d_logpdf,
                       float *d_logpdf4, double *d_t0, cudaStream_t st) {
    if (ngrd == 0) return;
    event_grid_kernel<<<(ngrd + 255) / 256, 256, 0, st>>>(ngrd, ldgrd, npicks, d_table_id, d_tobs, d_w_t0, d_w_obj, want_ot,
                                                          t0use, d_tables, d_logpdf, d_logpdf4, d_t0);
    MCEIK_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// L1 flavour, one event (locate_l1_gridSearch__double64, locate.c:1205-1335): origin time = weighted
// median of the residuals tobs_i - T_i[g] (weights 1/var, normalised), misfit = sum w_i |res_i - t0|.
// The reference declares its weighted median (locate.c:73) but defines it nowhere; the definition
// used here is in include/mceik_b200.h.  One thread per grid node: stable insertion sort of the
// (residual, weight) pairs in local memory, then one pass over the cumulative weights.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double weighted_median_sorted(int n, const double *x, const double *w) {
    double W = 0.0;
    for (int k = 0; k < n; ++k) W = __dadd_rn(W, w[k]);
    const double half = __dmul_rn(0.5, W);
    double cum = 0.0;
    for (int k = 0; k < n; ++k) {
        cum = __dadd_rn(cum, w[k]);
        if (cum > half) return x[k];
        if (cum == half) return k + 1 < n ? __dmul_rn(0.5, __dadd_rn(x[k], x[k + 1])) : x[k];
    }
    return n > 0 ? x[n - 1] : 0.0;
}

__global__ void l1_grid_kernel(int ngrd, size_t ldgrd, int nuse, const double *__restrict__ tobs,
                               const double *__restrict__ wt, int want_ot, double t0use,
                               const double *__restrict__ test, double *__restrict__ t0out, double *__restrict__ obj) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngrd) return;
    double t0 = t0use;
    if (want_ot) {
        double x[kL1MaxObs], w[kL1MaxObs];
        for (int i = 0; i < nuse; ++i) {  // picks arrive in catalogue order; equal residuals keep that order
            const double r = __dsub_rn(tobs[i], test[(size_t)i * ldgrd + g]);
            int k = i;
            while (k > 0 && x[k - 1] > r) { x[k] = x[k - 1]; w[k] = w[k - 1]; --k; }
            x[k] = r; w[k] = wt[i];
        }
        t0 = weighted_median_sorted(nuse, x, w);
    }
    double acc = 0.0;
    for (int i = 0; i < nuse; ++i)  // locate.c:1320-1321
        acc = __dadd_rn(acc, __dmul_rn(wt[i], fabs(__dsub_rn(__dsub_rn(tobs[i], test[(size_t)i * ldgrd + g]), t0))));
    t0out[g] = t0;
    obj[g] = acc;
}

void launch_l1_grid(int ngrd, size_t ldgrd, int nuse, const double *d_tobs, const double *d_wt, int want_ot, double t0use,
                    const double *d_test, double *d_t0, double *d_obj, cudaStream_t st) {
    if (ngrd == 0) return;
    if (nuse > kL1MaxObs) throw CudaError("L1 grid search: more than kL1MaxObs used observations");
    l1_grid_kernel<<<(ngrd + 127) / 128, 128, 0, st>>>(ngrd, ldgrd, nuse, d_tobs, d_wt, want_ot, t0use, d_test, d_t0, d_obj);
    MCEIK_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// locate_minLoc{Double64,Float64} (locate.c:811-851)
// ------------------------------------------------------------------------------------------
constexpr int kMinlocBlocks = 148 * 4;
struct MinPart { double v; int i; int pad; };
size_t minloc_scratch_bytes() { return sizeof(MinPart) * kMinlocBlocks; }

template <typename T>
__global__ void minloc_stage1(int n, const T *__restrict__ x, MinPart *__restrict__ parts, double sign) {
    double bv = d_inf();
    int bi = INT_MAX;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double v = sign * (double)x[i];  // float -> double and the sign flip (MAXLOC) are exact, order preserving
        if (better(v, i, bv, bi)) { bv = v; bi = i; }
    }
    __shared__ double sv[8];
    __shared__ int si[8];
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) if (better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
        parts[blockIdx.x].v = bv; parts[blockIdx.x].i = bi;
    }
}
template <typename T>
__global__ void minloc_stage2(int nparts, const MinPart *__restrict__ parts, const T *__restrict__ x, int *__restrict__ out) {
    double bv = d_inf();
    int bi = INT_MAX;
    for (int p = threadIdx.x; p < nparts; p += 32)
        if (better(parts[p].v, parts[p].i, bv, bi)) { bv = parts[p].v; bi = parts[p].i; }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (threadIdx.x == 0) {
        const T x0 = x[0];
        out[0] = (x0 != x0 || bi == INT_MAX) ? 0 : bi;  // a NaN at x[0] is sticky in the reference scan
    }
}

template <typename T>
void launch_minloc(int n, const T *d_x, int *d_out, void *d_scratch, size_t scratch_bytes, cudaStream_t st, bool maxloc) {
    if (scratch_bytes < minloc_scratch_bytes()) throw CudaError("minloc scratch too small");
    MinPart *parts = reinterpret_cast<MinPart *>(d_scratch);
    const int blocks = std::max(1, std::min(kMinlocBlocks, (n + 255) / 256));
    minloc_stage1<T><<<blocks, 256, 0, st>>>(n, d_x, parts, maxloc ? -1.0 : 1.0);
    MCEIK_LAUNCH_CHECK();
    minloc_stage2<T><<<1, 32, 0, st>>>(blocks, parts, d_x, d_out);
    MCEIK_LAUNCH_CHECK();
}
template void launch_minloc<double>(int, const double *, int *, void *, size_t, cudaStream_t, bool);
template void launch_minloc<float>(int, const float *, int *, void *, size_t, cudaStream_t, bool);

// ------------------------------------------------------------------------------------------
// LOCATE_NORMALIZE_PDF (locate.f90:43-64): sum of the grid (fixed reduction tree -> the same bits
// on every run; the order differs from the Fortran SUM, see the tolerance in the header) and DSCAL.
// ------------------------------------------------------------------------------------------
__global__ void sum_stage1(int n, const double *__restrict__ x, double *__restrict__ parts) {
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += x[i];
    __shared__ double sv[8];
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) acc += sv[w];
        parts[blockIdx.x] = acc;
    }
}
__global__ void sum_stage2(int nparts, const double *__restrict__ parts, double *__restrict__ out) {
    double acc = 0.0;
    for (int p = threadIdx.x; p < nparts; p += 32) acc += parts[p];
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if (threadIdx.x == 0) out[0] = acc;
}
__global__ void scale_kernel(int n, double factor, double *__restrict__ x) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = __dmul_rn(x[i], factor);
}
void launch_sum(int n, const double *d_x, double *d_out, void *d_scratch, size_t scratch_bytes, cudaStream_t st) {
    if (scratch_bytes < sizeof(double) * kMinlocBlocks) throw CudaError("sum scratch too small");
    const int blocks = std::max(1, std::min(kMinlocBlocks, (n + 255) / 256));
    sum_stage1<<<blocks, 256, 0, st>>>(n, d_x, reinterpret_cast<double *>(d_scratch));
    MCEIK_LAUNCH_CHECK();
    sum_stage2<<<1, 32, 0, st>>>(blocks, reinterpret_cast<const double *>(d_scratch), d_out);
    MCEIK_LAUNCH_CHECK();
}
void launch_scale(int n, double factor, double *d_x, cudaStream_t st) {
    if (n == 0) return;
    scale_kernel<<<std::min(kMinlocBlocks, (n + 255) / 256), 256, 0, st>>>(n, factor, d_x);
    MCEIK_LAUNCH_CHECK();
}

}  // namespace gs
}  // namespace mceik
