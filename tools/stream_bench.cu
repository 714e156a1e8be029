// stream_bench.cu -- how fast can one B200 stream brick columns the way the sweep kernel does (read u and slowness,
// write u; one warp per 64-node cross-section marching along z), as a function of the size of the contiguous pieces?
//   layout 0: natural [z][y][x], brick 8 x 8   -> 64-byte pieces at a 2 KB stride
//   layout 1: natural, brick 16 x 4            -> 128-byte pieces
//   layout 2: natural, brick 32 x 2            -> 256-byte pieces
//   layout 3: blocked [brick][z][8][8]         -> 512 contiguous bytes per brick plane
// Development aid for profiles/kernel_evolution_r2.md (memory-side ceiling of the brick walk).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/stream_bench.cu -o build/stream_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kDepth = 4;  // planes in flight per warp

__device__ __forceinline__ double2 ldcg(const double2 *p) { return __ldcg(p); }

template <int kLayout>
__global__ void __launch_bounds__(384, 1) walk(const double *__restrict__ u_in, const double *__restrict__ slow, double *__restrict__ u_out,
                                                int n, int nfields, unsigned long long *queue) {
    const int lane = threadIdx.x & 31;
    constexpr int bx = kLayout == 1 ? 16 : (kLayout == 2 ? 32 : 8), by = 64 / bx;
    const int nbx = n / bx, nby = n / by, nbricks = nbx * nby;
    const size_t N = (size_t)n * n * n, nxy = (size_t)n * n;
    const long long ntasks = (long long)nbricks * nfields;
    // lane -> 16-byte pair p of row r inside the cross-section
    const int ppr = bx / 2, r = lane / ppr, p = lane % ppr;
    while (true) {
        long long t = 0;
        if (lane == 0) t = (long long)atomicAdd(queue, 1ULL);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntasks) break;
        const int f = (int)(t % nfields), b = (int)(t / nfields);
        const int I = b % nbx, J = b / nbx;
        size_t base, zstride;
        if (kLayout == 3) { base = (size_t)b * n * 64 + (size_t)lane * 2; zstride = 64; }
        else { base = (size_t)(J * by + r) * n + (size_t)I * bx + 2 * p; zstride = nxy; }
        const double2 *pu = reinterpret_cast<const double2 *>(u_in + (size_t)f * N + base);
        const double2 *ps = reinterpret_cast<const double2 *>(slow + (size_t)(f & 1) * N + base);
        double2 *po = reinterpret_cast<double2 *>(u_out + (size_t)f * N + base);
        const size_t zs2 = zstride / 2;
        double2 a[kDepth], s[kDepth];
#pragma unroll
        for (int d = 0; d < kDepth; ++d) { a[d] = ldcg(pu + d * zs2); s[d] = ldcg(ps + d * zs2); }
        for (int k = 0; k < n; k += kDepth) {
            double2 o[kDepth];
#pragma unroll
            for (int d = 0; d < kDepth; ++d) o[d] = make_double2(a[d].x + s[d].x, a[d].y + s[d].y);
            if (k + kDepth < n) {
#pragma unroll
                for (int d = 0; d < kDepth; ++d) { a[d] = ldcg(pu + (k + kDepth + d) * zs2); s[d] = ldcg(ps + (k + kDepth + d) * zs2); }
            }
#pragma unroll
            for (int d = 0; d < kDepth; ++d) __stcg(po + (k + d) * zs2, o[d]);
        }
    }
}

template <int kLayout>
void run(const double *u, const double *sl, double *o, int n, int nf, unsigned long long *q, const char *name) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(q, 0, 8);
        cudaEventRecord(e0);
        walk<kLayout><<<148, 384>>>(u, sl, o, n, nf, q);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double bytes = 24.0 * (double)n * n * n * nf;
    printf("%-40s %8.2f ms  %7.0f GB/s  (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
    const int n = 256, nf = argc > 1 ? atoi(argv[1]) : 64;
    const size_t N = (size_t)n * n * n;
    double *u, *sl, *o;
    unsigned long long *q;
    cudaMalloc(&u, N * nf * 8);
    cudaMalloc(&o, N * nf * 8);
    cudaMalloc(&sl, N * 2 * 8);
    cudaMalloc(&q, 8);
    cudaMemset(u, 0, N * nf * 8);
    cudaMemset(sl, 0, N * 2 * 8);
    run<0>(u, sl, o, n, nf, q, "natural, 8x8 bricks (64 B pieces)");
    run<1>(u, sl, o, n, nf, q, "natural, 16x4 bricks (128 B pieces)");
    run<2>(u, sl, o, n, nf, q, "natural, 32x2 bricks (256 B pieces)");
    run<3>(u, sl, o, n, nf, q, "blocked, 512 B per brick plane");
    return 0;
}
