// solver_bench.cu -- throughput of the local Godunov solver alone (no memory traffic): how many
// node-updates/s the fp64 arithmetic of fsm_solve.cuh sustains on one B200 for NC independent chains
// per lane and W warps per SM.  Development aid for profiles/kernel_evolution_r2.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I mceik_b200/csrc tools/solver_bench.cu -o build/solver_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "fsm_solve.cuh"

using namespace mceik::fsm;

template <int NC>
__global__ void __launch_bounds__(NC == 1 ? 1024 : (NC == 2 ? 768 : 384), 1) bench(double *out, int iters) {
    double a[NC], b[NC], c[NC], f[NC], r[NC];
    bool act[NC];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        a[q] = 10.0 + 1e-3 * (t % 977) + 0.01 * q;
        f[q] = 0.2 + 1e-4 * (t % 13);
        b[q] = a[q] + 0.3 * f[q];
        c[q] = a[q] + 0.5 * f[q];
        act[q] = true;
    }
    for (int it = 0; it < iters; ++it) {
        local_solve_xn<NC>(a, b, c, f, act, r);
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            // the recurrence of the sweep: the new value is the next node's z neighbour
            const double m = dmin2(r[q], c[q]);
            c[q] = __dmul_rn(m, 0.99999);
            a[q] = __dmul_rn(a[q], 1.0000001);
        }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < NC; ++q) s += c[q];
    out[t] = s;
}

template <int NC>
void run(int warps, int iters, double *d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int nsm = 148;
    bench<NC><<<nsm, warps * 32>>>(d_out, iters / 8);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<NC><<<nsm, warps * 32>>>(d_out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double upd = (double)nsm * warps * 32 * NC * iters;
    printf("NC=%d warps/SM=%2d: %8.2f ms  %7.1f G solves/s  (%s)\n", NC, warps, ms, upd / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    double *d_out;
    cudaMalloc(&d_out, sizeof(double) * 148 * 1024);
    for (int w : {4, 8, 12, 16, 24, 32}) run<1>(w, iters, d_out);
    for (int w : {4, 8, 12, 16, 24}) run<2>(w, iters, d_out);
    for (int w : {4, 6, 8, 12}) run<4>(w, iters, d_out);
    return 0;
}
