// common.cuh -- shared helpers for libmceik_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

namespace mceik {

// Thread-local last error text exposed through mceik_last_error().
void set_error(const char *fmt, ...);
const char *last_error();
// Every kernel launch of this library is counted (bench.py reports it as gpu_launches).
void count_launch(long long n = 1);
long long launch_count();

struct CudaError : public std::runtime_error {
    explicit CudaError(const std::string &s) : std::runtime_error(s) {}
};

#define MCEIK_CUDA(expr)                                                                     \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            char _b[512];                                                                    \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                    \
            throw ::mceik::CudaError(_b);                                                    \
        }                                                                                    \
    } while (0)

#define MCEIK_LAUNCH_CHECK()                      \
    do {                                          \
        ::mceik::count_launch();                  \
        MCEIK_CUDA(cudaGetLastError());           \
    } while (0)

// Growable device buffer owned by a context (never shrinks).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void *ensure(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFree(p);
            p = nullptr;
            cap = 0;
            MCEIK_CUDA(cudaMalloc(&p, bytes));
            cap = bytes;
        }
        return p;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

#ifdef __CUDACC__
// u_nan of the reference: HUGE(1.d0) (module.F90:419)
__device__ __forceinline__ double huge_val() { return DBL_MAX; }

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int *p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif

}  // namespace mceik
