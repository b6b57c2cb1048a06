// common.cuh -- shared helpers for libpglb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "pglb.h"

namespace pglb {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

int fail(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define PGLB_CHECK_ARG(cond, code, ...)                  \
    do {                                                 \
        if (!(cond)) return ::pglb::fail(code, __VA_ARGS__); \
    } while (0)

#define PGLB_CUDA(call)                                             \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return ::pglb::cuda_fail(_e, #call); \
    } while (0)

#define PGLB_LAUNCH_CHECK(name)                                     \
    do {                                                            \
        ::pglb::count_launch();                                     \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return ::pglb::cuda_fail(_e, name);  \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}

// SM count of the CURRENT device (cached per device: a process may drive several GPUs).
inline int sm_count() {
    static std::atomic<int> cache[64];
    const int dev = current_device();
    int n = cache[dev & 63].load(std::memory_order_relaxed);
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cache[dev & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}

// Opt a kernel into more than 48 KB of dynamic shared memory.  The attribute is per device, so the
// "already done" flag is a bit per device (one mask per kernel instantiation, owned by the caller).
template <typename K>
inline cudaError_t ensure_dyn_smem(K kernel, int bytes, std::atomic<unsigned long long> &done) {
    const unsigned long long bit = 1ull << (current_device() & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

// multi-block int64 prefix sum (csr_build.cu): in and out may alias, tmp holds scan_i64_ws_bytes(n) bytes
size_t scan_i64_ws_bytes(int64_t n);
int scan_i64(const int64_t *in, int64_t *out, int64_t n, int inclusive, void *tmp, cudaStream_t stream);

// ---- device helpers -------------------------------------------------------------------

// streaming (evict-first) loads for index data that is read exactly once
__device__ __forceinline__ int64_t ld_stream(const int64_t *p) { return __ldcs((const long long *)p); }
__device__ __forceinline__ int64_t ld_ro(const int64_t *p) { return __ldg((const long long *)p); }

}  // namespace pglb
