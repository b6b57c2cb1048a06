// gather_probe_narrow.cu -- ceiling of a random gather of NARROW rows (64 / 128 / 256 bytes) on one B200.
//
// Companion of gather_probe.cu for the column-sharded layout (every GPU holds D/R columns of every row): the access
// pattern of csrc/spmm_narrow2.inl without any row bookkeeping -- LPR lanes read one row (one float4 each), a warp
// works on 32 / LPR rows at once, sub-warp `sub` walks 32 consecutive indices, U loads in flight per lane, everything
// summed into one float4 per lane (written once per warp chunk so the loads stay live).
// usage: gather_probe_narrow N E dist(0 uniform | 1 skewed)
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                       \
        }                                                                                  \
    } while (0)

template <int LPR, int U>
__global__ void __launch_bounds__(256) k_narrow(const float *__restrict__ x, const unsigned *__restrict__ idx,
                                                float *out, long long E, int T) {
    constexpr int EPW = 32 / LPR, CPL = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const long long task = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long t_beg = task * T;
    if (t_beg >= E) return;
    const long long t_end = (t_beg + T < E) ? t_beg + T : E;
    const int sub = lane / LPR, li = lane % LPR;
    const char *xl = reinterpret_cast<const char *>(x) + li * 16;
    float4 acc = make_float4(0, 0, 0, 0);
    for (long long chunk = t_beg; chunk < t_end; chunk += EPW * 32) {
        const long long rb = chunk + sub * 32;
        unsigned creg[CPL];
#pragma unroll
        for (int r = 0; r < CPL; ++r) creg[r] = __ldcs(idx + rb + li * CPL + r);
#pragma unroll 1
        for (int g = 0; g < 32 / U; ++g) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned c = __shfl_sync(~0u, creg[u % CPL], sub * LPR + g * (U / CPL) + u / CPL);
                v[u] = __ldg(reinterpret_cast<const float4 *>(xl + (size_t)c * (LPR * 16)));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
            }
        }
    }
    *reinterpret_cast<float4 *>(out + task * 128 + lane * 4) = acc;
}

__global__ void fill_x(float *x, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = (float)(i % 1021) * 0.001f;
}
__global__ void fill_idx(unsigned *idx, long long E, unsigned N, int dist) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0x1234567ull;
        z ^= z >> 31; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 29; z *= 0x94D049BB133111EBull; z ^= z >> 32;
        double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
        if (dist == 1) u = u * u * u * u;
        unsigned r = (unsigned)(u * N);
        if (r >= N) r = N - 1;
        if (dist == 1) r = (unsigned)(((unsigned long long)r * 2654435761ull) % N);
        idx[i] = r;
    }
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

template <int LPR, int U>
static void run(const float *x, const unsigned *idx, float *out, long long E, int T, const char *label) {
    const long long tasks = (E + T - 1) / T;
    const unsigned grid = (unsigned)((tasks * 32 + 255) / 256);
    float ms = time_ms([&] { k_narrow<LPR, U><<<grid, 256>>>(x, idx, out, E, T); }, 5);
    CK(cudaGetLastError());
    const double gb = (double)E * LPR * 16 / 1e9;
    printf("%s rows of %3d B, U=%d : %8.3f ms  %7.1f GB/s gathered  %6.2f G rows/s\n", label, LPR * 16, U, ms,
           gb / ms * 1e3, E / ms * 1e-6);
}

int main(int argc, char **argv) {
    const long long N = argc > 1 ? atoll(argv[1]) : 10000000;
    long long E = argc > 2 ? atoll(argv[2]) : 100000000;
    const int dist = argc > 3 ? atoi(argv[3]) : 0;
    const int T = 2048;
    E = E / T * T;
    float *x, *out;
    unsigned *idx;
    CK(cudaMalloc(&x, N * 256));  // widest case: 256-byte rows
    CK(cudaMalloc(&idx, E * 4));
    CK(cudaMalloc(&out, (E / T) * 512));
    fill_x<<<148 * 8, 256>>>(x, N * 64);
    fill_idx<<<148 * 8, 256>>>(idx, E, (unsigned)N, dist);
    CK(cudaDeviceSynchronize());
    printf("N=%lld E=%lld dist=%d\n", N, E, dist);
    run<4, 8>(x, idx, out, E, T, "LDG");
    run<4, 4>(x, idx, out, E, T, "LDG");
    run<8, 8>(x, idx, out, E, T, "LDG");
    run<16, 8>(x, idx, out, E, T, "LDG");
    return 0;
}
