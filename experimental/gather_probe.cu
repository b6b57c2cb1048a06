// gather_probe.cu -- what is the ceiling of a random 512-byte row gather on one B200, by mechanism?
//
// Standalone (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_probe gather_probe.cu).
// Not part of libpglb.so.  The product kernel (csrc/spmm_stream.cu) gathers with one LDGSTS per lane
// per row; VERDICT r1 asks whether TMA (one bulk copy per row, or one gather4 per four rows) lifts the
// ceiling.  This probe strips everything but the gather + an in-order sum so the mechanisms can be
// compared on equal terms:
//   mode 0  LDG.128 per lane per row, 8 rows in flight per lane (registers)
//   mode 1  cp.async.bulk (UBLKCP) 512 B per row from an elected lane, mbarrier per group of 8 rows
//   mode 2  cp.async.bulk.tensor.2d.tile::gather4 (UTMALDG.2D.GATHER4), 4 rows per instruction
//   mode 3  LDGSTS per lane per row (cp.async.cg 16 B), commit groups of 4 (the product's mechanism)
// Every warp reduces chunks of T indices; the sum is written once per chunk (keeps the loads live).
// usage: gather_probe N E dist(0 uniform | 1 skewed) [T]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ float4 ldg128(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned cnt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    // bounded: a wrong tensor-map box (fewer bytes than expected) must trap, not hang the box
    for (unsigned spin = 0; spin < (1u << 22); ++spin) {
        unsigned ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_row(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void gather4(unsigned dst, const CUtensorMap *tm, int c0, int r0, int r1, int r2, int r3,
                                        unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, "
        "%6}], [%7];" ::"r"(dst),
        "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
        : "memory");
}

// ---- mode 0 ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ldg(const float *__restrict__ x, const unsigned *__restrict__ idx, float *out,
                                             long long E, int T) {
    const int lane = threadIdx.x & 31;
    const long long chunk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long a = chunk * T;
    if (a >= E) return;
    const int cnt = (int)((E - a) < T ? (E - a) : T);
    float4 acc = make_float4(0, 0, 0, 0);
    for (int b = 0; b < cnt; b += 32) {
        const unsigned c = (b + lane < cnt) ? __ldcs(idx + a + b + lane) : 0u;
        const int nb = (cnt - b) < 32 ? (cnt - b) : 32;
        for (int k = 0; k < nb; k += 8) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned cj = __shfl_sync(~0u, c, k + j);
                v[j] = (k + j < nb) ? ldg128(x + (size_t)cj * 128 + lane * 4) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w;
            }
        }
    }
    *reinterpret_cast<float4 *>(out + chunk * 128 + lane * 4) = acc;
}

// ---- modes 1, 2 -----------------------------------------------------------------------------------
// per warp: NG group slots of GR rows (GR*512 B each), one mbarrier per slot; LAG = NG - 1 groups in flight
template <int MODE, int NG, int GR, int W>
__global__ void __launch_bounds__(W * 32) k_tma(const float *__restrict__ x, const __grid_constant__ CUtensorMap tm,
                                                 const unsigned *__restrict__ idx, float *out, long long E, int T) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[W * NG];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long chunk = (long long)blockIdx.x * W + wib;
    const long long a = chunk * T;
    const unsigned ring = (unsigned)__cvta_generic_to_shared(smem) + wib * (NG * GR * 512);
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&bars[wib * NG]);
    if (lane == 0)
        for (int s = 0; s < NG; ++s) mbar_init(bar0 + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (a >= E) return;
    const int cnt = (int)((E - a) < T ? (E - a) : T);  // multiple of 32 in this probe
    const int ngroups = cnt / GR;
    constexpr int LAG = NG - 1;
    constexpr int GPB = 32 / GR;
    float4 acc = make_float4(0, 0, 0, 0);
    unsigned c = __ldcs(idx + a + lane);
#pragma unroll 1
    for (int g = 0; g < ngroups + LAG; ++g) {
        if (g < ngroups) {
            const int sub = g % GPB;
            if (sub == 0 && g > 0) c = __ldcs(idx + a + g * GR + lane);
            const int s = g % NG;
            __syncwarp();  // every lane is done reading slot s (consumed LAG+1 groups ago)
            if (MODE == 1) {
                if (lane == 0) mbar_expect(bar0 + s * 8, GR * 512);
#pragma unroll
                for (int k = 0; k < GR; ++k) {
                    const unsigned ck = __shfl_sync(~0u, c, sub * GR + k);
                    if (lane == 0) bulk_row(ring + (s * GR + k) * 512, x + (size_t)ck * 128, 512, bar0 + s * 8);
                }
            } else {
                if (lane == 0) mbar_expect(bar0 + s * 8, GR * 512);
#pragma unroll
                for (int k = 0; k < GR; k += 4) {
                    const int r0 = __shfl_sync(~0u, c, sub * GR + k), r1 = __shfl_sync(~0u, c, sub * GR + k + 1);
                    const int r2 = __shfl_sync(~0u, c, sub * GR + k + 2), r3 = __shfl_sync(~0u, c, sub * GR + k + 3);
                    if (lane == 0) gather4(ring + (s * GR + k) * 512, &tm, 0, r0, r1, r2, r3, bar0 + s * 8);
                }
            }
        }
        if (g >= LAG) {
            const int gc = g - LAG;
            const int s = gc % NG;
            mbar_wait(bar0 + s * 8, (gc / NG) & 1);
            const unsigned base = ring + s * GR * 512 + lane * 16;
#pragma unroll
            for (int k = 0; k < GR; ++k) {
                const float4 v = lds128(base + k * 512);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
    }
    *reinterpret_cast<float4 *>(out + chunk * 128 + lane * 4) = acc;
}

// ---- mode 3 --------------------------------------------------------------------------------------
template <int RING, int GRP, int W>
__global__ void __launch_bounds__(W * 32) k_ldgsts(const float *__restrict__ x, const unsigned *__restrict__ idx,
                                                   float *out, long long E, int T) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long chunk = (long long)blockIdx.x * W + wib;
    const long long a = chunk * T;
    if (a >= E) return;
    const unsigned ring = (unsigned)__cvta_generic_to_shared(smem) + wib * (RING * 512) + lane * 16;
    const int cnt = (int)((E - a) < T ? (E - a) : T);
    const int ngroups = cnt / GRP;
    constexpr int LAG = RING / GRP - 1, GPB = 32 / GRP, RG = RING / GRP;
    const char *xl = reinterpret_cast<const char *>(x) + lane * 16;
    float4 acc = make_float4(0, 0, 0, 0);
    unsigned c = __ldcs(idx + a + lane);
#pragma unroll 1
    for (int g = 0; g < ngroups + LAG; ++g) {
        if (g < ngroups) {
            const int sub = g % GPB;
            if (sub == 0 && g > 0) c = __ldcs(idx + a + g * GRP + lane);
            const unsigned ga = ring + (g % RG) * (GRP * 512);
#pragma unroll
            for (int k = 0; k < GRP; ++k) {
                const unsigned ck = __shfl_sync(~0u, c, sub * GRP + k);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ga + k * 512), "l"(xl + (size_t)ck * 512)
                             : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (g >= LAG) {
            asm volatile("cp.async.wait_group %0;" ::"n"(LAG) : "memory");
            const unsigned ga = ring + ((g - LAG) % RG) * (GRP * 512);
#pragma unroll
            for (int k = 0; k < GRP; ++k) {
                const float4 v = lds128(ga + k * 512);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
    }
    *reinterpret_cast<float4 *>(out + chunk * 128 + lane * 4) = acc;
}

__global__ void fill_x(float *x, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i >> 7;
        x[i] = (float)((row * 7 + (i & 127)) % 1021) * 0.001f;
    }
}
__global__ void fill_idx(unsigned *idx, long long E, unsigned N, int dist) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0x1234567ull;
        z ^= z >> 31; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 29; z *= 0x94D049BB133111EBull; z ^= z >> 32;
        double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
        if (dist == 1) u = u * u * u * u;  // skewed: a quarter of the draws hit the first 0.4 % of the rows
        unsigned r = (unsigned)(u * N);
        if (r >= N) r = N - 1;
        if (dist == 1) r = (unsigned)(((unsigned long long)r * 2654435761ull) % N);  // scatter the hot rows
        idx[i] = r;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap *tm, float *x, unsigned long long N, unsigned box_rows) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn) return false;
    cuuint64_t dims[2] = {128, N};
    cuuint64_t strides[1] = {512};
    cuuint32_t box[2] = {128, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled(box_rows=%u) -> %d\n", box_rows, (int)r);
    return r == CUDA_SUCCESS;
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

static double checksum(const float *d_out, long long n) {
    std::vector<float> h(n);
    CK(cudaMemcpy(h.data(), d_out, n * 4, cudaMemcpyDeviceToHost));
    double s = 0;
    for (long long i = 0; i < n; ++i) s += h[i];
    return s;
}

int main(int argc, char **argv) {
    const long long N = argc > 1 ? atoll(argv[1]) : 10000000;
    long long E = argc > 2 ? atoll(argv[2]) : 100000000;
    const int dist = argc > 3 ? atoi(argv[3]) : 0;
    const int T = argc > 4 ? atoi(argv[4]) : 2048;
    E = E / T * T;
    float *x, *out;
    unsigned *idx;
    CK(cudaMalloc(&x, N * 512));
    CK(cudaMalloc(&idx, E * 4));
    const long long chunks = E / T;
    CK(cudaMalloc(&out, chunks * 512));
    fill_x<<<148 * 8, 256>>>(x, N * 128);
    fill_idx<<<148 * 8, 256>>>(idx, E, (unsigned)N, dist);
    CK(cudaDeviceSynchronize());
    const double gb = (double)E * 512 / 1e9;
    printf("N=%lld E=%lld dist=%d T=%d gather bytes %.2f GB\n", N, E, dist, T, gb);

    // mode 0
    {
        const int wpb = 8;
        const unsigned grid = (unsigned)((chunks + wpb - 1) / wpb);
        float ms = time_ms([&] { k_ldg<<<grid, wpb * 32>>>(x, idx, out, E, T); }, 5);
        CK(cudaGetLastError());
        printf("mode0 LDG.128 x8        : %8.3f ms  %7.1f GB/s  checksum %.6e\n", ms, gb / ms * 1e3, checksum(out, chunks * 128));
    }
    const double ref = checksum(out, chunks * 128);
    // mode 3: product geometry (16-slot ring, groups of 4, 14 warps x 2 CTAs) and a deeper one
    {
        constexpr int W = 14;
        CK(cudaFuncSetAttribute(k_ldgsts<16, 4, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 16 * 512));
        const unsigned grid = (unsigned)((chunks + W - 1) / W);
        float ms = time_ms([&] { k_ldgsts<16, 4, W><<<grid, W * 32, W * 16 * 512>>>(x, idx, out, E, T); }, 5);
        CK(cudaGetLastError());
        printf("mode3 LDGSTS r16 g4 w14 : %8.3f ms  %7.1f GB/s  checksum %.6e\n", ms, gb / ms * 1e3, checksum(out, chunks * 128));
    }
    {
        constexpr int W = 7;
        CK(cudaFuncSetAttribute(k_ldgsts<32, 8, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 32 * 512));
        const unsigned grid = (unsigned)((chunks + W - 1) / W);
        float ms = time_ms([&] { k_ldgsts<32, 8, W><<<grid, W * 32, W * 32 * 512>>>(x, idx, out, E, T); }, 5);
        CK(cudaGetLastError());
        printf("mode3 LDGSTS r32 g8 w7  : %8.3f ms  %7.1f GB/s  checksum %.6e\n", ms, gb / ms * 1e3, checksum(out, chunks * 128));
    }
    CUtensorMap tm1, tm4;
    const bool ok1 = make_map(&tm1, x, N, 1), ok4 = make_map(&tm4, x, N, 4);
    printf("tensor map box rows 1: %s, box rows 4: %s\n", ok1 ? "ok" : "rejected", ok4 ? "ok" : "rejected");
#define RUN_TMA(MODE, NG, GR, W, TM, LABEL)                                                                         \
    {                                                                                                               \
        auto kern = k_tma<MODE, NG, GR, W>;                                                                         \
        const int smem = W * NG * GR * 512;                                                                         \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                          \
        const unsigned grid = (unsigned)((chunks + W - 1) / W);                                                     \
        CK(cudaMemset(out, 0, chunks * 512));                                                                       \
        float ms = time_ms([&] { kern<<<grid, W * 32, smem>>>(x, TM, idx, out, E, T); }, 5);                        \
        cudaError_t e = cudaDeviceSynchronize();                                                                    \
        if (e != cudaSuccess) {                                                                                     \
            printf("%s : FAILED %s\n", LABEL, cudaGetErrorString(e));                                               \
            return 3;                                                                                               \
        }                                                                                                           \
        const double cs = checksum(out, chunks * 128);                                                              \
        printf("%s : %8.3f ms  %7.1f GB/s  checksum %.6e %s\n", LABEL, ms, gb / ms * 1e3, cs,                       \
               fabs(cs - ref) <= 1e-6 * fabs(ref) ? "OK" : "MISMATCH");                                             \
    }
    // bulk copy per row: smem per warp = NG*GR*512
    RUN_TMA(1, 4, 8, 7, tm1, "mode1 bulk  ng4 gr8 w7  ");    // 16 KB/warp, 112 KB/CTA x2
    RUN_TMA(1, 4, 4, 14, tm1, "mode1 bulk  ng4 gr4 w14 ");   // 8 KB/warp
    RUN_TMA(1, 8, 4, 7, tm1, "mode1 bulk  ng8 gr4 w7  ");
    RUN_TMA(1, 4, 8, 4, tm1, "mode1 bulk  ng4 gr8 w4  ");    // 64 KB/CTA x3
    if (ok1) {
        RUN_TMA(2, 4, 8, 7, tm1, "mode2 gat4/b1 ng4 gr8 w7 ");
        RUN_TMA(2, 4, 4, 14, tm1, "mode2 gat4/b1 ng4 gr4 w14");
        RUN_TMA(2, 8, 4, 7, tm1, "mode2 gat4/b1 ng8 gr4 w7 ");
        RUN_TMA(2, 4, 8, 4, tm1, "mode2 gat4/b1 ng4 gr8 w4 ");
        RUN_TMA(2, 2, 16, 7, tm1, "mode2 gat4/b1 ng2 gr16 w7");
    }
    if (ok4) {
        RUN_TMA(2, 4, 8, 7, tm4, "mode2 gat4/b4 ng4 gr8 w7 ");
    }
    return 0;
}
