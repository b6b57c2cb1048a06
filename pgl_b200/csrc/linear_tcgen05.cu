// linear_tcgen05.cu -- out[M,128] = act(x[M,128] @ W[128,128] + bias) on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM, x tiles staged by TMA), with 3xTF32 error compensation
// (x_lo*w_hi + x_hi*w_lo + x_hi*w_hi accumulated in one fp32 TMEM accumulator: fp32-level accuracy).
// The K = N = 128 fast path of pglb_linear_tf32x3_f32 (the dense transform of the conv layers, reference
// pgl/nn/conv.py:238-251); other shapes stay on the mma.sync kernel in linear_tc.cu.
//
// Round 1 shipped the mma.sync version (6.44 ms for the 10M-row cfg5 layer against a 1.56 ms HBM bound) and a
// blind tcgen05 draft (experimental/linear_tcgen05.cu) that turned out correct on hardware but slow (1.63 ms per 2M
// rows): its 128 producer threads kept only 16 KB of loads in flight per SM and its epilogue wrote one 16-byte piece
// per lane into 32 different rows.  This version changes three things:
//   * x tiles arrive by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B): one elected thread keeps STAGES x 16 KB in
//     flight, and the bytes land directly in the canonical K-major layout the MMA reads.  The raw fp32 tile IS the
//     "hi" operand (kind::tf32 ignores the low 13 mantissa bits); four splitter warps derive lo = x - trunc_tf32(x).
//   * the MMA computes out^T = W^T x^T (A = W^T from shared memory, B = the x tile): the accumulator's TMEM lane is
//     the OUTPUT COLUMN, so when the epilogue warps read it back (tcgen05.ld 32x32b: lane = column n, registers =
//     32 consecutive rows m) every store instruction writes one 128-byte row segment -- coalesced without staging.
//   * bias + ReLU are applied in registers on the way out (one bias value per thread).
// One persistent CTA per SM, 10 warps: 0-3 epilogue (TMEM lane quadrant = warp), 4 MMA issuer, 5 TMA producer,
// 6-9 splitters.  Shared memory: W^T hi / lo (128 KB, built once per CTA), STAGES x (raw + lo) K blocks of the x tile.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace pglb {
namespace tc5 {

constexpr int BM = 128, BN = 128, BK = 32, KTOT = 128, KBLOCKS = KTOT / BK;
constexpr int STAGES = 3;
constexpr int KB_BYTES = BM * BK * 4;                  // 16 KB: one [128][32] fp32 block
constexpr int W_BYTES = 2 * KBLOCKS * KB_BYTES;        // W^T hi + lo
constexpr int A_STAGE_BYTES = 2 * KB_BYTES;            // raw (= hi) + lo
constexpr int SMEM_BYTES = 1024 + W_BYTES + STAGES * A_STAGE_BYTES + 256;
constexpr int NUM_EPI_WARPS = 4, NUM_SPLIT_WARPS = 4;
constexpr int THREADS = (NUM_EPI_WARPS + 2 + NUM_SPLIT_WARPS) * 32;
constexpr uint32_t TMEM_COLS = 256;                    // two 128-column accumulators

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// a barrier that never completes (a protocol bug, a tensor map that does not describe the buffer) must surface as a
// launch failure on the host, not as a hung device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try(bar, parity))
        if (++spins > (1u << 24)) __trap();
}

// K-major SWIZZLE_128B operand block [rows][32 fp32]: row r at r * 128 B, its 16-byte chunk c at (c ^ (r & 7)).
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 [0,14) | LBO >> 4 [16,30) |
// SBO >> 4 [32,46) | version 1 [46,48) | layout SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;  // 8 rows x 128 B between core-matrix groups
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D fp32 [4,6) = 1 | A tf32 [7,10) = 2 | B tf32 [10,13) = 2 | K-major A, B | N >> 3 [17,23) | M >> 4 [24,29)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BM >> 3) << 17) | ((uint32_t)(BN >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
linear_tcgen05_kernel(const __grid_constant__ CUtensorMap tmx, const float *__restrict__ w,
                      const float *__restrict__ bias, float *__restrict__ out, int64_t ldo, int64_t M, int act) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *w_hi = smem;                            // [KBLOCKS][128 n][32 k]
    unsigned char *w_lo = smem + KBLOCKS * KB_BYTES;
    unsigned char *a_st = smem + W_BYTES;                  // [STAGES][raw, lo][128 m][32 k]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + W_BYTES + STAGES * A_STAGE_BYTES);
    // barriers: raw_full[S] (TMA bytes landed), a_full[S] (lo written), a_empty[S] (MMAs done with the stage),
    // t_full[2] (accumulator complete), t_empty[2] (accumulator drained)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * STAGES + 4);
    const uint32_t bar0 = smem_u32(bars);
    auto raw_full = [&](int s) { return bar0 + 8 * s; };
    auto a_full = [&](int s) { return bar0 + 8 * (STAGES + s); };
    auto a_empty = [&](int s) { return bar0 + 8 * (2 * STAGES + s); };
    auto t_full = [&](int s) { return bar0 + 8 * (3 * STAGES + s); };
    auto t_empty = [&](int s) { return bar0 + 8 * (3 * STAGES + 2 + s); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: barriers, TMEM, W^T split into hi / lo in the swizzled K-major layout ----
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(raw_full(s), 1);
            mbar_init(a_full(s), NUM_SPLIT_WARPS * 32);
            mbar_init(a_empty(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(t_full(s), 1);
            mbar_init(t_empty(s), NUM_EPI_WARPS * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < KTOT * BN; i += THREADS) {
        const int k = i / BN, n = i - k * BN;              // coalesced read of W[k][n]
        const float f = w[i];
        const uint32_t hi = to_tf32(f);
        const uint32_t lo = to_tf32(f - __uint_as_float(hi));
        const int kb = k / BK, kk = k - kb * BK;
        const uint32_t off = kb * KB_BYTES + sw128_offset(n, kk >> 2) + (kk & 3) * 4;
        *reinterpret_cast<uint32_t *>(w_hi + off) = hi;
        *reinterpret_cast<uint32_t *>(w_lo + off) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int64_t tiles = (M + BM - 1) / BM;

    if (warp == 5) {
        // ================= TMA producer (one thread) =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                for (int kb = 0; kb < KBLOCKS; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(a_empty(s), ph ^ 1);         // MMAs have drained this stage (passes on the first lap)
                    mbar_expect_tx(raw_full(s), KB_BYTES);
                    // rows past M are zero-filled by the TMA unit
                    tma_load_2d(smem_u32(a_st + s * A_STAGE_BYTES), &tmx, kb * BK, (int)(tile * BM), raw_full(s));
                }
            }
        }
    } else if (warp >= 6) {
        // ================= splitters: lo = x - trunc_tf32(x), same swizzled position =================
        const int pt = tid - 6 * 32;                       // 0..127
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            for (int kb = 0; kb < KBLOCKS; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(raw_full(s), ph);
                const unsigned char *raw = a_st + s * A_STAGE_BYTES;
                unsigned char *lo_b = a_st + s * A_STAGE_BYTES + KB_BYTES;
#pragma unroll
                for (int i = 0; i < KB_BYTES / 16 / (NUM_SPLIT_WARPS * 32); ++i) {
                    const int off = (pt + i * NUM_SPLIT_WARPS * 32) * 16;   // any 16-byte piece: lo sits where hi sits
                    const float4 v = *reinterpret_cast<const float4 *>(raw + off);
                    float4 l;
                    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    *reinterpret_cast<float4 *>(lo_b + off) = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(a_full(s));
            }
        }
    } else if (warp == 4) {
        // ================= MMA issuer =================
        uint32_t it = 0, tcount = 0;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tcount) {
            const int as = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            mbar_wait(t_empty(as), aph ^ 1);               // the epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + as * BM;
            for (int kb = 0; kb < KBLOCKS; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(a_full(s), ph);                  // raw landed (the splitters waited for it) and lo written
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t x_hi = smem_u32(a_st + s * A_STAGE_BYTES);
                    const uint32_t x_lo = x_hi + KB_BYTES;
                    const uint32_t wt_hi = smem_u32(w_hi + kb * KB_BYTES);
                    const uint32_t wt_lo = smem_u32(w_lo + kb * KB_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {     // 32 bytes of K per instruction inside the 128-B atom
                        const uint32_t ko = k * 32;
                        // D[n][m] += W^T[n][k] x[m][k]: A = W^T (rows = output columns), B = x tile
                        umma_tf32(tmem_d, make_desc(wt_hi + ko), make_desc(x_lo + ko), (kb | k) ? 1u : 0u);
                        umma_tf32(tmem_d, make_desc(wt_lo + ko), make_desc(x_hi + ko), 1u);
                        umma_tf32(tmem_d, make_desc(wt_hi + ko), make_desc(x_hi + ko), 1u);
                    }
                    umma_commit(a_empty(s));               // stage reusable once these MMAs have read it
                    if (kb == KBLOCKS - 1) umma_commit(t_full(as));
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue (warps 0-3: TMEM lanes 32 * warp .. +32 = output columns) =================
        const int n = warp * 32 + lane;                    // my output column
        const float b = bias ? __ldg(bias + n) : 0.0f;
        uint32_t tcount = 0;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tcount) {
            const int as = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            mbar_wait(t_full(as), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t row0 = tile * BM;
#pragma unroll 1
            for (int m0 = 0; m0 < BM; m0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * BM + m0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
                      "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
                      "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
                      "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float *orow = out + (row0 + m0) * ldo + n;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float v = __uint_as_float(r[j]) + b;
                    if (act == 1) v = fmaxf(v, 0.0f);
                    if (row0 + m0 + j < M) __stcs(orow + (int64_t)j * ldo, v);   // 32 lanes = 128 contiguous bytes of row m
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(t_empty(as));
        }
    }

    // ---- teardown ----
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*TmapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace tc5

// PGLB_LINEAR_TCGEN05 = 0 keeps every shape on the mma.sync kernel (linear_tc.cu)
bool linear_tcgen05_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PGLB_LINEAR_TCGEN05");
        v = (e && atoi(e) == 0) ? 0 : 1;
    }
    return v == 1;
}

// K = N = 128, x rows 16-byte aligned (ldx % 4 == 0): checked by the caller
int linear_tcgen05_run(const float *x, int64_t ldx, const float *w, const float *bias, float *out, int64_t ldo,
                       int64_t M, int act, cudaStream_t stream) {
    using namespace tc5;
    static std::atomic<void *> enc_fn{nullptr};
    void *f = enc_fn.load(std::memory_order_acquire);
    if (!f) {
        cudaDriverEntryPointQueryResult q;
        void *g = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &g, cudaEnableDefault, &q) == cudaSuccess && g) {
            enc_fn.store(g, std::memory_order_release);
            f = g;
        }
    }
    PGLB_CHECK_ARG(f != nullptr, PGLB_EINVAL, "cuTensorMapEncodeTiled not available from the driver");
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)KTOT, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)ldx * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = reinterpret_cast<TmapEncodeFn>(f)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), dims,
                                                         strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PGLB_CHECK_ARG(r == CUDA_SUCCESS, PGLB_EINVAL, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    static std::atomic<unsigned long long> attr_done{0};
    PGLB_CUDA(ensure_dyn_smem(linear_tcgen05_kernel, SMEM_BYTES, attr_done));
    const int64_t tiles = (M + BM - 1) / BM;
    const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
    linear_tcgen05_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tm, w, bias, out, ldo, M, act);
    PGLB_LAUNCH_CHECK("linear_tcgen05_kernel");
    return PGLB_OK;
}

}  // namespace pglb
