// linear_tcgen05.cu -- EXPERIMENTAL, NOT PART OF libpglb.so, NEVER RUN ON HARDWARE YET.
//
// tcgen05 / TMEM version of pgl_b200/csrc/linear_tc.cu: out[M,128] = act(x[M,128] @ W[128,128] + bias)
// with 3xTF32 error compensation (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo accumulated in one fp32 TMEM
// accumulator).  Written blind at the end of round 1 (no GPU minutes left) as the starting point for the
// "tcgen05 version of the dense transform" item of DESIGN.md section 4.11; experimental/check_linear_tcgen05.py
// builds it on its own and checks it against an fp64 matmul.  Run that under `timeout` first: a wrong
// barrier protocol hangs instead of failing.
//
// Structure (one persistent CTA per SM, 9 warps):
//   warps 0-3  epilogue   tcgen05.ld (32 lanes x 32 columns per call) -> bias, ReLU -> st.global
//   warp  4    MMA issuer one elected lane: per 32-wide K block 4 x 3 tcgen05.mma kind::tf32 (M128 N128 K8)
//   warps 5-8  producers  ld.global x (coalesced 128-B row segments) -> hi/lo split in registers ->
//                         st.shared into the canonical K-major SWIZZLE_128B layout -> fence.proxy.async
// Shared memory: W^T hi / lo, 4 K blocks x [128 n][32 k] each (128 KB, loaded once), 2 stages of
// x hi / lo K blocks [128 m][32 k] (64 KB), mbarriers.  TMEM: 2 x 128 columns (double-buffered accumulator).
//
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (CUTLASS): shared-memory matrix descriptor
// = start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type [61,64)
// (SWIZZLE_128B = 2); instruction descriptor = c_format F32=1 [4,6) | a_format TF32=2 [7,10) |
// b_format TF32=2 [10,13) | K-major A,B | N>>3 [17,23) | M>>4 [24,29).
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BM = 128, BN = 128, BK = 32, KTOT = 128, KBLOCKS = KTOT / BK;
constexpr int STAGES = 2;
constexpr int KB_BYTES = BM * BK * 4;                  // 16 KB: one [128][32] tf32 block
constexpr int W_BYTES = 2 * KBLOCKS * KB_BYTES;        // hi + lo
constexpr int A_STAGE_BYTES = 2 * KB_BYTES;            // hi + lo
constexpr int SMEM_BYTES = 1024 + W_BYTES + STAGES * A_STAGE_BYTES + 256;
constexpr int NUM_EPI_WARPS = 4, NUM_PROD_WARPS = 4;
constexpr int THREADS = (NUM_EPI_WARPS + 1 + NUM_PROD_WARPS) * 32;
constexpr uint32_t TMEM_COLS = 256;

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}

// K-major SWIZZLE_128B operand block [rows][32 tf32]: row r at r * 128 B, its 16-byte chunk c at (c ^ (r & 7)).
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);        // start address
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}

constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

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

__global__ void __launch_bounds__(THREADS, 1)
linear_tcgen05_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ w,
                      const float *__restrict__ bias, float *__restrict__ out, int64_t ldo, int64_t M, int act) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *w_hi = smem;                            // [KBLOCKS][128 n][32 k]
    unsigned char *w_lo = smem + KBLOCKS * KB_BYTES;
    unsigned char *a_st = smem + W_BYTES;                  // [STAGES][hi, lo][128 m][32 k]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + W_BYTES + STAGES * A_STAGE_BYTES);
    // bars: a_full[2], a_empty[2], t_full[2], t_empty[2]; then the TMEM base address
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    const uint32_t bar0 = smem_u32(bars);
    auto a_full = [&](int s) { return bar0 + 8 * s; };
    auto a_empty = [&](int s) { return bar0 + 8 * (2 + s); };
    auto t_full = [&](int s) { return bar0 + 8 * (4 + s); };
    auto t_empty = [&](int s) { return bar0 + 8 * (6 + s); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: barriers, TMEM, W^T split into hi / lo in the swizzled K-major layout ----
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(a_full(s), NUM_PROD_WARPS * 32);
            mbar_init(a_empty(s), 1);
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

    if (warp >= 5) {
        // ================= producers =================
        const int pt = tid - 5 * 32;                       // 0..127
        const int chunk = pt & 7, rsub = pt >> 3;          // 8 threads cover one 128-byte row segment
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            for (int kb = 0; kb < KBLOCKS; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {              // issue all loads before waiting for the slot
                    const int64_t row = tile * BM + rsub + 16 * i;
                    v[i] = row < M ? __ldcs(reinterpret_cast<const float4 *>(x + row * ldx + kb * BK + chunk * 4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                mbar_wait(a_empty(s), ph ^ 1);             // MMA has drained this stage (passes the first lap)
                unsigned char *hi_b = a_st + s * A_STAGE_BYTES;
                unsigned char *lo_b = hi_b + KB_BYTES;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rsub + 16 * i;
                    uint4 h, l;
                    h.x = to_tf32(v[i].x); l.x = to_tf32(v[i].x - __uint_as_float(h.x));
                    h.y = to_tf32(v[i].y); l.y = to_tf32(v[i].y - __uint_as_float(h.y));
                    h.z = to_tf32(v[i].z); l.z = to_tf32(v[i].z - __uint_as_float(h.z));
                    h.w = to_tf32(v[i].w); l.w = to_tf32(v[i].w - __uint_as_float(h.w));
                    const uint32_t off = sw128_offset(r, chunk);
                    *reinterpret_cast<uint4 *>(hi_b + off) = h;
                    *reinterpret_cast<uint4 *>(lo_b + off) = l;
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
            mbar_wait(t_empty(as), aph ^ 1);               // epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + as * BN;
            for (int kb = 0; kb < KBLOCKS; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(a_full(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t a_hi = smem_u32(a_st + s * A_STAGE_BYTES);
                    const uint32_t a_lo = a_hi + KB_BYTES;
                    const uint32_t b_hi = smem_u32(w_hi + kb * KB_BYTES);
                    const uint32_t b_lo = smem_u32(w_lo + kb * KB_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {     // 32 bytes of K per instruction inside the 128-B atom
                        const uint32_t ko = k * 32;
                        umma_tf32(tmem_d, make_desc(a_lo + ko), make_desc(b_hi + ko), (kb | k) ? 1u : 0u);
                        umma_tf32(tmem_d, make_desc(a_hi + ko), make_desc(b_lo + ko), 1u);
                        umma_tf32(tmem_d, make_desc(a_hi + ko), make_desc(b_hi + ko), 1u);
                    }
                    umma_commit(a_empty(s));               // stage reusable once these MMAs have read it
                    if (kb == KBLOCKS - 1) umma_commit(t_full(as));
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue (warps 0-3: TMEM lanes 32*warp .. +32) =================
        uint32_t tcount = 0;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tcount) {
            const int as = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            mbar_wait(t_full(as), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t row = tile * BM + warp * 32 + lane;
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * BN + c0;
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
                if (row < M) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o;
                        o.x = __uint_as_float(r[j]) + (bias ? __ldg(bias + c0 + j) : 0.f);
                        o.y = __uint_as_float(r[j + 1]) + (bias ? __ldg(bias + c0 + j + 1) : 0.f);
                        o.z = __uint_as_float(r[j + 2]) + (bias ? __ldg(bias + c0 + j + 2) : 0.f);
                        o.w = __uint_as_float(r[j + 3]) + (bias ? __ldg(bias + c0 + j + 3) : 0.f);
                        if (act == 1) {
                            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                        }
                        *reinterpret_cast<float4 *>(out + row * ldo + c0 + j) = o;
                    }
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

}  // namespace

extern "C" int exp_linear_tcgen05_f32(const float *x, int64_t ldx, const float *w, const float *bias, float *out,
                                      int64_t ldo, int64_t M, int act, int sm_count, void *stream) {
    if (M <= 0) return 0;
    cudaError_t e = cudaFuncSetAttribute(linear_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return 1000 + (int)e;
    const int64_t tiles = (M + BM - 1) / BM;
    const unsigned grid = (unsigned)(tiles < sm_count ? tiles : sm_count);
    linear_tcgen05_kernel<<<grid, THREADS, SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(x, ldx, w, bias, out,
                                                                                                ldo, M, act);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : 1000 + (int)e;
}
