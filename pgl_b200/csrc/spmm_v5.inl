// spmm_v5.inl -- included by spmm_stream.cu (inside namespace pglb, after the shared task / fix-up code).
//
// v5 of the wide-row copy-sum aggregation (sum / mean, D % 4 == 0, D <= 128): the same edge-balanced
// tasks, cut-row partials and fix-up as spmm_stream128_kernel, rebuilt around two findings of round 1's
// ncu capture (profiles/r01_ncu_summary.md: 65 warp instructions per edge, issue slots 65 % busy,
// 0.65 IPC x 4 schedulers x 148 SMs x 1.9 GHz / 65 = 11.2 G edges/s -- exactly the measured rate):
// the kernel was ISSUE bound, not byte bound.
//
//  * Gather by the TMA unit (ISSUE 1): one `cp.async.bulk.tensor.2d.tile::gather4` per FOUR feature rows
//    (SASS UTMALDG.2D.GATHER4), issued by one elected lane, completion counted in bytes on an mbarrier per
//    group of GRP rows -- 4 shuffles + ~8 uniform-datapath instructions per four rows instead of
//    (shuffle + IMAD.WIDE + LDGSTS + 3 dead LDS) per row per warp.  ISSUE 0 keeps the per-lane LDGSTS ring
//    (cp.async.cg) so the two mechanisms can be measured against each other on the same consume code.
//  * Consume side written for instruction count: a group that lies inside one CSR row is reduced by a
//    fully unrolled body with compile-time shared-memory offsets (LDS.128 + LDS.32 + 4 FFMA per gathered
//    row); row boundaries are handled per RUN of slots, not per slot; the source norms ride in a small
//    per-warp shared ring filled by one 4-byte cp.async per lane per 32 slots (no per-slot shuffle).
//
// The per-row summation order is unchanged (slot order, __fadd_rn / fmaf), so results are bit-identical to
// spmm_stream128_kernel and, for rows of <= max(T, 1024) slots, to the sequential oracle loop.

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned cnt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait blocks for a hardware-defined interval per attempt; a copy that never lands (a tensor map that
// does not describe the buffer) must become an error the host sees, not a hung device
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void tma_gather4(unsigned dst, const CUtensorMap *tm, int r0, int r1, int r2, int r3,
                                            unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, "
        "%4, %5, %6}], [%7];" ::"r"(dst),
        "l"(tm), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void tma_gather4_hint(unsigned dst, const CUtensorMap *tm, int r0, int r1, int r2, int r3,
                                                 unsigned bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, "
        "{%2, %3, %4, %5, %6}], [%7], %8;" ::"r"(dst),
        "l"(tm), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar), "l"(pol)
        : "memory");
}

// Geometry: GRP rows per group (one mbarrier / one commit group), NG groups per warp ring, LAG = NG - 1 groups
// in flight behind the issue point, W warps per CTA (two CTAs per SM).  Per warp: NG*GRP*512 B of rows +
// 512 B of source norms.
template <int GRP_, int NG_, int W_>
struct GeoV5 {
    static constexpr int kGrp = GRP_, kNg = NG_, kW = W_;
    static constexpr int kLag = NG_ - 1;
    static constexpr int kGpb = 32 / GRP_;
    static constexpr int kWarpBytes = NG_ * GRP_ * 512 + 512;
    static constexpr int kSmem = W_ * kWarpBytes + 128;  // +128: manual alignment of the dynamic base
};

// DYN: persistent warps draw task ids from a device counter (zeroed by task_plan_kernel) instead of the static
// block -> task map: a warp that finishes early takes the next task, so no SM idles while a neighbour's slowest warp
// finishes (p.dyn == 2 hands the ids out from the last task down).  Results do not depend on which warp ran a task.
template <int ISSUE, bool SCALED, bool D128, bool HOT, int GRP, int NG, int W, bool DYN>
__global__ void __launch_bounds__(W * 32, 2) spmm_v5_kernel(const StreamP p, const __grid_constant__ CUtensorMap tm) {
    typedef GeoV5<GRP, NG, W> G_;
    constexpr int LAG = G_::kLag, GPB = G_::kGpb;
    static_assert(GRP % 4 == 0 && 32 % GRP == 0, "groups are whole gather4 quads inside a 32-slot batch");
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) unsigned long long bars[W * NG];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const unsigned smem0 = ((unsigned)__cvta_generic_to_shared(smem_dyn) + 127u) & ~127u;
    const unsigned ring0 = smem0 + wib * G_::kWarpBytes;       // NG*GRP rows
    const unsigned sring = ring0 + NG * GRP * 512;              // 128 floats: norms of 4 column batches
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&bars[wib * NG]);
    if (ISSUE == 1) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < NG; ++s) mbar_init(bar0 + s * 8, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    if (!DYN && (int64_t)blockIdx.x * W + wib >= p.ntasks) return;
    // HOT: bit 31 of the packed ids marks the most frequently gathered sources (pglb_pack_cols); a quad that holds
    // one is fetched with the "keep" policy, a quad of cold rows with evict_first so the long tail cannot flush them
    const uint64_t pol_keep = !HOT ? 0 : (p.hot_mode == 3 ? policy_evict_normal() : policy_evict_last());
    const uint64_t pol_cold = !HOT ? 0 : (p.hot_mode == 2 ? policy_evict_normal() : policy_evict_first());
    // row pitch in shared memory: the TMA box packs rows of D floats; the LDGSTS ring uses 512-B slots
    // (D128: D == 128, every stride is a compile-time constant)
    const unsigned rp = (ISSUE == 1 && !D128) ? (unsigned)p.D * 4u : 512u;
    const unsigned qs = (ISSUE == 1 && !D128) ? ((rp * 4u + 127u) & ~127u) : 2048u;  // stride of a quad of rows
    const unsigned gs = qs * (GRP / 4);                                        // stride of a group
    const bool act = lane * 4 < p.D;
    const unsigned lane_off = act ? lane * 16 : 0;  // inactive lanes (D < 128) alias lane 0's bytes
    const char *xlane = reinterpret_cast<const char *>(p.x) + lane_off;
    const unsigned row_bytes = (unsigned)(p.ldx * 4);

    unsigned gtot = 0;  // groups this warp has pushed through its mbarrier ring in earlier tasks (DYN, ISSUE 1)
#pragma unroll 1
    for (;;) {
    int64_t task;
    if (DYN) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(p.counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);  // also: every lane has left the previous task's rings
        if ((int64_t)t >= p.ntasks) break;
        task = p.dyn == 2 ? p.ntasks - 1 - (int64_t)t : (int64_t)t;
    } else {
        task = (int64_t)blockIdx.x * W + wib;
    }
    const long long t_task0 = p.trace ? global_ns() : 0;
    const int64_t a = ld_ro(p.start + task);
    const int64_t b = ld_ro(p.start + task + 1);
    const int cnt = (int)(b - a);
    int64_t row = ld_ro(p.first_row + task);
    int64_t tail = -1;
    if (cnt > 0) {
        auto rel = [&](int64_t v) -> int {
            const int64_t d = v - a;
            return d < -(1 << 30) ? -(1 << 30) : (d > (1 << 30) ? (1 << 30) : (int)d);
        };
        int beg_rel = rel(ld_ro(p.indptr + row));
        int end_rel = rel(ld_ro(p.indptr + row + 1));
        int nxt_rel = (row + 2 <= p.n_rows) ? rel(ld_ro(p.indptr + row + 2)) : (1 << 30);
        bool head = beg_rel < 0;  // first row started in an earlier task (only rows longer than T)
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

        auto finish_row = [&]() {
            if (head) {
                if (act) *reinterpret_cast<float4 *>(p.partial + (2 * task) * p.dpad + lane * 4) = acc;
                head = false;  // the owner task's fix-up finishes this row
            } else if (act) {
                float4 v = acc;
                const int deg = end_rel - beg_rel;
                if (p.accumulate) {
                    const float4 o = *reinterpret_cast<const float4 *>(p.out + row * p.ldo + lane * 4);
                    v.x = __fadd_rn(o.x, v.x); v.y = __fadd_rn(o.y, v.y);
                    v.z = __fadd_rn(o.z, v.z); v.w = __fadd_rn(o.w, v.w);
                }
                if (p.reduce_op == PGLB_REDUCE_MEAN && deg != 0) {
                    const float c = (float)deg;
                    v.x = __fdiv_rn(v.x, c); v.y = __fdiv_rn(v.y, c);
                    v.z = __fdiv_rn(v.z, c); v.w = __fdiv_rn(v.w, c);
                }
                if (p.scale_dst) {
                    const float sd = __ldg(p.scale_dst + row);
                    v.x = __fmul_rn(v.x, sd); v.y = __fmul_rn(v.y, sd);
                    v.z = __fmul_rn(v.z, sd); v.w = __fmul_rn(v.w, sd);
                }
                __stcs(reinterpret_cast<float4 *>(p.out + row * p.ldo + lane * 4), v);
            }
            ++row;
            beg_rel = end_rel;
            end_rel = nxt_rel;
            nxt_rel = (row + 2 <= p.n_rows) ? rel(ld_ro(p.indptr + row + 2)) : (1 << 30);
            if (end_rel == beg_rel && row < p.n_rows) {
                // run of empty rows (zero-filled by empty_rows_kernel): jump to the row that owns the next slot
                const int64_t pos_abs = a + beg_rel;
                if (pos_abs >= p.E) {
                    row = p.n_rows;
                    end_rel = 1 << 30;
                } else {
                    row = row_of_slot_from(p.indptr, p.n_rows, row, pos_abs);
                    end_rel = rel(ld_ro(p.indptr + row + 1));
                    nxt_rel = (row + 2 <= p.n_rows) ? rel(ld_ro(p.indptr + row + 2)) : (1 << 30);
                }
            }
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
        };

        // column ids of the 32-slot batch being issued (lane j holds slot batch*32 + j) and of the next one.
        // Slots past the task end read id 0: a valid row, gathered and never consumed, so every TMA group
        // delivers exactly GRP rows and the expected byte count is a constant.
        auto load_col = [&](int batch) -> unsigned {
            const int j = batch * 32 + lane;
            if (j >= cnt) return 0u;
            if (p.cols32) return HOT ? __ldcs(p.cols32 + a + j) : (__ldcs(p.cols32 + a + j) & 0x7fffffffu);
            return (unsigned)(p.cols ? ld_stream(p.cols + a + j) : (a + j));
        };
        unsigned col_cur = load_col(0);
        unsigned col_nxt = load_col(1);
        if (SCALED) {
            cp_async4(sring + lane * 4, p.scale_src + (col_cur & 0x7fffffffu));
            if (ISSUE == 1) cp_async_commit();
        }

        const int ngroups = (cnt + GRP - 1) / GRP;
#pragma unroll 1
        for (int g = 0; g < ngroups + LAG; ++g) {
            if (g < ngroups) {
                const int sub = g % GPB;  // position of the group inside its 32-slot column batch
                if (sub == 0 && g > 0) {
                    col_cur = col_nxt;
                    col_nxt = load_col(g / GPB + 1);
                    if (SCALED) {
                        cp_async4(sring + (((g / GPB) & 3) * 32 + lane) * 4, p.scale_src + (col_cur & 0x7fffffffu));
                        if (ISSUE == 1) cp_async_commit();
                    }
                }
                const int s = (int)((gtot + (unsigned)g) % NG);
                if (ISSUE == 1) {
                    __syncwarp();  // every lane has finished reading slot s (consumed NG groups ago)
                    if (lane == 0) mbar_expect_tx(bar0 + s * 8, GRP * rp);
#pragma unroll
                    for (int q = 0; q < GRP / 4; ++q) {
                        const int r0 = __shfl_sync(0xffffffffu, col_cur, sub * GRP + q * 4 + 0);
                        const int r1 = __shfl_sync(0xffffffffu, col_cur, sub * GRP + q * 4 + 1);
                        const int r2 = __shfl_sync(0xffffffffu, col_cur, sub * GRP + q * 4 + 2);
                        const int r3 = __shfl_sync(0xffffffffu, col_cur, sub * GRP + q * 4 + 3);
                        if (HOT) {
                            const uint64_t pol = ((r0 | r1 | r2 | r3) < 0) ? pol_keep : pol_cold;
                            if (lane == 0)
                                tma_gather4_hint(ring0 + s * gs + q * qs, &tm, r0 & 0x7fffffff, r1 & 0x7fffffff,
                                                 r2 & 0x7fffffff, r3 & 0x7fffffff, bar0 + s * 8, pol);
                        } else if (lane == 0) {
                            tma_gather4(ring0 + s * gs + q * qs, &tm, r0, r1, r2, r3, bar0 + s * 8);
                        }
                    }
                } else {
                    const int valid = cnt - g * GRP;
                    const unsigned gaddr = ring0 + s * gs + lane * 16;
#pragma unroll
                    for (int k = 0; k < GRP; ++k) {
                        const unsigned c = __shfl_sync(0xffffffffu, col_cur, sub * GRP + k);
                        if (k < valid) {
                            if (HOT) cp_async16_hint(gaddr + k * 512, xlane + (size_t)(c & 0x7fffffffu) * row_bytes,
                                                     (c >> 31) ? pol_keep : pol_cold);
                            else cp_async16(gaddr + k * 512, xlane + (size_t)c * row_bytes);
                        }
                    }
                }
            }
            if (ISSUE == 0) cp_async_commit();
            if (g >= LAG) {
                const int gc = g - LAG;
                const int s = (int)((gtot + (unsigned)gc) % NG);
                const int base = gc * GRP;
                if (ISSUE == 1) {
                    if (SCALED && (gc % GPB) == 0) {
                        cp_async_wait<0>();  // this batch's norms (the only cp.async traffic of this variant)
                        __syncwarp();
                    }
                    mbar_wait(bar0 + s * 8, ((gtot + (unsigned)gc) / NG) & 1u);
                } else {
                    cp_async_wait<LAG>();
                    if (SCALED && (gc % GPB) == 0) __syncwarp();  // norms were written by other lanes
                }
                int valid = cnt - base;
                valid = valid > GRP ? GRP : valid;
                const unsigned gaddr = ring0 + s * gs + lane_off;
                const unsigned saddr = sring + (base & 127) * 4;
                int k = 0;
                while (true) {
                    int lim = end_rel - base;  // the current row ends before slot `lim` of this group
                    if (lim >= GRP && k == 0 && valid == GRP) {
                        // the whole group belongs to the current row: unrolled, compile-time offsets
#pragma unroll
                        for (int j = 0; j < GRP; ++j) {
                            const float4 v = lds128(gaddr + (j / 4) * qs + (j % 4) * rp);
                            if (SCALED) {
                                const float sc = lds32(saddr + j * 4);
                                acc.x = fmaf(v.x, sc, acc.x); acc.y = fmaf(v.y, sc, acc.y);
                                acc.z = fmaf(v.z, sc, acc.z); acc.w = fmaf(v.w, sc, acc.w);
                            } else {
                                acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
                                acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
                            }
                        }
                        k = GRP;
                    } else {
                        lim = lim < valid ? lim : valid;
                        if (D128 || ISSUE == 0) {
                            // running pointers: LDS.128 + LDS.32 + 4 FFMA + 2 adds + compare/branch per slot
                            unsigned pa = gaddr + k * 512, ps = saddr + k * 4;
                            const unsigned pe = gaddr + lim * 512;
#pragma unroll 1
                            for (; pa != pe; pa += 512, ps += 4) {
                                const float4 v = lds128(pa);
                                if (SCALED) {
                                    const float sc = lds32(ps);
                                    acc.x = fmaf(v.x, sc, acc.x); acc.y = fmaf(v.y, sc, acc.y);
                                    acc.z = fmaf(v.z, sc, acc.z); acc.w = fmaf(v.w, sc, acc.w);
                                } else {
                                    acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
                                    acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
                                }
                            }
                            k = lim > k ? lim : k;
                        } else {
#pragma unroll 1
                            for (; k < lim; ++k) {
                                const float4 v = lds128(gaddr + (k >> 2) * qs + (k & 3) * rp);
                                if (SCALED) {
                                    const float sc = lds32(saddr + k * 4);
                                    acc.x = fmaf(v.x, sc, acc.x); acc.y = fmaf(v.y, sc, acc.y);
                                    acc.z = fmaf(v.z, sc, acc.z); acc.w = fmaf(v.w, sc, acc.w);
                                } else {
                                    acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
                                    acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
                                }
                            }
                        }
                    }
                    if (k >= valid) break;
                    finish_row();  // slot k starts the next row
                }
            }
        }
        // rows that end exactly at b (and, for the last non-empty task, every trailing empty row)
        while (row < p.n_rows && end_rel == cnt) finish_row();
        if (row < p.n_rows && beg_rel < cnt) {
            // the open row (longer than T) continues in the next task(s)
            if (act)
                *reinterpret_cast<float4 *>(p.partial + (head ? (2 * task) : (2 * task + 1)) * p.dpad + lane * 4) = acc;
            if (!head) tail = row;
        }
        if (DYN && ISSUE == 1) gtot += (unsigned)ngroups;
    }
    if (lane == 0) p.tail_row[task] = tail;
    trace_task(p.trace, task, t_task0);
    if (!DYN) break;
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
typedef CUresult (*TmapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TmapEncodeFn tmap_encoder() {
    static std::atomic<void *> fn{nullptr};
    void *f = fn.load(std::memory_order_acquire);
    if (!f) {
        cudaDriverEntryPointQueryResult q;
        void *g = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &g, cudaEnableDefault, &q) == cudaSuccess && g) {
            fn.store(g, std::memory_order_release);
            f = g;
        }
    }
    return reinterpret_cast<TmapEncodeFn>(f);
}

// 2-D map over x[n_src, D] (row stride ldx floats) with a {D, 1} box: what tile::gather4 takes.
static int make_row_map(CUtensorMap *tm, const float *x, int64_t n_src, int64_t D, int64_t ldx) {
    TmapEncodeFn enc = tmap_encoder();
    PGLB_CHECK_ARG(enc != nullptr, PGLB_EINVAL, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)n_src};
    cuuint64_t strides[1] = {(cuuint64_t)ldx * 4};
    cuuint32_t box[2] = {(cuuint32_t)D, 1};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PGLB_CHECK_ARG(r == CUDA_SUCCESS, PGLB_EINVAL, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return PGLB_OK;
}

// PGLB_STREAM_V5: 0 = off (spmm_stream128_kernel), 1 = TMA gather4 (default), 2 = LDGSTS ring with the v5 consume code
static int v5_mode() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PGLB_STREAM_V5");
        v = e ? atoi(e) : 1;
        if (v < 0 || v > 2) v = 1;
    }
    return v;
}
// Ring geometry.  0 = groups of 4, ring of 4, 13 warps per CTA; 1 = groups of 8, ring of 4, 6 warps; 2 = groups of 8,
// ring of 3, 9 warps.  PGLB_V5_GEO pins one; otherwise the average row length decides.  Measured (128-float rows, sum;
// profiles/r02_v5_geo_sweep*.log), geometry 0 vs 2:
//     slots per row      6       10      10 (cfg5)   16      24      16 (10M nodes)   24 (10M)   50 (cfg4, mean)
//     geometry 0, ms    0.96    1.35     8.45       3.23    5.31       16.0            25.5        14.0
//     geometry 2, ms    1.00    1.28     8.98       1.81    2.44        9.8            14.2         7.7
// Up to ~10 slots per row most groups straddle a row boundary and the many-warps geometry holds its own (cfg5: 6 %
// better); from 16 slots per row on the fully unrolled 8-row group is 1.6-2.2x faster, whatever the size of the feature
// matrix.  The switch sits at 12.
static int v5_geo_env() {
    static int v = -2;
    if (v == -2) {
        const char *e = getenv("PGLB_V5_GEO");
        v = e ? atoi(e) : -1;
        if (v < -1 || v > 2) v = -1;
    }
    return v;
}
static int v5_geo(const StreamP &p) {
    const int e = v5_geo_env();
    if (e >= 0) return e;
    return (p.n_rows > 0 && p.E >= 12 * p.n_rows) ? 2 : 0;
}

template <int ISSUE, bool SCALED, bool D128, bool HOT, int GRP, int NG, int W>
static int launch_v5_geo(const StreamP &p, const CUtensorMap &tm, cudaStream_t stream) {
    typedef GeoV5<GRP, NG, W> G_;
    int64_t blocks = (p.ntasks + W - 1) / W;
    PGLB_CHECK_ARG(blocks <= 0x7fffffffLL, PGLB_ESHAPE, "spmm_v5: grid too large");
    bool launched = false;
    if constexpr (ISSUE == 1) {
      if (p.dyn) {
        static std::atomic<unsigned long long> attr_done_dyn{0};
        PGLB_CUDA(ensure_dyn_smem(spmm_v5_kernel<ISSUE, SCALED, D128, HOT, GRP, NG, W, true>, G_::kSmem, attr_done_dyn));
        const int64_t resident = (int64_t)sm_count() * 2;  // __launch_bounds__(W * 32, 2)
        if (blocks > resident) blocks = resident;
        spmm_v5_kernel<ISSUE, SCALED, D128, HOT, GRP, NG, W, true><<<(unsigned)blocks, W * 32, G_::kSmem, stream>>>(p, tm);
        launched = true;
      }
    }
    if (!launched) {
        static std::atomic<unsigned long long> attr_done{0};
        PGLB_CUDA(ensure_dyn_smem(spmm_v5_kernel<ISSUE, SCALED, D128, HOT, GRP, NG, W, false>, G_::kSmem, attr_done));
        spmm_v5_kernel<ISSUE, SCALED, D128, HOT, GRP, NG, W, false><<<(unsigned)blocks, W * 32, G_::kSmem, stream>>>(p, tm);
    }
    PGLB_LAUNCH_CHECK("spmm_v5_kernel");
    const int64_t fblocks = (p.ntasks * 32 + 255) / 256;
    spmm_stream_fixup_kernel<1, 0><<<(unsigned)fblocks, 256, 0, stream>>>(p);
    PGLB_LAUNCH_CHECK("spmm_stream_fixup_kernel");
    return PGLB_OK;
}

template <int ISSUE, bool SCALED, bool D128, bool HOT>
static int launch_v5_h(const StreamP &p, const CUtensorMap &tm, cudaStream_t stream) {
    switch (v5_geo(p)) {
        case 1: return launch_v5_geo<ISSUE, SCALED, D128, HOT, 8, 4, 6>(p, tm, stream);
        case 2: return launch_v5_geo<ISSUE, SCALED, D128, HOT, 8, 3, 9>(p, tm, stream);
        default: return launch_v5_geo<ISSUE, SCALED, D128, HOT, 4, 4, 13>(p, tm, stream);
    }
}

template <int ISSUE, bool SCALED>
static int launch_v5_issue(const StreamP &p, const CUtensorMap &tm, bool hot, cudaStream_t stream) {
    if (p.D == 128)
        return hot ? launch_v5_h<ISSUE, SCALED, true, true>(p, tm, stream) : launch_v5_h<ISSUE, SCALED, true, false>(p, tm, stream);
    return launch_v5_h<ISSUE, SCALED, false, false>(p, tm, stream);  // hints only for the 512-byte rows they were tuned on
}

// true when the v5 kernel can take this call (the caller falls back to spmm_stream128_kernel otherwise)
static bool v5_eligible(const StreamP &p, int64_t n_src, int rk, bool small_ids) {
    if (v5_mode() == 0 || rk != 0 || p.y || !small_ids || p.D % 4 != 0 || p.D > 128) return false;
    if (v5_mode() == 1) {
        // TMA: 16-byte aligned base and row stride, int32 row coordinates
        if ((reinterpret_cast<uintptr_t>(p.x) & 15) || ((p.ldx * 4) & 15) || n_src >= 0x7fffffffLL) return false;
        if (!p.cols && !p.cols32) return false;  // identity columns (segment ops): a plain stream, keep the old kernel
    }
    return true;
}

static int launch_v5(const StreamP &p, int64_t n_src, bool hot, cudaStream_t stream) {
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    const bool scaled = p.scale_src != nullptr;
    if (v5_mode() == 1) {
        const int rc = make_row_map(&tm, p.x, n_src, p.D, p.ldx);
        if (rc) return rc;
        return scaled ? launch_v5_issue<1, true>(p, tm, hot, stream) : launch_v5_issue<1, false>(p, tm, hot, stream);
    }
    return scaled ? launch_v5_issue<0, true>(p, tm, hot, stream) : launch_v5_issue<0, false>(p, tm, hot, stream);
}
