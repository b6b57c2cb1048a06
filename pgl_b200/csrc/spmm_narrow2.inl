// spmm_narrow2.inl -- included by spmm_stream.cu (inside namespace pglb).
//
// Narrow-row copy-sum aggregation (sum / mean, D = 4..64 floats, D % 4 == 0): the kernel behind the
// column-sharded multi-GPU layout (every GPU holds the whole CSR and D/R columns of every row, so rows are
// 64 / 128 / 256 bytes wide at 8 / 4 / 2 GPUs).  It replaces spmm_narrow_kernel, which kept the wide kernel's
// "one slot after the other" order and therefore needed a cross-lane reduction at every row end (measured in
// round 2's first GPU call: 4.47 ms for cfg5 at D = 16, 0.27 of the HBM roofline).
//
// Layout of the work.  LPR = D/4 rounded up to 4 / 8 / 16 lanes hold one feature row (one float4 each);
// a warp therefore works on EPW = 32 / LPR rows at once.  The unit of work is a CHUNK of EPW * 32 consecutive
// CSR slots: sub-warp `sub` owns the 32 CONSECUTIVE slots [chunk + 32 sub, chunk + 32 sub + 32) and walks them
// in order, so a row's slots are summed sequentially in registers by one sub-warp and a row end costs a
// predicated 16-byte store per lane, not a shuffle tree.  Rows that cross from one sub-warp's range into the
// next are stitched once per chunk by a serial carry over the EPW sub-warps (5 shuffles per sub-warp); rows that
// cross a task boundary (a task = T consecutive slots, one warp) leave partials for the shared fix-up kernel.
//
// What the kernel reads per slot is one 32-bit word of the PLAN (built once per graph, cached next to the packed
// column ids): bits 0..29 the source id, bit 30 "this slot starts a new row".  Two small side arrays turn those
// flags into row numbers without touching indptr: nz_row[k] = id of the k-th non-empty row, blk_k[b] = k of the
// row that owns slot 32 b - 1 (the row a sub-warp is in before it sees its first flag).
//
// Every lane keeps 8 gathered rows (and their source norms) in flight in registers; with ~24 resident warps per
// SM that is ~100 KB of outstanding reads per SM, and the instruction count is ~3.5 warp instructions per edge
// at D = 16 -- the kernel is bound by the memory system, not by issue.
//
// Summation order: slot order inside a sub-warp's range, ranges of one row added left to right (carry + head).
// Deterministic; equal to the sequential loop for rows that live inside one 32-slot range, equal to rounding
// (<= 1e-6 relative) otherwise.

struct NarrowP {
    const uint32_t *plan;    // [E] id | head << 30
    const int32_t *nz_row;   // [K + 2] ids of the non-empty rows, padded with the last one
    const int32_t *blk_k;    // [ceil(E / 32)] index into nz_row of the row owning slot 32 b - 1 (-1 for b = 0)
    const int64_t *indptr;   // mean only
    const float *x;
    int64_t ldx;
    float *out;
    int64_t ldo;
    int64_t E;
    int D;
    int mean;
    const float *scale_src;  // SC 1: gathered per slot (scale_src[id])
    const float *sval;       // SC 2: scale_src[id] cached per SLOT (sval[j], streamed like the plan)
    const float *scale_dst;
    int hot_mode;            // as StreamP: which L2 policies bit 31 of the plan selects
    int64_t T;       // slots per task, a multiple of the chunk size
    int64_t ntasks;
    float *partial;  // [2 * ntasks, dpad]
    int64_t dpad;
    int64_t *tail_row;  // [ntasks]
};

constexpr int kNarrowWarps = 8;

__device__ __forceinline__ float4 ldg128_na(const void *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg128_na_hint(const void *p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}

// SC: 0 no source scale, 1 gathered (scale_src[id]), 2 cached per slot (sval[j]).  HOT: bit 31 of the plan selects an
// L2 policy per row (keep the most gathered sources, stream the tail).  U: gathers in flight per lane (8 or 4; fewer
// registers = more resident warps).
template <int LPR, int SC, bool MEAN, bool HOT, int U>
__global__ void __launch_bounds__(kNarrowWarps * 32, (U == 4 ? 3 : 2)) spmm_narrow2_kernel(const NarrowP p) {
    constexpr int EPW = 32 / LPR;      // sub-warps (rows in flight) per warp
    constexpr int CPL = 32 / LPR;      // plan words each lane loads per 32-slot range
    constexpr bool SCALED = SC != 0;
    static_assert(U == 8 || U == 4, "8 or 4 gathers in flight per lane");
    const uint64_t pol_keep = !HOT ? 0 : (p.hot_mode == 3 ? policy_evict_normal() : policy_evict_last());
    const uint64_t pol_cold = !HOT ? 0 : (p.hot_mode == 2 ? policy_evict_normal() : policy_evict_first());
    static_assert(LPR == 4 || LPR == 8 || LPR == 16, "4, 8 or 16 lanes per row");
    const int lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * kNarrowWarps + (threadIdx.x >> 5);
    if (task >= p.ntasks) return;
    const int sub = lane / LPR, li = lane % LPR;
    const bool act = li * 4 < p.D;
    const char *xlane = reinterpret_cast<const char *>(p.x) + (act ? li * 16 : 0);
    const unsigned row_bytes = (unsigned)(p.ldx * 4);
    const int64_t t_beg = task * p.T;
    const int64_t t_end = (t_beg + p.T < p.E) ? t_beg + p.T : p.E;

    // the row that is open across sub-warp / chunk boundaries: its partial sum so far (meaningful in the lanes
    // of sub-warp 0 between chunks), and whether it began before this task (then it goes to the partial buffer)
    float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
    int carry_mode = 1;      // 1: the open row began in an earlier task ("task head")
    int carry_row = -1;      // row id of the open row (valid once carry_mode == 0)

    auto write_out = [&](int row, float sd, float4 v) {
        if (MEAN) {
            const float c = (float)(ld_ro(p.indptr + row + 1) - ld_ro(p.indptr + row));
            v.x = __fdiv_rn(v.x, c); v.y = __fdiv_rn(v.y, c); v.z = __fdiv_rn(v.z, c); v.w = __fdiv_rn(v.w, c);
        }
        if (p.scale_dst) {
            v.x = __fmul_rn(v.x, sd); v.y = __fmul_rn(v.y, sd); v.z = __fmul_rn(v.z, sd); v.w = __fmul_rn(v.w, sd);
        }
        if (act) __stcs(reinterpret_cast<float4 *>(p.out + (int64_t)row * p.ldo + li * 4), v);
    };

#pragma unroll 1
    for (int64_t chunk = t_beg; chunk < t_end; chunk += EPW * 32) {
        const int64_t rb = chunk + sub * 32;               // my sub-warp's range
        int nvalid = (int)((t_end - rb) < 32 ? (t_end - rb) : 32);
        if (nvalid < 0) nvalid = 0;
        // plan words of my range: lane li holds slots [li * CPL, li * CPL + CPL)
        unsigned creg[CPL];
        float sreg[SC == 2 ? CPL : 1];
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            const int j = li * CPL + r;
            creg[r] = (j < nvalid) ? __ldcs(p.plan + rb + j) : 0u;
            if (SC == 2) sreg[r] = (j < nvalid) ? __ldcs(p.sval + rb + j) : 0.0f;
        }
        // rows: the one I am in before my first flag, and the next non-empty one
        int k = (nvalid > 0) ? __ldg(p.blk_k + (rb >> 5)) : -1;
        int row_cur = (k >= 0) ? __ldg(p.nz_row + k) : 0;
        int row_nxt = __ldg(p.nz_row + k + 1);
        float sd_cur = p.scale_dst ? __ldg(p.scale_dst + row_cur) : 1.0f;
        const int row_head = row_cur;
        const float sd_head = sd_cur;

        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 head_acc = make_float4(0.f, 0.f, 0.f, 0.f);
        bool flagged = false;

#pragma unroll 1
        for (int g = 0; g < 32 / U; ++g) {
            unsigned c[U];
            float4 v[U];
            float s[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = g * U + u;                        // step inside the range
                // which lane of my sub-warp holds the plan word of step j, and in which register
                const int wl = (U >= CPL) ? (g * (U / CPL) + u / CPL) : ((g * U + u) / CPL);
                const int wr = (U >= CPL) ? (u % CPL) : ((g % (CPL / U)) * U + u);
                unsigned cw;
                float sw = 0.0f;
                if (U >= CPL) {
                    cw = __shfl_sync(0xffffffffu, creg[wr], sub * LPR + wl);
                    if (SC == 2) sw = __shfl_sync(0xffffffffu, sreg[SC == 2 ? wr : 0], sub * LPR + wl);
                } else {
                    // U < CPL: the register index depends on g -- select it without dynamic register indexing
                    unsigned pick = 0u;
                    float spick = 0.0f;
#pragma unroll
                    for (int q = 0; q < CPL / U; ++q)
                        if ((g % (CPL / U)) == q) {
                            pick = creg[q * U + u];
                            if (SC == 2) spick = sreg[SC == 2 ? q * U + u : 0];
                        }
                    cw = __shfl_sync(0xffffffffu, pick, sub * LPR + wl);
                    if (SC == 2) sw = __shfl_sync(0xffffffffu, spick, sub * LPR + wl);
                    (void)wr;
                }
                c[u] = cw;
                const unsigned id = c[u] & 0x3fffffffu;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                s[u] = sw;
                if (j < nvalid) {
                    const void *src = xlane + (size_t)id * row_bytes;
                    v[u] = HOT ? ldg128_na_hint(src, (c[u] >> 31) ? pol_keep : pol_cold) : ldg128_na(src);
                    if (SC == 1) s[u] = __ldg(p.scale_src + id);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (c[u] & 0x40000000u) {   // this slot starts a new row: the row so far is complete
                    if (!flagged) {
                        head_acc = acc;     // ... but it began before my range: stitched after the chunk
                        flagged = true;
                    } else {
                        write_out(row_cur, sd_cur, acc);
                    }
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    ++k;
                    row_cur = row_nxt;
                    sd_cur = p.scale_dst ? __ldg(p.scale_dst + row_cur) : 1.0f;
                    row_nxt = __ldg(p.nz_row + k + 1);
                }
                if (SCALED) {
                    acc.x = fmaf(v[u].x, s[u], acc.x); acc.y = fmaf(v[u].y, s[u], acc.y);
                    acc.z = fmaf(v[u].z, s[u], acc.z); acc.w = fmaf(v[u].w, s[u], acc.w);
                } else {
                    acc.x = __fadd_rn(acc.x, v[u].x); acc.y = __fadd_rn(acc.y, v[u].y);
                    acc.z = __fadd_rn(acc.z, v[u].z); acc.w = __fadd_rn(acc.w, v[u].w);
                }
            }
        }
        // my range as (head piece, [complete rows already written], tail piece)
        if (!flagged) head_acc = acc;       // one row covers the whole range: it is all "head"
        const float4 tail_acc = acc;
        const int tail_row = row_cur;

        // stitch the EPW ranges of this chunk, left to right
        float4 cin = carry;
        int cmode = carry_mode;
        int crow = carry_row;
#pragma unroll
        for (int sidx = 0; sidx < EPW; ++sidx) {
            float4 cout = cin;
            int cmode_o = cmode, crow_o = crow;
            if (sub == sidx) {
                float4 tot;
                tot.x = __fadd_rn(cin.x, head_acc.x); tot.y = __fadd_rn(cin.y, head_acc.y);
                tot.z = __fadd_rn(cin.z, head_acc.z); tot.w = __fadd_rn(cin.w, head_acc.w);
                if (flagged) {
                    // the open row ends inside my range
                    if (cmode) {
                        if (act) *reinterpret_cast<float4 *>(p.partial + (2 * task) * p.dpad + li * 4) = tot;
                    } else {
                        write_out(row_head, sd_head, tot);
                    }
                    cout = tail_acc;
                    cmode_o = 0;
                    crow_o = tail_row;
                } else {
                    cout = tot;
                }
            }
            // hand the open row to the next sub-warp (the last one hands it back to sub-warp 0 for the next chunk)
            const int from = (sidx * LPR + li) & 31;
            cin.x = __shfl_sync(0xffffffffu, cout.x, from); cin.y = __shfl_sync(0xffffffffu, cout.y, from);
            cin.z = __shfl_sync(0xffffffffu, cout.z, from); cin.w = __shfl_sync(0xffffffffu, cout.w, from);
            cmode = __shfl_sync(0xffffffffu, cmode_o, from);
            crow = __shfl_sync(0xffffffffu, crow_o, from);
        }
        carry = cin;
        carry_mode = cmode;
        carry_row = crow;
    }
    // the row that is still open at the end of the task
    int64_t tail = -1;
    if (sub == 0) {
        if (carry_mode) {
            if (act) *reinterpret_cast<float4 *>(p.partial + (2 * task) * p.dpad + li * 4) = carry;
        } else {
            if (act) *reinterpret_cast<float4 *>(p.partial + (2 * task + 1) * p.dpad + li * 4) = carry;
            tail = carry_row;
        }
    }
    if (lane == 0) p.tail_row[task] = tail;
}
