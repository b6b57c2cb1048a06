// spmm_gat5.inl -- included by spmm_stream.cu (inside namespace pglb, after spmm_v5.inl).
//
// Single-pass GAT aggregation, second version (reference pgl/nn/conv.py:333-339: send_uv + LeakyReLU + edge_softmax +
// send_ue_recv(mul, sum) in one kernel):
//     out[d,h,:] = sum_j softmax_j( leaky_relu(attn_src[src_j,h] + attn_dst[d,h]) ) * f[src_j,h,:]
// Round 1's version (spmm_stream128_kernel<YM = 2>) measured 2.5 ms on cfg3 (RMAT 1M / 10M, 8 heads x 16) with the
// 0.5 GB feature matrix sitting in L2: 95 warp instructions per edge at 54 % issue utilisation -- instruction bound,
// nowhere near memory (profiles/r02_ncu_summary.md section 3).  This version applies what worked for the copy-sum kernel:
//   * feature rows AND the source's attention row (H floats) arrive by TMA tile::gather4, four slots per instruction
//     each, counted on the same mbarrier -- no per-slot shuffle / address arithmetic / LDGSTS in any warp;
//   * the online softmax is taken FOUR SLOTS AT A TIME when they belong to one row: one running-max update and one
//     rescale of (l, acc) per quad, then four branch-free exp + FFMA groups -- ~16 instructions per slot instead of a
//     data-dependent branch per slot;
//   * row boundaries fall back to the slot-at-a-time update (same arithmetic as round 1's kernel).
// Tasks, cut-row partials (acc, running max, running sum per lane) and the merge kernel are shared with the old path.
// exp is ex2.approx of a pre-scaled argument (__expf): relative error ~1e-6, inside the path's 1e-4 bar.

template <int GRP_, int NG_, int W_>
struct GeoG5 {
    static constexpr int kGrp = GRP_, kNg = NG_, kW = W_;
    static constexpr int kLag = NG_ - 1;
    static constexpr int kGpb = 32 / GRP_;
};

struct Gat5P {
    StreamP s;        // indptr, cols32 / cols, out, ldo, n_rows, E, D, tasks, partial, partial_ml, tail_row, attn_dst, ldy (= H), head_dim, slope
    unsigned rp;      // feature row pitch in shared memory (D * 4)
    unsigned fq;      // stride of a quad of feature rows (aligned to 128)
    unsigned aq;      // stride of a quad of attention rows (aligned to 128)
    unsigned ap;      // attention row pitch (H * 4)
    int64_t task_mul; // task id = (linear warp id * task_mul) % ntasks, gcd(task_mul, ntasks) == 1
};

template <int GRP, int NG, int W, bool DYN>
__global__ void __launch_bounds__(W * 32, 2) spmm_gat5_kernel(const Gat5P gp, const __grid_constant__ CUtensorMap tmf,
                                                              const __grid_constant__ CUtensorMap tma) {
    typedef GeoG5<GRP, NG, W> G_;
    constexpr int LAG = G_::kLag, GPB = G_::kGpb;
    const StreamP &p = gp.s;
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) unsigned long long bars[W * NG];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    // optional multiplicative permutation of the task ids (task_mul = 1: identity, the default; see launch_gat5)
    const int64_t lin = (int64_t)blockIdx.x * W + wib;
    const unsigned rp = gp.rp, fq = gp.fq, aq = gp.aq, ap = gp.ap;
    const unsigned fgs = fq * (GRP / 4), ags = aq * (GRP / 4);       // group strides
    const unsigned warp_bytes = NG * (fgs + ags);
    const unsigned smem0 = ((unsigned)__cvta_generic_to_shared(smem_dyn) + 127u) & ~127u;
    const unsigned fring = smem0 + wib * warp_bytes;
    const unsigned aring = fring + NG * fgs;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&bars[wib * NG]);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NG; ++s) mbar_init(bar0 + s * 8, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (!DYN && lin >= p.ntasks) return;
    const bool act = lane * 4 < p.D;
    const unsigned lane_off = act ? lane * 16 : 0;
    const int yhead = act ? (lane * 4) / p.head_dim : 0;
    const bool hlead = act && p.lse != nullptr && (lane * 4) % p.head_dim == 0;  // first lane of a head: writes lse
    const float slope = p.slope;

    // DYN: persistent warps draw task ids from the device counter task_plan_kernel zeroed (see spmm_v5_kernel)
    unsigned gtot = 0;  // groups pushed through this warp's mbarrier ring by earlier tasks
#pragma unroll 1
    for (;;) {
    int64_t task;
    if (DYN) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(p.counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if ((int64_t)t >= p.ntasks) break;
        task = p.dyn == 2 ? p.ntasks - 1 - (int64_t)t : (int64_t)t;
    } else {
        task = (lin * gp.task_mul) % p.ntasks;
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
        bool head = beg_rel < 0;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float m_run = -INFINITY, l_run = 0.0f;
        // attn_dst of the current row and, prefetched, of the next one: a row of one or two slots ends before a
        // dependent load issued at its start would be back
        auto load_ad = [&](int64_t r) -> float { return (act && r < p.n_rows) ? __ldg(p.attn_dst + r * p.ldy + yhead) : 0.0f; };
        float ad = load_ad(row);
        float ad_nxt = load_ad(row + 1);

        auto finish_row = [&]() {
            if (head) {
                if (act) *reinterpret_cast<float4 *>(p.partial + (2 * task) * p.dpad + lane * 4) = acc;
                p.partial_ml[(2 * task) * 64 + lane * 2] = m_run;
                p.partial_ml[(2 * task) * 64 + lane * 2 + 1] = l_run;
                head = false;
            } else if (act) {
                float4 v = acc;
                if (end_rel - beg_rel != 0) {
                    const float inv = __frcp_rn(l_run);   // one reciprocal + 4 multiplies (within 1 ulp of the division)
                    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                } else {
                    v = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                __stcs(reinterpret_cast<float4 *>(p.out + row * p.ldo + lane * 4), v);
                if (hlead) p.lse[row * p.ldy + yhead] = m_run + logf(l_run);
            }
            ++row;
            beg_rel = end_rel;
            end_rel = nxt_rel;
            nxt_rel = (row + 2 <= p.n_rows) ? rel(ld_ro(p.indptr + row + 2)) : (1 << 30);
            ad = ad_nxt;
            if (end_rel == beg_rel && row < p.n_rows) {
                const int64_t pos_abs = a + beg_rel;
                if (pos_abs >= p.E) {
                    row = p.n_rows;
                    end_rel = 1 << 30;
                } else {
                    row = row_of_slot_from(p.indptr, p.n_rows, row, pos_abs);
                    end_rel = rel(ld_ro(p.indptr + row + 1));
                    nxt_rel = (row + 2 <= p.n_rows) ? rel(ld_ro(p.indptr + row + 2)) : (1 << 30);
                }
                ad = load_ad(row);
            }
            ad_nxt = load_ad(row + 1);
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
            m_run = -INFINITY;
            l_run = 0.0f;
        };

        auto load_col = [&](int batch) -> unsigned {
            const int j = batch * 32 + lane;
            if (j >= cnt) return 0u;
            if (p.cols32) return __ldcs(p.cols32 + a + j) & 0x7fffffffu;
            return (unsigned)ld_stream(p.cols + a + j);
        };
        unsigned col_cur = load_col(0);
        unsigned col_nxt = load_col(1);

        // one slot, the general (branchy) update: exact same arithmetic as round 1's kernel
        auto one_slot = [&](unsigned faddr, unsigned aaddr) {
            const float4 v = lds128(faddr);
            float lg = lds32(aaddr) + ad;
            lg = lg >= 0.0f ? lg : lg * slope;
            if (lg <= m_run) {
                const float pe = __expf(lg - m_run);
                l_run += pe;
                acc.x = fmaf(pe, v.x, acc.x); acc.y = fmaf(pe, v.y, acc.y);
                acc.z = fmaf(pe, v.z, acc.z); acc.w = fmaf(pe, v.w, acc.w);
            } else {
                const float sc = __expf(m_run - lg);   // exp(-inf) = 0 at a row start
                l_run = fmaf(l_run, sc, 1.0f);
                acc.x = fmaf(acc.x, sc, v.x); acc.y = fmaf(acc.y, sc, v.y);
                acc.z = fmaf(acc.z, sc, v.z); acc.w = fmaf(acc.w, sc, v.w);
                m_run = lg;
            }
        };

        const int ngroups = (cnt + GRP - 1) / GRP;
#pragma unroll 1
        for (int g = 0; g < ngroups + LAG; ++g) {
            if (g < ngroups) {
                const int sub = g % GPB;
                if (sub == 0 && g > 0) {
                    col_cur = col_nxt;
                    col_nxt = load_col(g / GPB + 1);
                }
                const int s = (int)((gtot + (unsigned)g) % NG);
                __syncwarp();  // every lane has finished reading slot s (consumed NG groups ago)
                if (lane == 0) mbar_expect_tx(bar0 + s * 8, GRP * (rp + ap));
#pragma unroll
                for (int q = 0; q < GRP / 4; ++q) {
                    const int r0 = __shfl_sync(0xffffffffu, col_cur, sub * GRP + q * 4 + 0);
                    const int r1 = __shfl_sync(0xffffffffu, col_cur, sub * GRP + q * 4 + 1);
                    const int r2 = __shfl_sync(0xffffffffu, col_cur, sub * GRP + q * 4 + 2);
                    const int r3 = __shfl_sync(0xffffffffu, col_cur, sub * GRP + q * 4 + 3);
                    if (lane == 0) {
                        tma_gather4(fring + s * fgs + q * fq, &tmf, r0, r1, r2, r3, bar0 + s * 8);
                        tma_gather4(aring + s * ags + q * aq, &tma, r0, r1, r2, r3, bar0 + s * 8);
                    }
                }
            }
            if (g >= LAG) {
                const int gc = g - LAG;
                const int s = (int)((gtot + (unsigned)gc) % NG);
                const int base = gc * GRP;
                mbar_wait(bar0 + s * 8, ((gtot + (unsigned)gc) / NG) & 1u);
                int valid = cnt - base;
                valid = valid > GRP ? GRP : valid;
                const unsigned fg = fring + s * fgs + lane_off;
                const unsigned ag = aring + s * ags + yhead * 4;
                int k = 0;
                while (true) {
                    int lim = end_rel - base;
                    lim = lim < valid ? lim : valid;
                    // whole quads inside the current row: one max / rescale per four slots
                    while ((k & 3) == 0 && lim - k >= 4) {
                        const unsigned fa = fg + (k >> 2) * fq;
                        const unsigned aa = ag + (k >> 2) * aq;
                        float lg0 = lds32(aa) + ad, lg1 = lds32(aa + ap) + ad, lg2 = lds32(aa + 2 * ap) + ad,
                              lg3 = lds32(aa + 3 * ap) + ad;
                        lg0 = fmaxf(lg0, lg0 * slope); lg1 = fmaxf(lg1, lg1 * slope);   // leaky relu, 0 <= slope <= 1
                        lg2 = fmaxf(lg2, lg2 * slope); lg3 = fmaxf(lg3, lg3 * slope);
                        const float m_new = fmaxf(fmaxf(fmaxf(lg0, lg1), fmaxf(lg2, lg3)), m_run);
                        const float sc = __expf(m_run - m_new);      // 0 at a row start (m_run = -inf), 1 when the max stands
                        m_run = m_new;
                        const float p0 = __expf(lg0 - m_new), p1 = __expf(lg1 - m_new), p2 = __expf(lg2 - m_new),
                                    p3 = __expf(lg3 - m_new);
                        l_run = fmaf(l_run, sc, (p0 + p1) + (p2 + p3));
                        const float4 v0 = lds128(fa), v1 = lds128(fa + rp), v2 = lds128(fa + 2 * rp), v3 = lds128(fa + 3 * rp);
                        acc.x = fmaf(p3, v3.x, fmaf(p2, v2.x, fmaf(p1, v1.x, fmaf(p0, v0.x, acc.x * sc))));
                        acc.y = fmaf(p3, v3.y, fmaf(p2, v2.y, fmaf(p1, v1.y, fmaf(p0, v0.y, acc.y * sc))));
                        acc.z = fmaf(p3, v3.z, fmaf(p2, v2.z, fmaf(p1, v1.z, fmaf(p0, v0.z, acc.z * sc))));
                        acc.w = fmaf(p3, v3.w, fmaf(p2, v2.w, fmaf(p1, v1.w, fmaf(p0, v0.w, acc.w * sc))));
                        k += 4;
                    }
#pragma unroll 1
                    for (; k < lim; ++k) one_slot(fg + (k >> 2) * fq + (k & 3) * rp, ag + (k >> 2) * aq + (k & 3) * ap);
                    if (k >= valid) break;
                    finish_row();
                }
            }
        }
        while (row < p.n_rows && end_rel == cnt) finish_row();
        if (row < p.n_rows && beg_rel < cnt) {
            const int64_t sl = head ? (2 * task) : (2 * task + 1);
            if (act) *reinterpret_cast<float4 *>(p.partial + sl * p.dpad + lane * 4) = acc;
            p.partial_ml[sl * 64 + lane * 2] = m_run;
            p.partial_ml[sl * 64 + lane * 2 + 1] = l_run;
            if (!head) tail = row;
        }
        if (DYN) gtot += (unsigned)ngroups;
    }
    if (lane == 0) p.tail_row[task] = tail;
    trace_task(p.trace, task, t_task0);
    if (!DYN) break;
    }
}

// PGLB_GAT_V5 = 0 keeps round 1's fused kernel
static int gat5_mode() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PGLB_GAT_V5");
        v = (e && atoi(e) == 0) ? 0 : 1;
    }
    return v;
}

// PGLB_GAT_GEO: 0 = groups of 8, ring of 3, 8 warps per CTA (default); 1 = groups of 4, ring of 4, 12 warps
static int gat5_geo() {
    static int geo = -1;
    if (geo < 0) {
        const char *e = getenv("PGLB_GAT_GEO");
        geo = (e && atoi(e) == 1) ? 1 : 0;
    }
    return geo;
}
constexpr int kGat5SmemBudget = 112 * 1024;   // two CTAs per SM
// dynamic shared memory of a launch: per warp NG groups of GRP feature rows + GRP attention rows (quads padded to 128 B)
static int64_t gat5_smem(int64_t D, int64_t H) {
    const int64_t fq = (D * 16 + 127) & ~127LL, aq = (H * 16 + 127) & ~127LL;
    const int64_t grp = gat5_geo() == 1 ? 4 : 8, ng = gat5_geo() == 1 ? 4 : 3, w = gat5_geo() == 1 ? 12 : 8;
    return w * ng * (fq + aq) * (grp / 4) + 128;
}

static bool gat5_eligible(const StreamP &p, const float *f, int64_t ldf, const float *attn_src, int64_t H, int64_t n_src) {
    if (!gat5_mode()) return false;
    if (p.D % 4 != 0 || p.D > 128 || H % 4 != 0 || H > 32) return false;
    if (gat5_smem(p.D, H) > kGat5SmemBudget) return false;   // e.g. 32 heads x 4: the attention rows no longer fit the rings
    if ((reinterpret_cast<uintptr_t>(f) & 15) || ((ldf * 4) & 15) || (reinterpret_cast<uintptr_t>(attn_src) & 15)) return false;
    if (n_src >= 0x7fffffffLL) return false;
    if (p.slope < 0.0f || p.slope > 1.0f) return false;   // the quad path writes leaky relu as max(x, slope * x)
    return true;
}

template <int GRP, int NG, int W>
static int launch_gat5_geo(const Gat5P &gp, const CUtensorMap &tmf, const CUtensorMap &tma, cudaStream_t stream) {
    const int smem = W * NG * (int)((gp.fq + gp.aq) * (GRP / 4)) + 128;
    PGLB_CHECK_ARG(smem <= 112 * 1024, PGLB_ESHAPE, "spmm_gat5: shared memory budget exceeded");
    const StreamP &p = gp.s;
    int64_t blocks = (p.ntasks + W - 1) / W;
    PGLB_CHECK_ARG(blocks <= 0x7fffffffLL, PGLB_ESHAPE, "spmm_gat5: grid too large");
    if (p.dyn) {
        static std::atomic<unsigned long long> attr_done_dyn{0};
        PGLB_CUDA(ensure_dyn_smem(spmm_gat5_kernel<GRP, NG, W, true>, 112 * 1024, attr_done_dyn));
        const int64_t resident = (int64_t)sm_count() * 2;  // __launch_bounds__(W * 32, 2), <= 112 KB of rings per CTA
        if (blocks > resident) blocks = resident;
        spmm_gat5_kernel<GRP, NG, W, true><<<(unsigned)blocks, W * 32, smem, stream>>>(gp, tmf, tma);
    } else {
        static std::atomic<unsigned long long> attr_done{0};
        PGLB_CUDA(ensure_dyn_smem(spmm_gat5_kernel<GRP, NG, W, false>, 112 * 1024, attr_done));
        spmm_gat5_kernel<GRP, NG, W, false><<<(unsigned)blocks, W * 32, smem, stream>>>(gp, tmf, tma);
    }
    PGLB_LAUNCH_CHECK("spmm_gat5_kernel");
    const int64_t fblocks = (p.ntasks * 32 + 255) / 256;
    spmm_stream_fixup_gat_kernel<<<(unsigned)fblocks, 256, 0, stream>>>(p);
    PGLB_LAUNCH_CHECK("spmm_stream_fixup_gat_kernel");
    return PGLB_OK;
}

static int launch_gat5(const StreamP &p, const float *attn_src, int64_t H, int64_t n_src, cudaStream_t stream) {
    Gat5P gp;
    gp.s = p;
    gp.rp = (unsigned)p.D * 4u;
    gp.fq = (gp.rp * 4u + 127u) & ~127u;
    gp.ap = (unsigned)H * 4u;
    gp.aq = (gp.ap * 4u + 127u) & ~127u;
    // PGLB_GAT_TASK_PERM=1 scatters the task ids with a multiplicative permutation (ncu showed the SMs busy only 66 % of
    // the launch: RMAT puts the hub rows first and the one-edge rows last).  Measured on cfg3: 2.17 ms without, 2.60 ms
    // with -- neighbouring tasks share gathered sources in L2 and the permutation throws that away.  Off by default.
    {
        static int perm = -1;
        if (perm < 0) {
            const char *e = getenv("PGLB_GAT_TASK_PERM");
            perm = (e && atoi(e) == 1) ? 1 : 0;
        }
        auto gcd = [](int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; };
        int64_t m = 9973;
        while (m > 1 && gcd(m, p.ntasks) != 1) --m;
        gp.task_mul = (perm && p.ntasks > 64) ? m : 1;
    }
    CUtensorMap tmf, tma;
    memset(&tmf, 0, sizeof(tmf));
    memset(&tma, 0, sizeof(tma));
    int rc = make_row_map(&tmf, p.x, n_src, p.D, p.ldx);
    if (rc) return rc;
    rc = make_row_map(&tma, attn_src, n_src, H, H);
    if (rc) return rc;
    return gat5_geo() == 1 ? launch_gat5_geo<4, 4, 12>(gp, tmf, tma, stream) : launch_gat5_geo<8, 3, 8>(gp, tmf, tma, stream);
}
