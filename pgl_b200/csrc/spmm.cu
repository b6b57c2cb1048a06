// spmm.cu -- node-parallel CSR aggregation for sm_100a (K1 send_u_recv, K2 send_ue_recv,
// K4 segment_* of SURVEY.md section 2.2).
//
// Design (DESIGN.md section "spmm_csr"):
//  * a "group" of G lanes (G = 4..32, one float4 or float per lane per iteration) owns a
//    task of `rpg` consecutive dst rows; the rows' slots are one contiguous range of the CSR,
//    so column indices are fetched G at a time with one coalesced streaming load and
//    broadcast by shuffle; U independent feature-row loads (U*ITERS = 8 x 16 B per lane)
//    are issued before the first add, so every group keeps >= 4 KB of gathers in flight
//    regardless of how short the individual rows are.
//  * accumulation is strictly sequential in slot order (= ascending edge id inside a row,
//    the order of the reference's CPU loop), with __fadd_rn/__fmul_rn so that no FMA
//    contraction changes the rounding: results are bit-identical to the sequential
//    restatement for every row below the hub threshold.
//  * rows longer than HUB_T slots ("hubs": ~4k rows holding ~18 % of the edges of the
//    10M/100M power-law graph) are not processed by the row's owner: they are appended to a
//    device-side list, cut into CH-slot chunks by a one-block planning kernel, and the SAME
//    kernel is re-launched twice over "virtual rows" (chunk -> partial, partials -> row).
//    No atomics on feature data, fixed chunking => deterministic run to run.
#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace pglb {

constexpr int64_t HUB_T = 1024;  // rows with more slots than this take the hub path
constexpr int64_t HUB_CH = 1024; // slots per hub chunk
constexpr int kThreads = 256;

struct SpmmP {
    const int64_t *rbeg;  // row r covers slots [rbeg[r], rend[r])
    const int64_t *rend;
    const int64_t *cols;  // nullable: identity
    const int64_t *eid;   // nullable: identity (only read when y != nullptr)
    const float *x;
    int64_t ldx;
    const float *y;
    int64_t ldy;
    int y_bcast;
    float *out;
    int64_t ldo;
    int64_t n_rows;
    const int64_t *n_rows_dev;  // nullable: overrides n_rows (hub passes)
    int D;
    int head_dim;
    int msg_op;
    int reduce_op;
    const float *scale_src;
    const float *scale_dst;
    const int64_t *row_map;      // nullable: out row = row_map[row]
    const int64_t *mean_indptr;  // nullable: MEAN divides by mean_indptr[o+1]-mean_indptr[o]
    int rpg;                     // rows per group task
    int contig;                  // 1: rows of a task are contiguous in slot space
    int64_t hub_threshold;
    unsigned long long *hub_count;
    int64_t *hub_rows;
    int accumulate;  // SUM only: add the previous contents of out before the row epilogue
};

template <int VEC>
struct VecIO;
template <>
struct VecIO<4> {
    static __device__ __forceinline__ void ld(const float *p, float (&v)[4]) {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float *p, const float (&v)[4]) {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    }
};
template <>
struct VecIO<1> {
    static __device__ __forceinline__ void ld(const float *p, float (&v)[1]) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void st(float *p, const float (&v)[1]) { __stcs(p, v[0]); }
};

__device__ __forceinline__ float apply_msg(int op, float a, float b) {
    switch (op) {
        case PGLB_MSG_ADD: return __fadd_rn(a, b);
        case PGLB_MSG_SUB: return __fsub_rn(a, b);
        case PGLB_MSG_MUL: return __fmul_rn(a, b);
        case PGLB_MSG_DIV: return __fdiv_rn(a, b);
        default: return a;
    }
}

// MODE 0: msg = x[col] (* scale_src);  MODE 1: msg = (x[col] (* scale_src)) op y[eid]
// RK 0: sum / mean;  RK 1: max / min
template <int VEC, int G, int ITERS, int MODE, int RK>
__global__ void __launch_bounds__(kThreads) spmm_csr_kernel(const SpmmP p) {
    constexpr int U = (8 / ITERS) < G ? (8 / ITERS) : G;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const int64_t groups_per_block = kThreads / G;
    const int64_t gid = (int64_t)blockIdx.x * groups_per_block + threadIdx.x / G;
    const int64_t ngroups = (int64_t)gridDim.x * groups_per_block;
    const int64_t n_rows = p.n_rows_dev ? *p.n_rows_dev : p.n_rows;
    const int col_tile = blockIdx.y * (G * VEC * ITERS);
    const bool is_max = (p.reduce_op == PGLB_REDUCE_MAX);
    const float ident = (RK == 0) ? 0.0f : (is_max ? -INFINITY : INFINITY);

    int col[ITERS];
    bool act[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        col[it] = col_tile + (it * G + gl) * VEC;
        act[it] = col[it] < p.D;
    }

    for (int64_t task = gid; task * p.rpg < n_rows; task += ngroups) {
        int64_t row = task * p.rpg;
        int64_t row_end = row + p.rpg;
        if (row_end > n_rows) row_end = n_rows;
        const int64_t task_end = p.contig ? ld_ro(p.rend + row_end - 1) : 0;

        float acc[ITERS][VEC];
        int64_t pos = 0, cur_beg = 0, cur_end = 0;
        bool hub = false;

        auto open_row = [&]() {
            cur_beg = ld_ro(p.rbeg + row);
            cur_end = ld_ro(p.rend + row);
            pos = cur_beg;
            hub = (cur_end - cur_beg) > p.hub_threshold;
            if (hub && gl == 0 && blockIdx.y == 0) {
                unsigned long long s = atomicAdd(p.hub_count, 1ull);
                p.hub_rows[s] = row;
            }
#pragma unroll
            for (int it = 0; it < ITERS; ++it)
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[it][j] = ident;
        };
        auto close_row = [&]() {
            if (hub) return;
            const int64_t orow = p.row_map ? ld_ro(p.row_map + row) : row;
            const int64_t deg = cur_end - cur_beg;
            float inv_cnt = 1.0f;
            if (p.reduce_op == PGLB_REDUCE_MEAN) {
                int64_t cnt = p.mean_indptr
                                  ? (ld_ro(p.mean_indptr + orow + 1) - ld_ro(p.mean_indptr + orow))
                                  : deg;
                inv_cnt = (float)cnt;
            }
            const float sd = p.scale_dst ? __ldg(p.scale_dst + orow) : 1.0f;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                if (!act[it]) continue;
                float v[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    float a = acc[it][j];
                    if (deg == 0) a = 0.0f;
                    if (p.accumulate) a = __fadd_rn(p.out[orow * p.ldo + col[it] + j], a);
                    if (p.reduce_op == PGLB_REDUCE_MEAN && deg != 0) a = __fdiv_rn(a, inv_cnt);
                    if (p.scale_dst) a = __fmul_rn(a, sd);
                    v[j] = a;
                }
                VecIO<VEC>::st(p.out + orow * p.ldo + col[it], v);
            }
        };

        open_row();
        for (;;) {
            if (hub || pos == cur_end) {
                close_row();
                ++row;
                if (row == row_end) break;
                open_row();
                continue;
            }
            const int64_t lim = p.contig ? task_end : cur_end;
            const int64_t rem = lim - pos;
            const int nb = rem < (int64_t)G ? (int)rem : G;
            int64_t my_col = 0, my_e = 0;
            if (gl < nb) {
                my_col = p.cols ? ld_stream(p.cols + pos + gl) : (pos + gl);
                if (MODE == 1) my_e = p.eid ? ld_stream(p.eid + pos + gl) : (pos + gl);
            }
            bool abandon = false;
            for (int k0 = 0; k0 < nb && !abandon; k0 += U) {
                float xv[U][ITERS][VEC];
                float yv[MODE == 1 ? U : 1][ITERS][VEC];
                float sv[U];
                // ---- phase A: issue every load of this sub-batch ----
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int64_t c = __shfl_sync(gmask, my_col, k0 + k, G);
                    int64_t e = 0;
                    if (MODE == 1) e = __shfl_sync(gmask, my_e, k0 + k, G);
                    if (k0 + k < nb) {
                        const float *xr = p.x + c * p.ldx;
#pragma unroll
                        for (int it = 0; it < ITERS; ++it) {
                            if (act[it]) VecIO<VEC>::ld(xr + col[it], xv[k][it]);
                        }
                        if (p.scale_src) sv[k] = __ldg(p.scale_src + c);
                        if (MODE == 1) {
                            const float *yr = p.y + e * p.ldy;
#pragma unroll
                            for (int it = 0; it < ITERS; ++it) {
                                if (!act[it]) continue;
                                if (p.y_bcast == PGLB_BCAST_FULL) {
                                    VecIO<VEC>::ld(yr + col[it], yv[k][it]);
                                } else if (p.y_bcast == PGLB_BCAST_HEAD) {
#pragma unroll
                                    for (int j = 0; j < VEC; ++j)
                                        yv[k][it][j] = __ldg(yr + (col[it] + j) / p.head_dim);
                                } else {
                                    const float s = __ldg(yr);
#pragma unroll
                                    for (int j = 0; j < VEC; ++j) yv[k][it][j] = s;
                                }
                            }
                        }
                    }
                }
                // ---- phase B: consume in slot order, closing rows on the way ----
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    if (k0 + k < nb && !abandon) {
                        while (pos == cur_end) {
                            close_row();
                            ++row;
                            open_row();
                            if (hub) break;
                        }
                        if (hub) {
                            abandon = true;
                        } else {
#pragma unroll
                            for (int it = 0; it < ITERS; ++it) {
                                if (!act[it]) continue;
#pragma unroll
                                for (int j = 0; j < VEC; ++j) {
                                    float m = xv[k][it][j];
                                    if (p.scale_src) m = __fmul_rn(m, sv[k]);
                                    if (MODE == 1) m = apply_msg(p.msg_op, m, yv[k][it][j]);
                                    if (RK == 0) {
                                        acc[it][j] = __fadd_rn(acc[it][j], m);
                                    } else {
                                        acc[it][j] = is_max ? fmaxf(acc[it][j], m)
                                                            : fminf(acc[it][j], m);
                                    }
                                }
                            }
                            ++pos;
                        }
                    }
                }
            }
        }
    }
}

// One block: turn the hub list into chunk tables.
//   chunk_off[h]   exclusive prefix of ceil(deg_h / CH); chunk_off[n] = total
//   vbeg/vend[c]   slot range of chunk c
__global__ void __launch_bounds__(1024) hub_plan_kernel(const int64_t *indptr,
                                                        const unsigned long long *hub_count,
                                                        const int64_t *hub_rows,
                                                        int64_t *chunk_off, int64_t *vbeg,
                                                        int64_t *vend, int64_t *total_chunks,
                                                        int64_t *n_hubs_out) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t running;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int64_t n = (int64_t)*hub_count;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += blockDim.x) {
        const int64_t h = base + tid;
        int64_t b = 0, e = 0, nch = 0;
        if (h < n) {
            const int64_t r = hub_rows[h];
            b = indptr[r];
            e = indptr[r + 1];
            nch = (e - b + HUB_CH - 1) / HUB_CH;
        }
        // block exclusive scan of nch
        int64_t incl = nch;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int64_t w = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
            int64_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - w;  // exclusive
            if (lane == 31) warp_sums[31] = wi - w;
        }
        __syncthreads();
        const int64_t off = running + warp_sums[wid] + (incl - nch);
        if (h < n) {
            chunk_off[h] = off;
            for (int64_t c = 0; c < nch; ++c) {
                vbeg[off + c] = b + c * HUB_CH;
                const int64_t ce = b + (c + 1) * HUB_CH;
                vend[off + c] = ce < e ? ce : e;
            }
        }
        __syncthreads();
        if (tid == blockDim.x - 1) running = off + nch;
        __syncthreads();
    }
    if (tid == 0) {
        chunk_off[n] = running;
        *total_chunks = running;
        *n_hubs_out = n;
    }
}

// ---- host side dispatch ----------------------------------------------------------------

struct Shape {
    int vec, g, iters, tiles;
};

static Shape pick_shape(int64_t D, bool vec4) {
    Shape s;
    s.vec = vec4 ? 4 : 1;
    const int64_t cv = D / s.vec;
    s.iters = 1;
    s.tiles = 1;
    if (cv <= 4) s.g = 4;
    else if (cv <= 8) s.g = 8;
    else if (cv <= 16) s.g = 16;
    else {
        s.g = 32;
        if (cv <= 32) s.iters = 1;
        else if (cv <= 64) s.iters = 2;
        else s.iters = 4;
        s.tiles = (int)((cv + 32 * s.iters - 1) / (32 * s.iters));
    }
    return s;
}

typedef void (*SpmmKernel)(const SpmmP);

template <int VEC, int G, int ITERS>
static SpmmKernel pick_mode(int mode, int rk) {
    if (mode == 0) return rk == 0 ? spmm_csr_kernel<VEC, G, ITERS, 0, 0> : spmm_csr_kernel<VEC, G, ITERS, 0, 1>;
    return rk == 0 ? spmm_csr_kernel<VEC, G, ITERS, 1, 0> : spmm_csr_kernel<VEC, G, ITERS, 1, 1>;
}

template <int VEC>
static SpmmKernel pick_kernel_v(const Shape &s, int mode, int rk) {
    switch (s.g) {
        case 4: return pick_mode<VEC, 4, 1>(mode, rk);
        case 8: return pick_mode<VEC, 8, 1>(mode, rk);
        case 16: return pick_mode<VEC, 16, 1>(mode, rk);
        default:
            if (s.iters == 1) return pick_mode<VEC, 32, 1>(mode, rk);
            if (s.iters == 2) return pick_mode<VEC, 32, 2>(mode, rk);
            return pick_mode<VEC, 32, 4>(mode, rk);
    }
}

static SpmmKernel pick_kernel(const Shape &s, int mode, int rk) {
    return s.vec == 4 ? pick_kernel_v<4>(s, mode, rk) : pick_kernel_v<1>(s, mode, rk);
}

struct HubWs {
    unsigned long long *hub_count;  // [1]
    int64_t *total_chunks;          // [1]
    int64_t *n_hubs;                // [1]
    int64_t *hub_rows;              // [hub_cap]
    int64_t *chunk_off;             // [hub_cap + 1]
    int64_t *vbeg;                  // [chunk_cap]
    int64_t *vend;                  // [chunk_cap]
    float *partial;                 // [chunk_cap * dpad]
    int64_t hub_cap, chunk_cap, dpad;
    size_t bytes;
};

static HubWs layout_ws(void *ws, int64_t E, int64_t D) {
    HubWs h;
    h.hub_cap = E / (HUB_T + 1) + 1;
    h.chunk_cap = E / HUB_CH + h.hub_cap + 1;
    h.dpad = (D + 3) / 4 * 4;
    char *p = reinterpret_cast<char *>(ws);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *r = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return r;
    };
    h.hub_count = reinterpret_cast<unsigned long long *>(take(8));
    h.total_chunks = reinterpret_cast<int64_t *>(take(8));
    h.n_hubs = reinterpret_cast<int64_t *>(take(8));
    h.hub_rows = reinterpret_cast<int64_t *>(take(sizeof(int64_t) * h.hub_cap));
    h.chunk_off = reinterpret_cast<int64_t *>(take(sizeof(int64_t) * (h.hub_cap + 1)));
    h.vbeg = reinterpret_cast<int64_t *>(take(sizeof(int64_t) * h.chunk_cap));
    h.vend = reinterpret_cast<int64_t *>(take(sizeof(int64_t) * h.chunk_cap));
    h.partial = reinterpret_cast<float *>(take(sizeof(float) * h.chunk_cap * h.dpad));
    h.bytes = off;
    return h;
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// spmm_stream.cu
size_t stream_ws_bytes(int64_t E, int64_t D);
int spmm_stream_run(const int64_t *indptr, const int64_t *cols, const float *x, int64_t ldx,
                    float *out, int64_t ldo, int64_t n_dst, int64_t n_src, int64_t E, int64_t D,
                    int reduce_op, const float *scale_src, const float *scale_dst,
                    const uint32_t *cols32, int l2_hints, int accumulate, const int64_t *eid,
                    const float *y, int64_t ldy, int y_bcast, int head_dim, int msg_op, void *ws,
                    size_t ws_bytes, cudaStream_t stream);

int gat_fused_run(const int64_t *indptr, const int64_t *cols, const float *f, int64_t ldf, float *out,
                  int64_t ldo, int64_t n_dst, int64_t n_src, int64_t E, int64_t D, int64_t H,
                  const float *attn_src, const float *attn_dst, float slope, float *lse, void *ws, size_t ws_bytes,
                  cudaStream_t stream);

bool stream_narrow_enabled();  // PGLB_NARROW=1: experimental narrow-row streaming kernel (spmm_stream.cu)

static bool use_stream_path() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PGLB_SPMM_IMPL");
        v = (e && strcmp(e, "legacy") == 0) ? 0 : 1;
    }
    return v == 1;
}

}  // namespace pglb

using namespace pglb;

extern "C" int pglb_spmm_csr_ws(int64_t n_dst, int64_t num_edges, int64_t D, size_t *ws_bytes) {
    PGLB_CHECK_ARG(ws_bytes != nullptr, PGLB_EINVAL, "pglb_spmm_csr_ws: ws_bytes is NULL");
    PGLB_CHECK_ARG(n_dst >= 0 && num_edges >= 0 && D >= 0, PGLB_EINVAL,
                   "pglb_spmm_csr_ws: negative size");
    size_t a = layout_ws(nullptr, num_edges, D).bytes;
    size_t b = stream_ws_bytes(num_edges, D);
    *ws_bytes = a > b ? a : b;
    return PGLB_OK;
}

extern "C" int pglb_spmm_csr_f32(const int64_t *indptr, const int64_t *cols, const int64_t *eid,
                                 const float *x, int64_t ldx, const float *y, int64_t ldy,
                                 int y_bcast, float *out, int64_t ldo, int64_t n_dst,
                                 int64_t n_src, int64_t num_edges, int64_t D, int64_t head_dim,
                                 int msg_op, int reduce_op, const float *scale_src,
                                 const float *scale_dst, const uint32_t *cols_packed,
                                 int64_t max_degree_hint, int flags, void *ws, size_t ws_bytes,
                                 void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(n_dst >= 0 && n_src >= 0 && num_edges >= 0 && D >= 0, PGLB_EINVAL,
                   "pglb_spmm_csr_f32: negative size");
    PGLB_CHECK_ARG(reduce_op >= PGLB_REDUCE_SUM && reduce_op <= PGLB_REDUCE_MIN, PGLB_EINVAL,
                   "pglb_spmm_csr_f32: unknown reduce_op %d", reduce_op);
    PGLB_CHECK_ARG(msg_op >= PGLB_MSG_COPY && msg_op <= PGLB_MSG_DIV, PGLB_EINVAL,
                   "pglb_spmm_csr_f32: unknown msg_op %d", msg_op);
    if (n_dst == 0 || D == 0) return PGLB_OK;
    PGLB_CHECK_ARG(indptr && out, PGLB_EINVAL, "pglb_spmm_csr_f32: indptr/out is NULL");
    PGLB_CHECK_ARG(num_edges == 0 || x, PGLB_EINVAL, "pglb_spmm_csr_f32: x is NULL");
    PGLB_CHECK_ARG(ldx >= D && ldo >= D, PGLB_ESHAPE,
                   "pglb_spmm_csr_f32: leading dimension smaller than D");
    PGLB_CHECK_ARG(D <= INT32_MAX, PGLB_ESHAPE, "pglb_spmm_csr_f32: D too large");
    const int accumulate = (flags & PGLB_SPMM_ACCUMULATE) ? 1 : 0;
    PGLB_CHECK_ARG(!accumulate || reduce_op == PGLB_REDUCE_SUM, PGLB_EUNSUPPORTED,
                   "pglb_spmm_csr_f32: PGLB_SPMM_ACCUMULATE needs reduce_op SUM");
    const int mode = (msg_op == PGLB_MSG_COPY) ? 0 : 1;
    if (mode == 1) {
        PGLB_CHECK_ARG(y != nullptr, PGLB_EINVAL, "pglb_spmm_csr_f32: msg_op needs y");
        PGLB_CHECK_ARG(y_bcast >= PGLB_BCAST_FULL && y_bcast <= PGLB_BCAST_SCALAR, PGLB_EINVAL,
                       "pglb_spmm_csr_f32: unknown y_bcast %d", y_bcast);
        if (y_bcast == PGLB_BCAST_HEAD)
            PGLB_CHECK_ARG(head_dim > 0 && D % head_dim == 0, PGLB_ESHAPE,
                           "pglb_spmm_csr_f32: head_dim must divide D");
    }
    const bool vec4 = (D % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && aligned16(x) &&
                      aligned16(out) &&
                      (mode == 0 || y_bcast != PGLB_BCAST_FULL || ((ldy % 4 == 0) && aligned16(y)));
    // wide rows: the edge-balanced cp.async-ring kernels (spmm_stream.cu).  With an edge operand
    // only the GAT shape qualifies: per-head (head_dim % 4 == 0) or scalar y, mul / add, D <= 128.
    const bool ue_fast = mode == 1 && D <= 128 && scale_src == nullptr && num_edges < 0x7fffffffLL &&
                         (msg_op == PGLB_MSG_MUL || msg_op == PGLB_MSG_ADD) &&
                         (y_bcast == PGLB_BCAST_SCALAR || (y_bcast == PGLB_BCAST_HEAD && head_dim % 4 == 0));
    const bool narrow = mode == 0 && D <= 64 && reduce_op < PGLB_REDUCE_MAX && stream_narrow_enabled();
    if ((mode == 0 || ue_fast) && vec4 && (D > 64 || narrow) && num_edges > 0 && use_stream_path()) {
        PGLB_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, PGLB_EWORKSPACE,
                       "pglb_spmm_csr_f32: workspace must be 256-byte aligned");
        return spmm_stream_run(indptr, cols, x, ldx, out, ldo, n_dst, n_src, num_edges, D,
                               reduce_op, scale_src, scale_dst, mode == 0 ? cols_packed : nullptr,
                               (flags & PGLB_SPMM_L2_HINTS) ? 1 : 0, accumulate, eid,
                               mode == 1 ? y : nullptr, ldy, y_bcast, (int)head_dim, msg_op, ws,
                               ws_bytes, stream);
    }
    const Shape s = pick_shape(D, vec4);
    const int rk = (reduce_op >= PGLB_REDUCE_MAX) ? 1 : 0;

    const bool need_hub = !(max_degree_hint >= 0 && max_degree_hint <= HUB_T) && num_edges > HUB_T;
    HubWs h{};
    if (need_hub) {
        h = layout_ws(ws, num_edges, D);
        PGLB_CHECK_ARG(ws != nullptr && ws_bytes >= h.bytes, PGLB_EWORKSPACE,
                       "pglb_spmm_csr_f32: workspace of %zu bytes needed (got %zu)", h.bytes,
                       ws_bytes);
        PGLB_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, PGLB_EWORKSPACE,
                       "pglb_spmm_csr_f32: workspace must be 256-byte aligned");
        PGLB_CUDA(cudaMemsetAsync(h.hub_count, 0, 8, stream));
    }

    SpmmP p{};
    p.rbeg = indptr;
    p.rend = indptr + 1;
    p.cols = cols;
    p.eid = eid;
    p.x = x;
    p.ldx = ldx;
    p.y = y;
    p.ldy = ldy;
    p.y_bcast = y_bcast;
    p.out = out;
    p.ldo = ldo;
    p.n_rows = n_dst;
    p.n_rows_dev = nullptr;
    p.D = (int)D;
    p.head_dim = (int)(head_dim > 0 ? head_dim : 1);
    p.msg_op = msg_op;
    p.reduce_op = reduce_op;
    p.scale_src = scale_src;
    p.scale_dst = scale_dst;
    p.row_map = nullptr;
    p.mean_indptr = nullptr;
    p.rpg = 8;
    p.contig = 1;
    p.hub_threshold = need_hub ? HUB_T : INT64_MAX;
    p.hub_count = h.hub_count;
    p.hub_rows = h.hub_rows;
    p.accumulate = accumulate;

    const int64_t gpb = kThreads / s.g;
    {
        const int64_t tasks = (n_dst + p.rpg - 1) / p.rpg;
        const int64_t blocks = (tasks + gpb - 1) / gpb;
        PGLB_CHECK_ARG(blocks <= 0x7fffffffLL, PGLB_ESHAPE, "pglb_spmm_csr_f32: grid too large");
        dim3 grid((unsigned)blocks, (unsigned)s.tiles);
        pick_kernel(s, mode, rk)<<<grid, kThreads, 0, stream>>>(p);
        PGLB_LAUNCH_CHECK("spmm_csr_kernel");
    }
    if (need_hub) {
        hub_plan_kernel<<<1, 1024, 0, stream>>>(indptr, h.hub_count, h.hub_rows, h.chunk_off,
                                                h.vbeg, h.vend, h.total_chunks, h.n_hubs);
        PGLB_LAUNCH_CHECK("hub_plan_kernel");
        const int64_t max_blocks = (int64_t)sm_count() * 8;
        // pass 1: chunk -> partial (same message, SUM/MAX/MIN, no row epilogue)
        SpmmP p1 = p;
        p1.rbeg = h.vbeg;
        p1.rend = h.vend;
        p1.out = h.partial;
        p1.ldo = h.dpad;
        p1.n_rows = 0;
        p1.n_rows_dev = h.total_chunks;
        p1.reduce_op = (reduce_op == PGLB_REDUCE_MEAN) ? PGLB_REDUCE_SUM : reduce_op;
        p1.scale_dst = nullptr;
        p1.rpg = 1;
        p1.contig = 0;
        p1.accumulate = 0;
        p1.hub_threshold = INT64_MAX;
        {
            int64_t blocks = (h.chunk_cap + gpb - 1) / gpb;
            if (blocks > max_blocks) blocks = max_blocks;
            // partial rows are dpad floats, 16-byte aligned => same vector shape is valid
            dim3 grid((unsigned)blocks, (unsigned)s.tiles);
            pick_kernel(s, mode, rk)<<<grid, kThreads, 0, stream>>>(p1);
            PGLB_LAUNCH_CHECK("spmm_csr_kernel(hub pass 1)");
        }
        // pass 2: partials of a hub -> its output row
        SpmmP p2 = p;
        p2.rbeg = h.chunk_off;
        p2.rend = h.chunk_off + 1;
        p2.cols = nullptr;
        p2.eid = nullptr;
        p2.x = h.partial;
        p2.ldx = h.dpad;
        p2.y = nullptr;
        p2.msg_op = PGLB_MSG_COPY;
        p2.scale_src = nullptr;
        p2.n_rows = 0;
        p2.n_rows_dev = h.n_hubs;
        p2.row_map = h.hub_rows;
        p2.mean_indptr = indptr;
        p2.rpg = 1;
        p2.contig = 0;
        p2.hub_threshold = INT64_MAX;
        {
            int64_t blocks = (h.hub_cap + gpb - 1) / gpb;
            if (blocks > max_blocks) blocks = max_blocks;
            dim3 grid((unsigned)blocks, (unsigned)s.tiles);
            pick_kernel(s, 0, rk)<<<grid, kThreads, 0, stream>>>(p2);
            PGLB_LAUNCH_CHECK("spmm_csr_kernel(hub pass 2)");
        }
    }
    return PGLB_OK;
}

static int gat_fused_entry(const char *name, const int64_t *indptr, const int64_t *cols, const float *f, int64_t ldf,
                           const float *attn_src, const float *attn_dst, float negative_slope, float *out, int64_t ldo,
                           float *lse, int64_t n_dst, int64_t n_src, int64_t num_edges, int64_t H, int64_t head_dim,
                           void *ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(n_dst >= 0 && n_src >= 0 && num_edges >= 0 && H > 0 && head_dim > 0, PGLB_EINVAL, "%s: bad size", name);
    const int64_t D = H * head_dim;
    PGLB_CHECK_ARG(D <= 128 && head_dim % 4 == 0 && num_edges < 0x7fffffffLL && n_src < 0xffffffffLL,
                   PGLB_EUNSUPPORTED, "%s: needs H*head_dim <= 128 and head_dim %% 4 == 0", name);
    if (n_dst == 0) return PGLB_OK;
    PGLB_CHECK_ARG(indptr && out, PGLB_EINVAL, "%s: NULL pointer", name);
    PGLB_CHECK_ARG(ldf >= D && ldo >= D && ldf % 4 == 0 && ldo % 4 == 0 && aligned16(f) && aligned16(out),
                   PGLB_ESHAPE, "%s: rows must be 16-byte aligned", name);
    if (num_edges == 0) {
        PGLB_CUDA(cudaMemset2DAsync(out, sizeof(float) * ldo, 0, sizeof(float) * D, n_dst, stream));
        return PGLB_OK;
    }
    PGLB_CHECK_ARG(cols && f && attn_src && attn_dst, PGLB_EINVAL, "%s: NULL pointer", name);
    PGLB_CHECK_ARG(ws && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, PGLB_EWORKSPACE,
                   "%s: workspace NULL or not 256-byte aligned", name);
    return gat_fused_run(indptr, cols, f, ldf, out, ldo, n_dst, n_src, num_edges, D, H, attn_src,
                         attn_dst, negative_slope, lse, ws, ws_bytes, stream);
}

extern "C" int pglb_gat_fused_csr_f32(const int64_t *indptr, const int64_t *cols, const float *f,
                                      int64_t ldf, const float *attn_src, const float *attn_dst,
                                      float negative_slope, float *out, int64_t ldo, int64_t n_dst,
                                      int64_t n_src, int64_t num_edges, int64_t H, int64_t head_dim,
                                      void *ws, size_t ws_bytes, void *stream_) {
    return gat_fused_entry("pglb_gat_fused_csr_f32", indptr, cols, f, ldf, attn_src, attn_dst, negative_slope, out, ldo,
                           nullptr, n_dst, n_src, num_edges, H, head_dim, ws, ws_bytes, stream_);
}

// Training forward: the same launch, which additionally leaves lse[d, h] = log sum_j exp(leaky(as[src_j,h] + ad[d,h]))
// for every row with at least one in-edge (rows without are not written) -- all pglb_gat_bwd_edge_f32 needs to rebuild
// the attention weights.  PGLB_EUNSUPPORTED when the shape is outside the TMA kernel (the caller keeps the op-by-op path).
extern "C" int pglb_gat_fused_train_csr_f32(const int64_t *indptr, const int64_t *cols, const float *f,
                                            int64_t ldf, const float *attn_src, const float *attn_dst,
                                            float negative_slope, float *out, int64_t ldo, float *lse,
                                            int64_t n_dst, int64_t n_src, int64_t num_edges, int64_t H,
                                            int64_t head_dim, void *ws, size_t ws_bytes, void *stream_) {
    PGLB_CHECK_ARG(lse != nullptr || n_dst == 0, PGLB_EINVAL, "pglb_gat_fused_train_csr_f32: NULL lse");
    return gat_fused_entry("pglb_gat_fused_train_csr_f32", indptr, cols, f, ldf, attn_src, attn_dst, negative_slope, out,
                           ldo, lse, n_dst, n_src, num_edges, H, head_dim, ws, ws_bytes, stream_);
}
