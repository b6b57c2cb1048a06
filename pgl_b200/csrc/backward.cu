// backward.cu -- gradient kernels of the send/recv path.
// The reference gets these from Paddle autograd; here:
//   * sum / mean copy aggregation backward  = the forward kernel on the reverse CSR (no new code)
//   * send_ue_recv(mul) wrt the edge operand = per-head dot of x[src] and grad[dst]  (sddmm_dot)
//   * edge_softmax backward                 = alpha * (g - sum_row(alpha * g))
//   * max / min copy aggregation backward   = grad routed to the entries that equal the output
//   * fused GAT aggregation backward        = one edge kernel (attention weights rebuilt from the saved row
//     log-sum-exp, SDDMM dot, softmax / LeakyReLU backward) + the existing reverse-CSR aggregations
#include "common.cuh"

namespace pglb {

// out[e, h] = sum_k a[ia[e], h, k] * b[ib[e], h, k]
template <int VEC>
__global__ void __launch_bounds__(256) sddmm_dot_kernel(const float *__restrict__ a,
                                                        const float *__restrict__ b,
                                                        const int64_t *__restrict__ ia, int64_t sa,
                                                        const int64_t *__restrict__ ib, int64_t sb,
                                                        int64_t E, int H, int Dh,
                                                        float *__restrict__ out) {
    const int64_t total = E * H;
    const int64_t D = (int64_t)H * Dh;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / H;
        const int h = (int)(i - e * H);
        const float *pa = a + __ldg((const long long *)ia + e * sa) * D + (int64_t)h * Dh;
        const float *pb = b + __ldg((const long long *)ib + e * sb) * D + (int64_t)h * Dh;
        float acc = 0.0f;
        if (VEC == 4) {
            for (int k = 0; k < Dh; k += 4) {
                const float4 x = __ldg(reinterpret_cast<const float4 *>(pa + k));
                const float4 y = __ldg(reinterpret_cast<const float4 *>(pb + k));
                acc = fmaf(x.x, y.x, acc);
                acc = fmaf(x.y, y.y, acc);
                acc = fmaf(x.z, y.z, acc);
                acc = fmaf(x.w, y.w, acc);
            }
        } else {
            for (int k = 0; k < Dh; ++k) acc = fmaf(__ldg(pa + k), __ldg(pb + k), acc);
        }
        out[i] = acc;
    }
}

// per CSR row: s_h = sum_j alpha[e_j,h] * g[e_j,h];  gl[e_j,h] = alpha[e_j,h] * (g[e_j,h] - s_h)
__global__ void __launch_bounds__(256) edge_softmax_bwd_kernel(
    const int64_t *__restrict__ indptr, const int64_t *__restrict__ eid,
    const float *__restrict__ alpha, const float *__restrict__ g, float *__restrict__ gl,
    int64_t n_rows, int H, int hp) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int head = blockIdx.y * hp + (lane % hp);
    const int srow = lane / hp, sstep = 32 / hp;
    const bool hact = head < H;
    for (int64_t r = warp; r < n_rows; r += nwarps) {
        const int64_t b = ld_ro(indptr + r), e = ld_ro(indptr + r + 1);
        if (e == b) continue;
        float s = 0.0f;
        for (int64_t j = b + srow; j < e; j += sstep) {
            const int64_t id = eid ? __ldg((const long long *)eid + j) : j;
            if (hact) s = fmaf(__ldg(alpha + id * H + head), __ldg(g + id * H + head), s);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1)
            if (o >= hp) s += __shfl_xor_sync(0xffffffffu, s, o);
        for (int64_t j = b + srow; j < e; j += sstep) {
            const int64_t id = eid ? __ldg((const long long *)eid + j) : j;
            if (hact) gl[id * H + head] = __ldg(alpha + id * H + head) * (__ldg(g + id * H + head) - s);
        }
    }
}

// reverse (src-keyed) CSR: gx[s, k] = sum_{slots j of row s} g[cols[j], k] * (x[s, k] == out[cols[j], k])
__global__ void __launch_bounds__(256) maxmin_bwd_kernel(const int64_t *__restrict__ indptr,
                                                         const int64_t *__restrict__ cols,
                                                         const float *__restrict__ x,
                                                         const float *__restrict__ out,
                                                         const float *__restrict__ g,
                                                         float *__restrict__ gx, int64_t n_src,
                                                         int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < n_src; s += nwarps) {
        const int64_t b = ld_ro(indptr + s), e = ld_ro(indptr + s + 1);
        for (int c = lane; c < D; c += 32) {
            const float xv = __ldg(x + s * D + c);
            float acc = 0.0f;
            for (int64_t j = b; j < e; ++j) {
                const int64_t d = ld_ro(cols + j);
                if (__ldg(out + d * D + c) == xv) acc += __ldg(g + d * D + c);
            }
            gx[s * D + c] = acc;
        }
    }
}


// Backward of the single-pass GAT aggregation (pglb_gat_fused_train_csr_f32), the per-edge part.  With
//   z = as[src,h] + ad[dst,h],  lz = leaky_relu(z),  alpha = exp(lz - lse[dst,h]),  out[dst,h,:] = sum alpha f[src,h,:]
// and G = d loss / d out:
//   d alpha = <G[dst,h,:], f[src,h,:]>,   d lz = alpha (d alpha - <G[dst,h,:], out[dst,h,:]>),   d z = d lz * leaky'(z)
//   (leaky'(z) = 1 for z > 0, slope otherwise).
// One launch rebuilds alpha and produces dz for every edge, both written in ORIGINAL edge order, which is what the
// reverse-CSR aggregations that finish the job read (grad f = sum over out-edges alpha G[dst]; grad as / grad ad =
// segment sums of dz over the src- / dst-CSR, exactly send_uv(add)'s backward).  Neither logits nor alpha were kept by
// the forward.  Walks the dst-CSR slots (rows[] = destination of every slot, so G / out / ad / lse are re-read only when
// the row changes); a warp owns 32 consecutive slots, the D/4 active lanes hold one float4 of the row each, LPH = Dh/4
// lanes form a head; four gathered feature rows are in flight per lane before the first dot product.
template <int LPH>
__global__ void __launch_bounds__(256) gat_bwd_edge_kernel(
    const int64_t *__restrict__ rows, const int64_t *__restrict__ cols, const int64_t *__restrict__ eid,
    const float *__restrict__ f, int64_t ldf, const float *__restrict__ g, int64_t ldg,
    const float *__restrict__ out, int64_t ldo, const float *__restrict__ a_s, const float *__restrict__ a_d,
    const float *__restrict__ lse, float slope, int64_t E, int H, int D, float *__restrict__ alpha_e,
    float *__restrict__ dz_e) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool act = lane * 4 < D;
    const int head = act ? lane / LPH : 0;
    const bool lead = act && (lane % LPH) == 0;
    auto head_sum = [&](float v) {
#pragma unroll
        for (int o = LPH / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    auto dot4 = [](const float4 &a, const float4 &b) {
        return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
    };
    for (int64_t base = warp * 32; base < E; base += nwarps * 32) {
        const int64_t j = base + lane;
        const bool in = j < E;
        const long long r_l = in ? ld_stream(rows + j) : 0;
        const long long c_l = in ? ld_stream(cols + j) : 0;
        const long long i_l = in ? (eid ? ld_stream(eid + j) : j) : 0;
        const int cnt = (E - base) < 32 ? (int)(E - base) : 32;
        long long cur = -1;
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float ad = 0.f, ls = 0.f, delta = 0.f;
        for (int k0 = 0; k0 < cnt; k0 += 4) {
            float4 fv[4];
            float as[4];
            long long rr[4], ii[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = (k0 + u) & 31;
                const long long cc = __shfl_sync(0xffffffffu, c_l, k);
                rr[u] = __shfl_sync(0xffffffffu, r_l, k);
                ii[u] = __shfl_sync(0xffffffffu, i_l, k);
                fv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                as[u] = 0.f;
                if (k0 + u < cnt && act) {
                    fv[u] = __ldg(reinterpret_cast<const float4 *>(f + cc * ldf + lane * 4));
                    as[u] = __ldg(a_s + cc * H + head);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (k0 + u >= cnt) break;
                if (rr[u] != cur) {  // warp-uniform
                    cur = rr[u];
                    float d = 0.f;
                    if (act) {
                        g4 = __ldg(reinterpret_cast<const float4 *>(g + cur * ldg + lane * 4));
                        const float4 o4 = __ldg(reinterpret_cast<const float4 *>(out + cur * ldo + lane * 4));
                        ad = __ldg(a_d + cur * H + head);
                        ls = __ldg(lse + cur * H + head);
                        d = dot4(g4, o4);
                    }
                    delta = head_sum(d);
                }
                const float da = head_sum(dot4(fv[u], g4));
                const float z = as[u] + ad;
                const bool pos = z > 0.0f;   // leaky_relu'(0) = slope, the convention of the framework's LeakyReLU backward
                const float alpha = expf((pos ? z : z * slope) - ls);
                const float dl = alpha * (da - delta);
                if (lead) {
                    alpha_e[ii[u] * H + head] = alpha;
                    dz_e[ii[u] * H + head] = pos ? dl : dl * slope;
                }
            }
        }
    }
}

static inline int grid_of(int64_t total, int cap_mult) {
    int64_t b = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * cap_mult;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace pglb

using namespace pglb;

extern "C" int pglb_sddmm_dot_f32(const float *a, const float *b, const int64_t *ia, int64_t ia_stride,
                                  const int64_t *ib, int64_t ib_stride, int64_t E, int64_t H,
                                  int64_t Dh, float *out, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(E >= 0 && H >= 0 && Dh >= 0 && H * Dh <= INT32_MAX, PGLB_EINVAL,
                   "pglb_sddmm_dot_f32: bad size");
    if (E == 0 || H == 0) return PGLB_OK;
    PGLB_CHECK_ARG(a && b && ia && ib && out, PGLB_EINVAL, "pglb_sddmm_dot_f32: NULL pointer");
    PGLB_CHECK_ARG(ia_stride >= 1 && ib_stride >= 1, PGLB_EINVAL, "pglb_sddmm_dot_f32: bad stride");
    const bool v4 = (Dh % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0;
    if (v4)
        sddmm_dot_kernel<4><<<grid_of(E * H, 32), 256, 0, stream>>>(a, b, ia, ia_stride, ib, ib_stride,
                                                                    E, (int)H, (int)Dh, out);
    else
        sddmm_dot_kernel<1><<<grid_of(E * H, 32), 256, 0, stream>>>(a, b, ia, ia_stride, ib, ib_stride,
                                                                    E, (int)H, (int)Dh, out);
    PGLB_LAUNCH_CHECK("sddmm_dot_kernel");
    return PGLB_OK;
}

extern "C" int pglb_edge_softmax_bwd_csr_f32(const int64_t *indptr, const int64_t *eid,
                                             const float *alpha, const float *grad,
                                             float *grad_logits, int64_t n_rows, int64_t E,
                                             int64_t H, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(n_rows >= 0 && E >= 0 && H >= 0 && H <= INT32_MAX, PGLB_EINVAL,
                   "pglb_edge_softmax_bwd_csr_f32: bad size");
    if (n_rows == 0 || E == 0 || H == 0) return PGLB_OK;
    PGLB_CHECK_ARG(indptr && alpha && grad && grad_logits, PGLB_EINVAL,
                   "pglb_edge_softmax_bwd_csr_f32: NULL pointer");
    int hp = 1;
    while (hp < H && hp < 32) hp <<= 1;
    const int tiles = (int)((H + hp - 1) / hp);
    dim3 grid((unsigned)grid_of(n_rows * 32, 64), (unsigned)tiles);
    edge_softmax_bwd_kernel<<<grid, 256, 0, stream>>>(indptr, eid, alpha, grad, grad_logits, n_rows,
                                                      (int)H, hp);
    PGLB_LAUNCH_CHECK("edge_softmax_bwd_kernel");
    return PGLB_OK;
}

extern "C" int pglb_gat_bwd_edge_f32(const int64_t *rows, const int64_t *cols, const int64_t *eid, const float *f,
                                     int64_t ldf, const float *grad_out, int64_t ldg, const float *out, int64_t ldo,
                                     const float *attn_src, const float *attn_dst, const float *lse,
                                     float negative_slope, int64_t num_edges, int64_t H, int64_t head_dim,
                                     float *alpha_e, float *dz_e, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(num_edges >= 0 && H > 0 && head_dim > 0, PGLB_EINVAL, "pglb_gat_bwd_edge_f32: bad size");
    const int64_t D = H * head_dim;
    const int64_t lph = head_dim / 4;
    PGLB_CHECK_ARG(D <= 128 && head_dim % 4 == 0 && (lph & (lph - 1)) == 0, PGLB_EUNSUPPORTED,
                   "pglb_gat_bwd_edge_f32: needs H*head_dim <= 128 and head_dim in {4, 8, 16, 32, 64, 128}");
    if (num_edges == 0) return PGLB_OK;
    PGLB_CHECK_ARG(rows && cols && f && grad_out && out && attn_src && attn_dst && lse && alpha_e && dz_e, PGLB_EINVAL,
                   "pglb_gat_bwd_edge_f32: NULL pointer");
    auto a16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    PGLB_CHECK_ARG(ldf >= D && ldg >= D && ldo >= D && ldf % 4 == 0 && ldg % 4 == 0 && ldo % 4 == 0 && a16(f) &&
                       a16(grad_out) && a16(out),
                   PGLB_ESHAPE, "pglb_gat_bwd_edge_f32: rows must be 16-byte aligned");
    const int grid = grid_of(num_edges * 1, 64);  // a warp per 32 slots, grid-stride
#define PGLB_GBE(L)                                                                                                   \
    gat_bwd_edge_kernel<L><<<grid, 256, 0, stream>>>(rows, cols, eid, f, ldf, grad_out, ldg, out, ldo, attn_src,      \
                                                    attn_dst, lse, negative_slope, num_edges, (int)H, (int)D,        \
                                                    alpha_e, dz_e)
    switch ((int)lph) {
        case 1: PGLB_GBE(1); break;
        case 2: PGLB_GBE(2); break;
        case 4: PGLB_GBE(4); break;
        case 8: PGLB_GBE(8); break;
        case 16: PGLB_GBE(16); break;
        default: PGLB_GBE(32); break;
    }
#undef PGLB_GBE
    PGLB_LAUNCH_CHECK("gat_bwd_edge_kernel");
    return PGLB_OK;
}

extern "C" int pglb_maxmin_bwd_f32(const int64_t *src_indptr, const int64_t *dst_of_slot,
                                   const float *x, const float *out, const float *grad_out,
                                   float *grad_x, int64_t n_src, int64_t D, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(n_src >= 0 && D >= 0 && D <= INT32_MAX, PGLB_EINVAL, "pglb_maxmin_bwd_f32: bad size");
    if (n_src == 0 || D == 0) return PGLB_OK;
    PGLB_CHECK_ARG(src_indptr && x && out && grad_out && grad_x, PGLB_EINVAL,
                   "pglb_maxmin_bwd_f32: NULL pointer");
    maxmin_bwd_kernel<<<grid_of(n_src * 32, 64), 256, 0, stream>>>(src_indptr, dst_of_slot, x, out,
                                                                   grad_out, grad_x, n_src, (int)D);
    PGLB_LAUNCH_CHECK("maxmin_bwd_kernel");
    return PGLB_OK;
}
