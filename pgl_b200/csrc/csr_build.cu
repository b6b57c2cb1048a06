// csr_build.cu -- device CSR/CSC construction, bit-identical to the reference's stable
// counting sort (pgl/graph_kernel.pyx:59-88): degree histogram, exclusive scan, and a
// STABLE key sort of (u, edge id) so that every bucket keeps ascending edge id.
//
// The key sort and the scans use CUB (device-wide LSD radix sort is stable by construction;
// CUB ships inside the CUDA toolkit and is compiled into this library for sm_100a).  It is
// one-off index preparation, cached on the EdgeIndex; the per-layer hot loop never calls it.
#include <cub/cub.cuh>

#include "common.cuh"

namespace pglb {

template <typename KeyT, typename ValT>
__global__ void __launch_bounds__(256) pack_keys_kernel(const int64_t *u, int64_t u_stride,
                                                        int64_t E, KeyT *keys, ValT *vals,
                                                        unsigned long long *degree) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = u[i * u_stride];
        keys[i] = (KeyT)k;
        vals[i] = (ValT)i;
        atomicAdd(degree + k, 1ull);
    }
}

template <typename KeyT, typename ValT>
__global__ void __launch_bounds__(256) emit_sorted_kernel(const KeyT *keys, const ValT *vals,
                                                          const int64_t *v, int64_t v_stride,
                                                          int64_t E, int64_t *sorted_u,
                                                          int64_t *sorted_v,
                                                          int64_t *sorted_eid) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = (int64_t)vals[i];
        sorted_u[i] = (int64_t)keys[i];
        sorted_eid[i] = e;
        sorted_v[i] = v[e * v_stride];
    }
}

__global__ void set_first_kernel(int64_t *indptr) { indptr[0] = 0; }

static int key_bits(int64_t n) {
    int b = 1;
    while (b < 63 && ((int64_t)1 << b) < n) ++b;
    return b;
}

template <typename KeyT, typename ValT>
struct SortPlan {
    size_t keys_in, keys_out, vals_in, vals_out, scan_tmp, sort_tmp, total;
    size_t scan_tmp_bytes, sort_tmp_bytes;
};

template <typename KeyT, typename ValT>
static cudaError_t plan(int64_t E, int64_t N, SortPlan<KeyT, ValT> &pl) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t r = off;
        off += align_up(bytes ? bytes : 1, 256);
        return r;
    };
    pl.keys_in = take(sizeof(KeyT) * E);
    pl.keys_out = take(sizeof(KeyT) * E);
    pl.vals_in = take(sizeof(ValT) * E);
    pl.vals_out = take(sizeof(ValT) * E);
    pl.scan_tmp_bytes = 0;
    cudaError_t e = cub::DeviceScan::InclusiveSum(nullptr, pl.scan_tmp_bytes, (const int64_t *)nullptr,
                                                  (int64_t *)nullptr, N);
    if (e != cudaSuccess) return e;
    pl.sort_tmp_bytes = 0;
    e = cub::DeviceRadixSort::SortPairs(nullptr, pl.sort_tmp_bytes, (const KeyT *)nullptr,
                                        (KeyT *)nullptr, (const ValT *)nullptr, (ValT *)nullptr,
                                        E, 0, key_bits(N));
    if (e != cudaSuccess) return e;
    pl.scan_tmp = take(pl.scan_tmp_bytes);
    pl.sort_tmp = take(pl.sort_tmp_bytes);
    pl.total = off;
    return cudaSuccess;
}

template <typename KeyT, typename ValT>
static int run(const int64_t *u, int64_t us, const int64_t *v, int64_t vs, int64_t E, int64_t N,
               int64_t *degree, int64_t *indptr, int64_t *su, int64_t *sv, int64_t *se, void *ws,
               size_t ws_bytes, cudaStream_t stream) {
    SortPlan<KeyT, ValT> pl;
    PGLB_CUDA((plan<KeyT, ValT>(E, N, pl)));
    PGLB_CHECK_ARG(ws_bytes >= pl.total, PGLB_EWORKSPACE,
                   "pglb_csr_build: workspace of %zu bytes needed (got %zu)", pl.total, ws_bytes);
    char *base = reinterpret_cast<char *>(ws);
    KeyT *keys_in = reinterpret_cast<KeyT *>(base + pl.keys_in);
    KeyT *keys_out = reinterpret_cast<KeyT *>(base + pl.keys_out);
    ValT *vals_in = reinterpret_cast<ValT *>(base + pl.vals_in);
    ValT *vals_out = reinterpret_cast<ValT *>(base + pl.vals_out);

    PGLB_CUDA(cudaMemsetAsync(degree, 0, sizeof(int64_t) * N, stream));
    const int blocks = (int)std::min<int64_t>((E + 255) / 256, (int64_t)sm_count() * 16);
    if (E > 0) {
        pack_keys_kernel<KeyT, ValT><<<blocks, 256, 0, stream>>>(
            u, us, E, keys_in, vals_in, reinterpret_cast<unsigned long long *>(degree));
        PGLB_LAUNCH_CHECK("pack_keys_kernel");
    }
    set_first_kernel<<<1, 1, 0, stream>>>(indptr);
    PGLB_LAUNCH_CHECK("set_first_kernel");
    if (N > 0) {
        size_t tb = pl.scan_tmp_bytes;
        PGLB_CUDA(cub::DeviceScan::InclusiveSum(base + pl.scan_tmp, tb, degree, indptr + 1, N, stream));
        count_launch(2);
    }
    if (E > 0) {
        size_t tb = pl.sort_tmp_bytes;
        PGLB_CUDA(cub::DeviceRadixSort::SortPairs(base + pl.sort_tmp, tb, keys_in, keys_out, vals_in,
                                                  vals_out, E, 0, key_bits(N), stream));
        count_launch(4);
        emit_sorted_kernel<KeyT, ValT><<<blocks, 256, 0, stream>>>(keys_out, vals_out, v, vs, E, su,
                                                                   sv, se);
        PGLB_LAUNCH_CHECK("emit_sorted_kernel");
    }
    return PGLB_OK;
}

// ---- segment ids ------------------------------------------------------------------------

// flag[r] = degree(r) > 0 ; then uniq = compaction of rows, seg id of slot j = rank(row(j))
__global__ void __launch_bounds__(256) nonempty_flag_kernel(const int64_t *indptr, int64_t N,
                                                            int64_t *flag) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N;
         r += (int64_t)gridDim.x * blockDim.x)
        flag[r] = (indptr[r + 1] > indptr[r]) ? 1 : 0;
}

// rank[] is the EXCLUSIVE prefix sum of flag[]
__global__ void __launch_bounds__(256) emit_segments_kernel(const int64_t *indptr,
                                                            const int64_t *rank, int64_t N,
                                                            int64_t *uniq, int64_t *seg,
                                                            int64_t *num_uniq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < N; r += nwarps) {
        const int64_t b = indptr[r], e = indptr[r + 1];
        if (e > b) {
            const int64_t k = rank[r];
            if (lane == 0) uniq[k] = r;
            for (int64_t j = b + lane; j < e; j += 32) seg[j] = k;
        }
        if (r == N - 1 && lane == 0) *num_uniq = rank[r] + (e > b ? 1 : 0);
    }
}

// sorted ids -> indptr[K+1]: indptr[k] = first slot whose id >= k
__global__ void __launch_bounds__(256) segment_indptr_kernel(const int64_t *ids, int64_t E,
                                                             int64_t K, int64_t *indptr) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= E;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t prev = (i == 0) ? -1 : ids[i - 1];
        const int64_t cur = (i == E) ? K : ids[i];
        for (int64_t k = prev + 1; k <= cur && k <= K; ++k) indptr[k] = i;
    }
}

}  // namespace pglb

using namespace pglb;

static bool use_small(int64_t E, int64_t N) { return E < 0xffffffffLL && N < 0xffffffffLL; }

extern "C" int pglb_csr_build_ws(int64_t E, int64_t N, size_t *ws_bytes) {
    PGLB_CHECK_ARG(ws_bytes != nullptr, PGLB_EINVAL, "pglb_csr_build_ws: ws_bytes is NULL");
    PGLB_CHECK_ARG(E >= 0 && N >= 0, PGLB_EINVAL, "pglb_csr_build_ws: negative size");
    if (use_small(E, N)) {
        SortPlan<uint32_t, uint32_t> pl;
        PGLB_CUDA((plan<uint32_t, uint32_t>(E, N, pl)));
        *ws_bytes = pl.total;
    } else {
        SortPlan<uint64_t, uint64_t> pl;
        PGLB_CUDA((plan<uint64_t, uint64_t>(E, N, pl)));
        *ws_bytes = pl.total;
    }
    return PGLB_OK;
}

extern "C" int pglb_csr_build(const int64_t *u, int64_t u_stride, const int64_t *v,
                              int64_t v_stride, int64_t E, int64_t N, int64_t *degree,
                              int64_t *indptr, int64_t *sorted_u, int64_t *sorted_v,
                              int64_t *sorted_eid, void *ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(E >= 0 && N >= 0, PGLB_EINVAL, "pglb_csr_build: negative size");
    PGLB_CHECK_ARG(indptr != nullptr, PGLB_EINVAL, "pglb_csr_build: indptr is NULL");
    PGLB_CHECK_ARG(N == 0 || degree != nullptr, PGLB_EINVAL, "pglb_csr_build: degree is NULL");
    PGLB_CHECK_ARG(E == 0 || (u && v && sorted_u && sorted_v && sorted_eid), PGLB_EINVAL,
                   "pglb_csr_build: NULL edge pointer");
    PGLB_CHECK_ARG(E == 0 || N > 0, PGLB_ESHAPE, "pglb_csr_build: edges but no nodes");
    PGLB_CHECK_ARG(u_stride >= 1 && v_stride >= 1, PGLB_EINVAL, "pglb_csr_build: bad stride");
    PGLB_CHECK_ARG(ws != nullptr && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0,
                   PGLB_EWORKSPACE, "pglb_csr_build: workspace NULL or not 256-byte aligned");
    if (use_small(E, N))
        return run<uint32_t, uint32_t>(u, u_stride, v, v_stride, E, N, degree, indptr, sorted_u,
                                       sorted_v, sorted_eid, ws, ws_bytes, stream);
    return run<uint64_t, uint64_t>(u, u_stride, v, v_stride, E, N, degree, indptr, sorted_u,
                                   sorted_v, sorted_eid, ws, ws_bytes, stream);
}

extern "C" int pglb_segment_ids_ws(int64_t N, size_t *ws_bytes) {
    PGLB_CHECK_ARG(ws_bytes != nullptr && N >= 0, PGLB_EINVAL, "pglb_segment_ids_ws: bad args");
    size_t tmp = 0;
    PGLB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int64_t *)nullptr,
                                            (int64_t *)nullptr, N));
    *ws_bytes = align_up(sizeof(int64_t) * (N ? N : 1), 256) * 2 + align_up(tmp ? tmp : 1, 256);
    return PGLB_OK;
}

extern "C" int pglb_segment_ids(const int64_t *indptr, int64_t N, int64_t E, int64_t *uniq_ind,
                                int64_t *segment_ids, int64_t *num_uniq, void *ws,
                                size_t ws_bytes, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(N >= 0 && E >= 0 && num_uniq != nullptr, PGLB_EINVAL,
                   "pglb_segment_ids: bad args");
    if (N == 0) {
        PGLB_CUDA(cudaMemsetAsync(num_uniq, 0, sizeof(int64_t), stream));
        return PGLB_OK;
    }
    PGLB_CHECK_ARG(indptr && uniq_ind && (E == 0 || segment_ids), PGLB_EINVAL,
                   "pglb_segment_ids: NULL pointer");
    size_t need = 0;
    int rc = pglb_segment_ids_ws(N, &need);
    if (rc) return rc;
    PGLB_CHECK_ARG(ws && ws_bytes >= need, PGLB_EWORKSPACE,
                   "pglb_segment_ids: workspace of %zu bytes needed (got %zu)", need, ws_bytes);
    char *base = reinterpret_cast<char *>(ws);
    const size_t a = align_up(sizeof(int64_t) * N, 256);
    int64_t *flag = reinterpret_cast<int64_t *>(base);
    int64_t *rank = reinterpret_cast<int64_t *>(base + a);
    void *tmp = base + 2 * a;
    size_t tmp_bytes = ws_bytes - 2 * a;
    const int blocks = (int)std::min<int64_t>((N + 255) / 256, (int64_t)sm_count() * 16);
    nonempty_flag_kernel<<<blocks, 256, 0, stream>>>(indptr, N, flag);
    PGLB_LAUNCH_CHECK("nonempty_flag_kernel");
    PGLB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag, rank, N, stream));
    count_launch(2);
    const int wblocks = (int)std::min<int64_t>((N + 7) / 8, (int64_t)sm_count() * 16);
    emit_segments_kernel<<<wblocks, 256, 0, stream>>>(indptr, rank, N, uniq_ind, segment_ids,
                                                      num_uniq);
    PGLB_LAUNCH_CHECK("emit_segments_kernel");
    return PGLB_OK;
}

extern "C" int pglb_segment_indptr(const int64_t *segment_ids, int64_t E, int64_t K,
                                   int64_t *indptr, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(E >= 0 && K >= 0 && indptr != nullptr, PGLB_EINVAL,
                   "pglb_segment_indptr: bad args");
    PGLB_CHECK_ARG(E == 0 || segment_ids != nullptr, PGLB_EINVAL,
                   "pglb_segment_indptr: segment_ids is NULL");
    const int blocks = (int)std::min<int64_t>((E + 256) / 256, (int64_t)sm_count() * 16);
    segment_indptr_kernel<<<blocks, 256, 0, stream>>>(segment_ids, E, K, indptr);
    PGLB_LAUNCH_CHECK("segment_indptr_kernel");
    return PGLB_OK;
}
