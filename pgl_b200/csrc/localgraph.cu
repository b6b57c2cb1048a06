// localgraph.cu -- partition -> local-graph pipeline on the device.
//
//   * pglb_map_nodes / pglb_map_edges: the reference's graph_kernel.map_nodes / map_edges
//     (pgl/graph_kernel.pyx:104-138; caller pgl/sampling/custom.py:66-68).  The reference walks a C++
//     unordered_map per element on one host thread; here the mapping is a dense table new_id[old_id]
//     and every element is one coalesced load + one gather.
//   * pglb_invert_perm: new_id[perm[j]] = j -- the second half of "stable sort the nodes by part" (the
//     permutation / offsets convention of apps/GNNAutoScale/graph_partition.py:94-101); the first half is
//     pglb_csr_build with u = part (its stable radix sort, its histogram and its scan ARE that sort).
//   * pglb_halo_plan_count / pglb_halo_plan_fill: one rank's local graph of a 1-D node partition
//     (SURVEY.md section 8e): the edges whose destination the rank owns, their destination rows renumbered
//     to [0, n_local), their sources renumbered into [own rows | halo rows], and the sorted list of distinct
//     remote sources (the halo) with the number of them every peer owns.  "sorted distinct" is a flag array
//     over the node ids + a prefix sum (no sort, no hash): ascending id = grouped by owner because parts are
//     contiguous id ranges.
// Integer work throughout: results are defined exactly (tests compare with numpy restatements bit for bit).
#include "common.cuh"

namespace pglb {

__global__ void __launch_bounds__(256) map_nodes_kernel(const int64_t *__restrict__ nodes, int64_t n,
                                                        const int64_t *__restrict__ table, int64_t table_size,
                                                        int64_t *__restrict__ out, int *__restrict__ bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = ld_stream(nodes + i);
        int64_t r = -1;
        if (v >= 0 && v < table_size) r = ld_ro(table + v);
        else if (bad) *bad = 1;
        out[i] = r;
    }
}

// out[i, :] = table[edges[eid[i], :]]   (eid NULL: identity)
__global__ void __launch_bounds__(256) map_edges_kernel(const int64_t *__restrict__ eid, int64_t n,
                                                        const int64_t *__restrict__ edges, int64_t E,
                                                        const int64_t *__restrict__ table, int64_t table_size,
                                                        int64_t *__restrict__ out, int *__restrict__ bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = eid ? ld_stream(eid + i) : i;
        int64_t a = -1, b = -1;
        if (j >= 0 && j < E) {
            const longlong2 uv = __ldg(reinterpret_cast<const longlong2 *>(edges) + j);
            if (uv.x >= 0 && uv.x < table_size) a = ld_ro(table + uv.x);
            else if (bad) *bad = 1;
            if (uv.y >= 0 && uv.y < table_size) b = ld_ro(table + uv.y);
            else if (bad) *bad = 1;
        } else if (bad) {
            *bad = 1;
        }
        reinterpret_cast<longlong2 *>(out)[i] = make_longlong2(a, b);
    }
}

__global__ void __launch_bounds__(256) invert_perm_kernel(const int64_t *__restrict__ perm, int64_t n,
                                                          int64_t *__restrict__ inv) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        inv[ld_stream(perm + i)] = i;
}

// mine[e] = 1 when this rank owns the edge's destination; mark[s] = 1 for every remote source of such an edge
__global__ void __launch_bounds__(256) halo_mark_kernel(const int64_t *__restrict__ edges, int64_t E, int64_t lo,
                                                        int64_t hi, int64_t N, int64_t *__restrict__ mine,
                                                        int64_t *__restrict__ mark, int *__restrict__ bad) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const longlong2 uv = __ldcs(reinterpret_cast<const longlong2 *>(edges) + e);
        const int64_t s = uv.x, d = uv.y;
        if (s < 0 || s >= N || d < 0 || d >= N) {
            *bad = 1;
            mine[e] = 0;
            continue;
        }
        const bool m = d >= lo && d < hi;
        mine[e] = m ? 1 : 0;
        if (m && (s < lo || s >= hi)) mark[s] = 1;  // every writer stores the same value
    }
}

__global__ void halo_counts_kernel(const int64_t *__restrict__ mine_incl, int64_t E, const int64_t *__restrict__ mark_incl,
                                   int64_t N, int64_t *__restrict__ counts) {
    counts[0] = E > 0 ? mine_incl[E - 1] : 0;
    counts[1] = N > 0 ? mark_incl[N - 1] : 0;
}

__global__ void __launch_bounds__(256) halo_fill_edges_kernel(const int64_t *__restrict__ edges, int64_t E, int64_t lo,
                                                              int64_t hi, const int64_t *__restrict__ mine_incl,
                                                              const int64_t *__restrict__ mark_incl,
                                                              int64_t *__restrict__ eid, int64_t *__restrict__ dst_local,
                                                              int64_t *__restrict__ col_local) {
    const int64_t n_local = hi - lo;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t inc = mine_incl[e];
        const int64_t prev = e ? mine_incl[e - 1] : 0;
        if (inc == prev) continue;
        const longlong2 uv = __ldcs(reinterpret_cast<const longlong2 *>(edges) + e);
        const int64_t p = inc - 1;
        eid[p] = e;
        dst_local[p] = uv.y - lo;
        const bool remote = uv.x < lo || uv.x >= hi;
        col_local[p] = remote ? n_local + ld_ro(mark_incl + uv.x) - 1 : uv.x - lo;
    }
}

__global__ void __launch_bounds__(256) halo_fill_nodes_kernel(const int64_t *__restrict__ mark_incl, int64_t N,
                                                              int64_t *__restrict__ halo_ids) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t inc = mark_incl[i];
        const int64_t prev = i ? mark_incl[i - 1] : 0;
        if (inc != prev) halo_ids[inc - 1] = i;
    }
}

// recv_counts[p] = number of halo ids inside [offsets[p], offsets[p+1])
__global__ void halo_recv_counts_kernel(const int64_t *__restrict__ mark_incl, int64_t N,
                                        const int64_t *__restrict__ offsets, int K, int64_t *__restrict__ recv_counts) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= K) return;
    auto before = [&](int64_t x) -> int64_t {
        x = x < 0 ? 0 : (x > N ? N : x);
        return x > 0 ? mark_incl[x - 1] : 0;
    };
    recv_counts[p] = before(offsets[p + 1]) - before(offsets[p]);
}

static inline int grid_1d(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

struct HaloWs {
    int64_t *mine;  // [E] flags -> inclusive prefix
    int64_t *mark;  // [N] flags -> inclusive prefix
    void *scan_tmp;
    int *bad;
    size_t bytes;
};

static HaloWs halo_layout(void *ws, int64_t E, int64_t N) {
    HaloWs w;
    char *base = reinterpret_cast<char *>(ws);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *r = base ? base + off : nullptr;
        off += align_up(bytes ? bytes : 1, 256);
        return r;
    };
    w.mine = reinterpret_cast<int64_t *>(take(sizeof(int64_t) * (size_t)E));
    w.mark = reinterpret_cast<int64_t *>(take(sizeof(int64_t) * (size_t)N));
    const size_t se = scan_i64_ws_bytes(E), sn = scan_i64_ws_bytes(N);
    w.scan_tmp = take(se > sn ? se : sn);
    w.bad = reinterpret_cast<int *>(take(sizeof(int)));
    w.bytes = off;
    return w;
}

}  // namespace pglb

using namespace pglb;

extern "C" int pglb_map_nodes(const int64_t *nodes, int64_t n, const int64_t *table, int64_t table_size,
                              int64_t *out, int32_t *bad_flag, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(n >= 0 && table_size >= 0, PGLB_EINVAL, "pglb_map_nodes: bad size");
    if (n == 0) return PGLB_OK;
    PGLB_CHECK_ARG(nodes && out && (table || table_size == 0), PGLB_EINVAL, "pglb_map_nodes: NULL pointer");
    map_nodes_kernel<<<grid_1d(n), 256, 0, stream>>>(nodes, n, table, table_size, out, bad_flag);
    PGLB_LAUNCH_CHECK("map_nodes_kernel");
    return PGLB_OK;
}

extern "C" int pglb_map_edges(const int64_t *eid, int64_t n, const int64_t *edges, int64_t num_edges,
                              const int64_t *table, int64_t table_size, int64_t *out, int32_t *bad_flag,
                              void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(n >= 0 && num_edges >= 0 && table_size >= 0, PGLB_EINVAL, "pglb_map_edges: bad size");
    if (n == 0) return PGLB_OK;
    PGLB_CHECK_ARG(edges && out && (table || table_size == 0), PGLB_EINVAL, "pglb_map_edges: NULL pointer");
    PGLB_CHECK_ARG((reinterpret_cast<uintptr_t>(edges) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0,
                   PGLB_ESHAPE, "pglb_map_edges: edges / out must be contiguous [*, 2] int64, 16-byte aligned");
    map_edges_kernel<<<grid_1d(n), 256, 0, stream>>>(eid, n, edges, num_edges, table, table_size, out, bad_flag);
    PGLB_LAUNCH_CHECK("map_edges_kernel");
    return PGLB_OK;
}

extern "C" int pglb_invert_perm(const int64_t *perm, int64_t n, int64_t *inverse, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(n >= 0, PGLB_EINVAL, "pglb_invert_perm: bad size");
    if (n == 0) return PGLB_OK;
    PGLB_CHECK_ARG(perm && inverse, PGLB_EINVAL, "pglb_invert_perm: NULL pointer");
    invert_perm_kernel<<<grid_1d(n), 256, 0, stream>>>(perm, n, inverse);
    PGLB_LAUNCH_CHECK("invert_perm_kernel");
    return PGLB_OK;
}

extern "C" int pglb_halo_plan_ws(int64_t num_edges, int64_t num_nodes, size_t *ws_bytes) {
    PGLB_CHECK_ARG(ws_bytes != nullptr && num_edges >= 0 && num_nodes >= 0, PGLB_EINVAL, "pglb_halo_plan_ws: bad argument");
    *ws_bytes = halo_layout(nullptr, num_edges, num_nodes).bytes;
    return PGLB_OK;
}

extern "C" int pglb_halo_plan_count(const int64_t *edges, int64_t num_edges, int64_t num_nodes, int64_t lo, int64_t hi,
                                    int64_t *counts, int32_t *bad_flag, void *ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(num_edges >= 0 && num_nodes >= 0 && 0 <= lo && lo <= hi && hi <= num_nodes, PGLB_EINVAL,
                   "pglb_halo_plan_count: bad size / node range");
    PGLB_CHECK_ARG(counts && bad_flag, PGLB_EINVAL, "pglb_halo_plan_count: NULL pointer");
    HaloWs w = halo_layout(ws, num_edges, num_nodes);
    PGLB_CHECK_ARG(ws != nullptr && ws_bytes >= w.bytes && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, PGLB_EWORKSPACE,
                   "pglb_halo_plan_count: workspace of %zu bytes (256-byte aligned) needed (got %zu)", w.bytes, ws_bytes);
    PGLB_CHECK_ARG(num_edges == 0 || (edges && (reinterpret_cast<uintptr_t>(edges) & 15u) == 0), PGLB_EINVAL,
                   "pglb_halo_plan_count: edges must be a contiguous [E, 2] int64 array, 16-byte aligned");
    if (num_nodes) PGLB_CUDA(cudaMemsetAsync(w.mark, 0, sizeof(int64_t) * (size_t)num_nodes, stream));
    if (num_edges) {
        halo_mark_kernel<<<grid_1d(num_edges), 256, 0, stream>>>(edges, num_edges, lo, hi, num_nodes, w.mine, w.mark, bad_flag);
        PGLB_LAUNCH_CHECK("halo_mark_kernel");
    }
    int rc = scan_i64(w.mine, w.mine, num_edges, 1, w.scan_tmp, stream);
    if (rc) return rc;
    rc = scan_i64(w.mark, w.mark, num_nodes, 1, w.scan_tmp, stream);
    if (rc) return rc;
    halo_counts_kernel<<<1, 1, 0, stream>>>(w.mine, num_edges, w.mark, num_nodes, counts);
    PGLB_LAUNCH_CHECK("halo_counts_kernel");
    return PGLB_OK;
}

extern "C" int pglb_halo_plan_fill(const int64_t *edges, int64_t num_edges, int64_t num_nodes, int64_t lo, int64_t hi,
                                   const int64_t *offsets, int64_t num_parts, int64_t *eid, int64_t *dst_local,
                                   int64_t *col_local, int64_t *halo_ids, int64_t *recv_counts, void *ws,
                                   size_t ws_bytes, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PGLB_CHECK_ARG(num_edges >= 0 && num_nodes >= 0 && 0 <= lo && lo <= hi && hi <= num_nodes && num_parts >= 0 &&
                       num_parts <= 65536,
                   PGLB_EINVAL, "pglb_halo_plan_fill: bad size / node range");
    HaloWs w = halo_layout(ws, num_edges, num_nodes);
    PGLB_CHECK_ARG(ws != nullptr && ws_bytes >= w.bytes, PGLB_EWORKSPACE,
                   "pglb_halo_plan_fill: workspace of %zu bytes needed (got %zu)", w.bytes, ws_bytes);
    if (num_edges) {
        PGLB_CHECK_ARG(edges, PGLB_EINVAL, "pglb_halo_plan_fill: NULL pointer");
        halo_fill_edges_kernel<<<grid_1d(num_edges), 256, 0, stream>>>(edges, num_edges, lo, hi, w.mine, w.mark, eid,
                                                                      dst_local, col_local);
        PGLB_LAUNCH_CHECK("halo_fill_edges_kernel");
    }
    if (num_nodes) {
        halo_fill_nodes_kernel<<<grid_1d(num_nodes), 256, 0, stream>>>(w.mark, num_nodes, halo_ids);
        PGLB_LAUNCH_CHECK("halo_fill_nodes_kernel");
    }
    if (num_parts) {
        PGLB_CHECK_ARG(offsets && recv_counts, PGLB_EINVAL, "pglb_halo_plan_fill: NULL offsets / recv_counts");
        halo_recv_counts_kernel<<<(unsigned)((num_parts + 63) / 64), 64, 0, stream>>>(w.mark, num_nodes, offsets,
                                                                                      (int)num_parts, recv_counts);
        PGLB_LAUNCH_CHECK("halo_recv_counts_kernel");
    }
    return PGLB_OK;
}
