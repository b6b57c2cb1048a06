/*
 * oracle_c.c -- plain-C CPU restatement of the reference's send/recv hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Never by pgl_b200/.
 *
 * Each function cites what it follows (paths relative to /root/reference).  The float
 * kernels live in PaddlePaddle (unvendored, unpinned, absent here): the loops below
 * restate Paddle phi's CPU send_u_recv / segment_pool contract -- zero-initialised
 * output, sequential loop over edges in input order, mean = sum / count where
 * count > 0, max/min seeded by the first message of a row, rows without a message 0.
 * Pinned by the reference's own known-answer tests (tests/golden/kat_*.json).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { ORC_SUM = 0, ORC_MEAN = 1, ORC_MAX = 2, ORC_MIN = 3 };

/* pgl/graph_kernel.pyx:59-88 build_index: histogram, exclusive scan, stable counting sort */
int orc_build_index(const int64_t *u, const int64_t *v, int64_t h, int64_t n,
                    int64_t *degree, int64_t *sorted_v, int64_t *sorted_u,
                    int64_t *sorted_eid, int64_t *indptr) {
    int64_t *count = (int64_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int64_t));
    if (!count) return -1;
    memset(degree, 0, (size_t)n * sizeof(int64_t));
    for (int64_t i = 0; i < h; ++i) degree[u[i]] += 1;
    indptr[0] = 0;
    for (int64_t i = 0; i < n; ++i) indptr[i + 1] = indptr[i] + degree[i];
    for (int64_t i = 0; i < h; ++i) {
        int64_t p = indptr[u[i]] + count[u[i]];
        sorted_v[p] = v[i];
        sorted_eid[p] = i;
        sorted_u[p] = u[i];
        count[u[i]] += 1;
    }
    free(count);
    return 0;
}

/* paddle.geometric.send_u_recv (CPU contract) as called at pgl/graph.py:860,886.
 * x [n_src, D] row-major, COO edges in input order, out [n_out, D]. Single thread,
 * exactly the reference's loop structure. */
int orc_send_u_recv_f32(const float *x, const int64_t *src, const int64_t *dst, int64_t E,
                        int64_t n_out, int64_t D, int op, float *out) {
    memset(out, 0, (size_t)n_out * (size_t)D * sizeof(float));
    if (op == ORC_SUM || op == ORC_MEAN) {
        for (int64_t e = 0; e < E; ++e) {
            const float *xs = x + src[e] * D;
            float *od = out + dst[e] * D;
            for (int64_t k = 0; k < D; ++k) od[k] += xs[k];
        }
        if (op == ORC_MEAN) {
            int64_t *cnt = (int64_t *)calloc((size_t)(n_out > 0 ? n_out : 1), sizeof(int64_t));
            if (!cnt) return -1;
            for (int64_t e = 0; e < E; ++e) cnt[dst[e]] += 1;
            for (int64_t i = 0; i < n_out; ++i) {
                if (cnt[i] == 0) continue;
                float c = (float)cnt[i];
                float *od = out + i * D;
                for (int64_t k = 0; k < D; ++k) od[k] = od[k] / c;
            }
            free(cnt);
        }
    } else if (op == ORC_MAX || op == ORC_MIN) {
        unsigned char *seen = (unsigned char *)calloc((size_t)(n_out > 0 ? n_out : 1), 1);
        if (!seen) return -1;
        for (int64_t e = 0; e < E; ++e) {
            const float *xs = x + src[e] * D;
            float *od = out + dst[e] * D;
            if (!seen[dst[e]]) {
                seen[dst[e]] = 1;
                memcpy(od, xs, (size_t)D * sizeof(float));
            } else if (op == ORC_MAX) {
                for (int64_t k = 0; k < D; ++k) od[k] = xs[k] > od[k] ? xs[k] : od[k];
            } else {
                for (int64_t k = 0; k < D; ++k) od[k] = xs[k] < od[k] ? xs[k] : od[k];
            }
        }
        free(seen);
    } else {
        return -2;
    }
    return 0;
}

/* Same contract, threaded: rows of a dst-keyed CSR (pgl/graph.py:1319-1328 via
 * build_index) are independent, and inside one row the CSR keeps ascending edge id, so
 * every output row sees the SAME summation order as the sequential COO loop above --
 * results are bit-identical.  Used as the "all host threads" reference arm. */
int orc_send_u_recv_csr_f32(const float *x, const int64_t *indptr, const int64_t *sorted_src,
                            int64_t n_out, int64_t D, int op, float *out, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < n_out; ++r) {
        float *od = out + r * D;
        int64_t b = indptr[r], e = indptr[r + 1];
        for (int64_t k = 0; k < D; ++k) od[k] = 0.0f;
        if (b == e) continue;
        if (op == ORC_SUM || op == ORC_MEAN) {
            for (int64_t j = b; j < e; ++j) {
                const float *xs = x + sorted_src[j] * D;
                for (int64_t k = 0; k < D; ++k) od[k] += xs[k];
            }
            if (op == ORC_MEAN) {
                float c = (float)(e - b);
                for (int64_t k = 0; k < D; ++k) od[k] = od[k] / c;
            }
        } else {
            memcpy(od, x + sorted_src[b] * D, (size_t)D * sizeof(float));
            for (int64_t j = b + 1; j < e; ++j) {
                const float *xs = x + sorted_src[j] * D;
                if (op == ORC_MAX)
                    for (int64_t k = 0; k < D; ++k) od[k] = xs[k] > od[k] ? xs[k] : od[k];
                else
                    for (int64_t k = 0; k < D; ++k) od[k] = xs[k] < od[k] ? xs[k] : od[k];
            }
        }
    }
    return 0;
}

/* paddle.geometric.segment_{sum,mean,max,min} (CPU contract) as called at
 * pgl/math.py:36-42: ids sorted, out rows = ids[E-1]+1 (caller allocates), gaps stay 0. */
int orc_segment_pool_f32(const float *data, const int64_t *ids, int64_t E, int64_t D, int op,
                         float *out, int64_t n_out) {
    memset(out, 0, (size_t)n_out * (size_t)D * sizeof(float));
    int64_t b = 0;
    while (b < E) {
        int64_t e = b + 1;
        while (e < E && ids[e] == ids[b]) ++e;
        float *od = out + ids[b] * D;
        memcpy(od, data + b * D, (size_t)D * sizeof(float));
        for (int64_t j = b + 1; j < e; ++j) {
            const float *xs = data + j * D;
            if (op == ORC_SUM || op == ORC_MEAN)
                for (int64_t k = 0; k < D; ++k) od[k] += xs[k];
            else if (op == ORC_MAX)
                for (int64_t k = 0; k < D; ++k) od[k] = xs[k] > od[k] ? xs[k] : od[k];
            else
                for (int64_t k = 0; k < D; ++k) od[k] = xs[k] < od[k] ? xs[k] : od[k];
        }
        if (op == ORC_MEAN) {
            float c = (float)(e - b);
            for (int64_t k = 0; k < D; ++k) od[k] = od[k] / c;
        }
        b = e;
    }
    return 0;
}

/* `feature * norm` and `output * norm` of GCNConv.forward (pgl/nn/conv.py:242,250): an elementwise
 * multiply with the [N, 1] degree norm broadcast along the feature axis (Paddle's elementwise_mul CPU
 * contract: out[i, k] = x[i, k] * s[i], one rounding).  In place when out == x.  Rows are independent, so
 * the threaded form is bit-identical to the sequential one. */
int orc_scale_rows_f32(const float *x, const float *s, int64_t n, int64_t D, float *out, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)nthreads;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float si = s[i];
        const float *xi = x + i * D;
        float *oi = out + i * D;
        for (int64_t k = 0; k < D; ++k) oi[k] = xi[k] * si;
    }
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
