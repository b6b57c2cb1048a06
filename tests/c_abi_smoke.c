/* Plain-C consumer of include/pglb.h: what a cgo / JNI / ctypes-free binding on the reference side would
 * do.  Host entry points only (no GPU needed): version, error convention, workspace query, and the
 * twin of the reference's Cython build_index (pgl/graph_kernel.pyx:59-88) on the 5-edge graph of
 * tests/test_graph.py:337-357.  Built and run by tests/test_abi_host.py. */
#include <stdio.h>
#include <string.h>

#include "pglb.h"

int main(void) {
    if (pglb_version() != 100) return 1;
    /* edges (src, dst): (0,1) (1,2) (3,4) (4,1) (1,0); index keyed by dst (u = dst, v = src) */
    const int64_t src[5] = {0, 1, 3, 4, 1}, dst[5] = {1, 2, 4, 1, 0};
    int64_t degree[5], indptr[6], su[5], sv[5], se[5];
    int rc = pglb_build_index_host(dst, 1, src, 1, 5, 5, degree, indptr, su, sv, se);
    if (rc != PGLB_OK) { printf("rc %d %s\n", rc, pglb_last_error()); return 2; }
    const int64_t want_deg[5] = {1, 2, 1, 0, 1}, want_ptr[6] = {0, 1, 3, 4, 4, 5}, want_eid[5] = {4, 0, 3, 1, 2};
    if (memcmp(degree, want_deg, sizeof want_deg) || memcmp(indptr, want_ptr, sizeof want_ptr) ||
        memcmp(se, want_eid, sizeof want_eid)) return 3;
    size_t ws = 0;
    if (pglb_csr_build_ws(100, 10, &ws) != PGLB_OK || ws == 0) return 4;
    if (pglb_csr_build_ws(-1, 10, &ws) != PGLB_EINVAL) return 5;
    if (strstr(pglb_last_error(), "negative") == NULL) return 6;
    /* an id outside [0, N) is rejected, not dereferenced */
    const int64_t bad[1] = {7}, zero[1] = {0};
    int64_t d2[3], p2[4], a2[1], b2[1], c2[1];
    if (pglb_build_index_host(bad, 1, zero, 1, 1, 3, d2, p2, a2, b2, c2) >= 0) return 7;
    printf("C_ABI_OK indegree=[%lld %lld %lld %lld %lld]\n", (long long)degree[0], (long long)degree[1],
           (long long)degree[2], (long long)degree[3], (long long)degree[4]);
    return 0;
}
