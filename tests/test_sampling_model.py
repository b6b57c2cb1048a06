"""CPU models of the EXPERIMENTAL sampling kernels (csrc/sampling.cu), statement by statement, checked
against the oracle -- the kernels have not run on hardware yet.

* reindex: rx_mark_seeds / rx_first_pos (atomicMin) / rx_flag / exclusive scan / rx_emit / rx_reset
  against oracle.reindex_graph (first-appearance order), including seeds that re-appear as neighbours,
  zero counts, duplicates, and the "table is clean again" post-condition.
* sample_fill: Floyd's subset selection with the kernel's hash (mix64) and multiply-high range
  reduction: k distinct positions inside the row, deterministic, every position about equally likely."""
import numpy as np

from oracle import oracle as O

EMPTY = np.iinfo(np.int64).max
M64 = (1 << 64) - 1


def reindex_model(x, nb, count, num_nodes, table):
    x, nb, count = (np.asarray(v, np.int64) for v in (x, nb, count))
    n, m = len(x), len(nb)
    offsets = np.concatenate([[0], np.cumsum(count)]).astype(np.int64)
    assert (table == EMPTY).all()
    for i in range(n):                                   # rx_mark_seeds_kernel
        table[x[i]] = i
    for p in range(m):                                   # rx_first_pos_kernel: atomicMin(table[nb[p]], n + p)
        table[nb[p]] = min(table[nb[p]], n + p)
    flag = np.array([1 if table[nb[p]] == n + p else 0 for p in range(m)], np.int64)   # rx_flag_kernel
    rank = np.concatenate([[0], np.cumsum(flag)[:-1]]).astype(np.int64) if m else np.zeros(0, np.int64)
    src = np.zeros(m, np.int64)
    dst = np.zeros(m, np.int64)
    out_nodes = np.full(n + m, -1, np.int64)
    num_out = n + (rank[m - 1] + flag[m - 1] if m > 0 else 0)
    for p in range(max(n, m)):                           # rx_emit_kernel
        if p < n:
            out_nodes[p] = x[p]
        if p >= m:
            continue
        v = nb[p]
        t = table[v]
        if t < n:
            src[p] = t
        else:
            idx = n + rank[t - n]
            src[p] = idx
            if flag[p]:
                out_nodes[idx] = v
        lo, hi = 0, n
        while hi - lo > 1:
            mid = (lo + hi) >> 1
            if offsets[mid] <= p:
                lo = mid
            else:
                hi = mid
        dst[p] = lo
    for p in range(max(n, m)):                           # rx_reset_kernel
        if p < n:
            table[x[p]] = EMPTY
        if p < m:
            table[nb[p]] = EMPTY
    return src, dst, out_nodes[:num_out]


def test_reindex_model_matches_oracle():
    rng = np.random.default_rng(1001)
    big = 500
    table = np.full(big, EMPTY, np.int64)
    cases = [([0, 1, 2], [8, 9, 0, 4, 7, 6, 7], [2, 3, 2])]
    for n in (1, 7, 60):
        x = rng.permutation(big)[:n]
        count = rng.integers(0, 6, n)
        count[rng.integers(0, n)] = 0
        nb = rng.integers(0, big, int(count.sum()))
        if len(nb) > 3:
            nb[1] = x[0]                                  # a seed re-appears as a neighbour
            nb[2] = nb[0]                                 # duplicate
        cases.append((x, nb, count))
    cases.append(([5], [], [0]))
    for x, nb, count in cases:
        src, dst, out = reindex_model(x, nb, count, big, table)
        ws, wd, wo = O.reindex_graph(x, nb, count)
        assert (src == ws).all() and (dst == wd).all() and (out == wo).all()
        assert (table == EMPTY).all()                     # clean for the next call


def mix64(z):
    z = (z + 0x9e3779b97f4a7c15) & M64
    z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & M64
    z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & M64
    return z ^ (z >> 31)


def floyd_model(deg, k, seed, i):
    base = mix64(seed ^ mix64(i))
    chosen = []
    for c in range(k):
        j = deg - k + c
        r = mix64((base + c) & M64)
        t = (r * (j + 1)) >> 64                          # __umul64hi
        if t in chosen:
            t = j
        chosen.append(t)
    return chosen


def test_floyd_subset_model():
    for deg, k in ((9, 8), (40, 5), (1000, 25), (26, 25)):
        hits = np.zeros(deg)
        trials = 3000
        for i in range(trials):
            ch = floyd_model(deg, k, seed=12345, i=i)
            assert len(set(ch)) == k and min(ch) >= 0 and max(ch) < deg
            assert ch == floyd_model(deg, k, seed=12345, i=i)      # pure function of (seed, i)
            hits[ch] += 1
        expect = trials * k / deg
        assert abs(hits.sum() - trials * k) < 1e-9
        # binomial spread: every position within 6 sigma of the uniform expectation
        sigma = np.sqrt(expect * (1 - k / deg)) + 1e-9
        assert np.abs(hits - expect).max() < 6 * sigma + 1, (deg, k, np.abs(hits - expect).max(), sigma)
