"""CPU model of the narrow-row kernel's control flow (csrc/spmm_narrow2.inl + the plan kernels in
csrc/csr_build.cu + the shared fix-up of csrc/spmm_stream.cu), statement by statement, checked against the oracle.

What is modelled: the plan (head flags in bit 30, nz_row, blk_k), the per-sub-warp walk over 32 consecutive slots
(flag -> first one parks the head piece, later ones write a complete row), the serial stitch of the EPW ranges of a
chunk with the carried open row, the task-head / task-tail partials and the fix-up that sums them.  Features are
small integers stored as float32, so every summation order gives the same bits and equality is exact."""
import numpy as np
import pytest

from oracle import oracle as O

HEAD = 1 << 30


def build_plan(indptr, cols):
    """plan_cols_kernel / plan_rows_kernel / plan_blocks_kernel."""
    n = len(indptr) - 1
    e = len(cols)
    flag = (indptr[1:] > indptr[:-1]).astype(np.int64)
    rank = np.cumsum(flag) - flag                      # exclusive scan
    plan = cols.astype(np.uint32).copy()
    k_total = int(flag.sum())
    nz_row = np.zeros(n + 2, np.int32)
    for r in range(n):
        if indptr[r + 1] > indptr[r]:
            nz_row[rank[r]] = r
            plan[indptr[r]] |= HEAD
    last = int(np.searchsorted(indptr, e - 1, side="right") - 1) if k_total else 0
    nz_row[k_total] = last
    nz_row[k_total + 1] = last
    nblk = (e + 31) // 32
    blk_k = np.zeros(nblk, np.int32)
    blk_k[0] = -1
    for b in range(1, nblk):
        v = b * 32 - 1
        r = int(np.searchsorted(indptr, v, side="right") - 1)
        blk_k[b] = rank[r]
    return plan, nz_row, blk_k


def run_model(indptr, cols, x, n_rows, lpr, T, scale_src=None, scale_dst=None, mean=False):
    e = len(cols)
    epw = 32 // lpr
    chunk_sz = epw * 32
    assert T % chunk_sz == 0
    plan, nz_row, blk_k = build_plan(indptr, cols)
    ntasks = (e + T - 1) // T
    out = np.full(n_rows, np.nan, np.float32)
    # empty_rows_kernel
    for r in range(n_rows):
        if indptr[r + 1] == indptr[r]:
            out[r] = 0.0
    partial = np.zeros(2 * ntasks, np.float32)
    tail_row = np.full(ntasks, -1, np.int64)

    def write_out(row, v):
        if mean:
            v = np.float32(v / np.float32(indptr[row + 1] - indptr[row]))
        if scale_dst is not None:
            v = np.float32(v * scale_dst[row])
        assert np.isnan(out[row]), "row %d written twice" % row
        out[row] = v

    for task in range(ntasks):
        t_beg = task * T
        t_end = min(e, t_beg + T)
        carry, carry_mode, carry_row = np.float32(0), 1, -1
        for chunk in range(t_beg, t_end, chunk_sz):
            subs = []
            for sub in range(epw):
                rb = chunk + sub * 32
                nvalid = max(0, min(32, t_end - rb))
                k = int(blk_k[rb >> 5]) if nvalid > 0 else -1
                row_cur = int(nz_row[k]) if k >= 0 else 0
                row_nxt = int(nz_row[k + 1])
                row_head = row_cur
                acc = np.float32(0)
                head_acc = np.float32(0)
                flagged = False
                for j in range(32):
                    c = int(plan[rb + j]) if j < nvalid else 0
                    v = np.float32(0)
                    if j < nvalid:
                        cid = c & (HEAD - 1)
                        v = x[cid]
                        if scale_src is not None:
                            v = np.float32(v * scale_src[cid])
                    if c & HEAD:
                        if not flagged:
                            head_acc = acc
                            flagged = True
                        else:
                            write_out(row_cur, acc)
                        acc = np.float32(0)
                        k += 1
                        row_cur = row_nxt
                        row_nxt = int(nz_row[k + 1])
                    acc = np.float32(acc + v)
                if not flagged:
                    head_acc = acc
                subs.append((head_acc, flagged, acc, row_cur, row_head))
            cin, cmode, crow = carry, carry_mode, carry_row
            for sidx in range(epw):
                head_acc, flagged, tail_acc, tail_r, row_head = subs[sidx]
                tot = np.float32(cin + head_acc)
                if flagged:
                    if cmode:
                        partial[2 * task] = tot
                    else:
                        assert row_head == crow
                        write_out(row_head, tot)
                    cin, cmode, crow = tail_acc, 0, tail_r
                else:
                    cin = tot
            carry, carry_mode, carry_row = cin, cmode, crow
        if carry_mode:
            partial[2 * task] = carry
        else:
            partial[2 * task + 1] = carry
            tail_row[task] = carry_row
    # spmm_stream_fixup_kernel<1, 0>
    for t in range(ntasks):
        r = tail_row[t]
        if r < 0:
            continue
        e_r = indptr[r + 1]
        acc = partial[2 * t + 1]
        u_end = (e_r + T - 1) // T
        for u in range(t + 1, u_end):
            acc = np.float32(acc + partial[2 * u])
        write_out(r, acc)
    return out


def graphs():
    rng = np.random.default_rng(5)
    yield "powerlaw", 300, O.chung_lu_edges(300, 4000, exponent=0.9, seed=801)
    yield "uniform", 200, rng.integers(0, 200, (3000, 2))
    e = O.chung_lu_edges(500, 2000, exponent=0.8, seed=803)
    e[:, 1] = e[:, 1] // 7 * 7
    yield "gaps", 500, e
    yield "tiny", 5, np.array([[0, 1], [1, 2], [3, 4], [4, 1], [1, 0]])
    hub = rng.integers(0, 40, (3000, 2))
    hub[:2500, 1] = 7
    yield "hub", 40, hub
    one = rng.integers(0, 50, (777, 2))
    one[:, 1] = 49
    yield "single_last_row", 50, one
    aligned = np.stack([rng.integers(0, 64, 64 * 32), np.repeat(np.arange(64), 32)], 1)   # every row exactly 32 slots
    yield "aligned32", 64, aligned


@pytest.mark.parametrize("lpr", [4, 8, 16])
def test_model_equals_oracle(lpr):
    chunk = (32 // lpr) * 32
    for name, n, edges in graphs():
        edges = np.asarray(edges, np.int64)
        deg, sv, su, se, ip = O.adj_dst_index(edges, n)
        x = np.random.default_rng(11).integers(-8, 9, n).astype(np.float32)
        want = O.send_u_recv(x.reshape(-1, 1), edges[:, 0], edges[:, 1], "sum").reshape(-1)
        for T in (chunk, 2 * chunk, 8 * chunk):
            got = run_model(ip, sv, x, n, lpr, T)
            assert not np.isnan(got).any(), (name, lpr, T)
            assert np.array_equal(got, want), (name, lpr, T)


def test_model_scales_and_mean():
    n = 300
    edges = O.chung_lu_edges(n, 4000, exponent=0.9, seed=12)
    deg, sv, su, se, ip = O.adj_dst_index(edges, n)
    rng = np.random.default_rng(13)
    x = rng.integers(-8, 9, n).astype(np.float32)
    s = (2.0 ** rng.integers(-2, 3, n)).astype(np.float32)     # powers of two: products stay exact
    got = run_model(ip, sv, x, n, 4, 256, scale_src=s, scale_dst=s)
    want = O.send_u_recv((x * s).reshape(-1, 1), edges[:, 0], edges[:, 1], "sum").reshape(-1) * s
    assert np.array_equal(got, want)
    got = run_model(ip, sv, x, n, 8, 128, mean=True)
    want = O.send_u_recv(x.reshape(-1, 1), edges[:, 0], edges[:, 1], "mean").reshape(-1)
    assert np.allclose(got, want, rtol=1e-6, atol=0)
