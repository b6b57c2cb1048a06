"""Lane-level emulation of the EXPERIMENTAL narrow-row kernel (csrc/spmm_stream.cu, spmm_narrow_kernel):
32 explicit lanes, the column-batch registers and their rotation, the per-lane source-lane arithmetic of
the shuffles, the cp.async ring (slots keep stale data until overwritten, consumed LAG groups after they
were issued), the sub / float4 lane mapping, the xor-shuffle row reduction and the writer lanes.
tests/test_narrow_model.py checks the row bookkeeping with the data path abstracted away; this file
checks the data path (which lane reads which column id, which ring slot, which 16 bytes) with the SAME
bookkeeping, again against the oracle.  Still no substitute for running the kernel."""
import numpy as np
import pytest

from oracle import oracle as O
from test_narrow_model import BIG, GRP, LAG, fixup, row_of_slot, task_plan

RING = 16
RG = RING // GRP


def shfl(vals, src_lanes):
    return [vals[s] for s in src_lanes]


def narrow_task_lanes(task, EPW, indptr, cols, x, scale, n_rows, E, start, first_row, out, partial, tail_row):
    D = x.shape[1]
    LPR = 32 // EPW
    SPG = GRP * EPW
    GPB = 32 // SPG
    SH = (LAG + GPB - 1) // GPB + 1
    lanes = range(32)
    sub = [l // LPR for l in lanes]
    li = [l % LPR for l in lanes]
    act = [li[l] * 4 < D for l in lanes]
    a, b = int(start[task]), int(start[task + 1])
    cnt = b - a
    row = int(first_row[task])
    tail = -1
    if cnt <= 0:
        tail_row[task] = tail
        return

    def rel(v):
        d = int(v) - a
        return -BIG if d < -BIG else (BIG if d > BIG else d)

    def xrow(c, l):   # the 16 bytes lane l copies from row c (inactive lanes alias li = 0)
        off = li[l] * 4 if act[l] else 0
        v = np.zeros(4, np.float32)
        seg = x[c, off:off + 4]
        v[:len(seg)] = seg
        return v

    st = {"row": row, "beg": rel(indptr[row]), "end": rel(indptr[row + 1]),
          "nxt": rel(indptr[row + 2]) if row + 2 <= n_rows else BIG, "head": rel(indptr[row]) < 0}
    acc = [np.zeros(4, np.float32) for _ in lanes]
    ring = [[np.full(4, np.nan, np.float32) for _ in lanes] for _ in range(RING)]   # stale = NaN

    def reduce_subs():
        o = LPR
        while o < 32:
            other = [acc[l ^ o].copy() for l in lanes]
            for l in lanes:
                acc[l] = acc[l] + other[l]
            o <<= 1

    def store(dst_row_vec, l):
        c0 = li[l] * 4
        dst_row_vec[c0:c0 + 4] = acc[l][:max(0, min(4, D - c0))] if D - c0 < 4 else acc[l]

    def finish_row():
        reduce_subs()
        writers = [l for l in lanes if act[l] and sub[l] == 0]
        if st["head"]:
            for l in writers:
                store(partial[2 * task], l)
            st["head"] = False
        else:
            deg = st["end"] - st["beg"]
            for l in writers:
                if deg == 0:
                    acc[l] = np.zeros(4, np.float32)
                store(out[st["row"]], l)
        st["row"] += 1
        st["beg"] = st["end"]
        st["end"] = st["nxt"]
        st["nxt"] = rel(indptr[st["row"] + 2]) if st["row"] + 2 <= n_rows else BIG
        if st["end"] == st["beg"] and st["row"] < n_rows:
            pos_abs = a + st["beg"]
            if pos_abs >= E:
                st["row"] = n_rows
                st["end"] = BIG
            else:
                st["row"] = row_of_slot(indptr, n_rows, pos_abs)
                st["end"] = rel(indptr[st["row"] + 1])
                st["nxt"] = rel(indptr[st["row"] + 2]) if st["row"] + 2 <= n_rows else BIG
        for l in lanes:
            acc[l] = np.zeros(4, np.float32)

    def load_col(batch):
        return [int(cols[a + batch * 32 + l]) if batch * 32 + l < cnt else 0 for l in lanes]

    def load_scale(colreg, base):
        return [np.float32(scale[colreg[l]]) if base + l < cnt else np.float32(1) for l in lanes]

    col_cur, col_nxt = load_col(0), load_col(1)
    sc_hist = [[np.float32(1)] * 32 for _ in range(SH)]
    sc_hist[0] = load_scale(col_cur, 0)
    ngroups = (cnt + SPG - 1) // SPG
    for g in range(ngroups + LAG):
        if g < ngroups:
            gsub = g % GPB
            if gsub == 0 and g > 0:
                col_cur = col_nxt
                col_nxt = load_col(g // GPB + 1)
                for i in range(SH - 1, 0, -1):
                    sc_hist[i] = sc_hist[i - 1]
                sc_hist[0] = load_scale(col_cur, g * SPG)
            rs = g % RG
            for k in range(GRP):
                src_lane = [(gsub * GRP + k) * EPW + sub[l] for l in lanes]
                c = shfl(col_cur, src_lane)
                for l in lanes:
                    my = (g * GRP + k) * EPW + sub[l]
                    if my < cnt:
                        ring[rs * GRP + k][l] = xrow(c[l], l)       # cp.async of this lane's 16 bytes
        if g < LAG:
            continue
        gc = g - LAG
        csub, crs = gc % GPB, gc % RG
        bcur = (g if g < ngroups else ngroups - 1) // GPB
        back = bcur - gc // GPB
        sc_reg = sc_hist[back] if 0 <= back < SH else None
        for k in range(GRP):
            s0 = (gc * GRP + k) * EPW
            if s0 >= cnt:
                break
            hi = min(s0 + EPW, cnt)
            v = [ring[crs * GRP + k][l] for l in lanes]                # lds128
            s = shfl(sc_reg, [(csub * GRP + k) * EPW + sub[l] for l in lanes])
            while st["end"] <= s0:
                finish_row()
            lo = s0
            while True:
                e = min(st["end"], hi)
                for l in lanes:
                    my = s0 + sub[l]
                    if lo <= my < e:
                        acc[l] = acc[l] + v[l] * s[l]
                if st["end"] < hi:
                    lo = st["end"]
                    finish_row()
                else:
                    break
    while st["row"] < n_rows and st["end"] <= cnt:
        finish_row()
    if st["row"] < n_rows and st["beg"] < cnt:
        reduce_subs()
        for l in lanes:
            if act[l] and sub[l] == 0:
                store(partial[2 * task if st["head"] else 2 * task + 1], l)
        if not st["head"]:
            tail = st["row"]
    tail_row[task] = tail


def lanes_spmm(edges, n, x, EPW, T, scale):
    deg, cols, _, _, indptr = O.build_index(edges[:, 1], edges[:, 0], n)
    E = len(edges)
    ntasks, first_row, start = task_plan(indptr, n, E, T, max(T, 1024))
    D = x.shape[1]
    out = np.full((n, D), np.nan, np.float32)
    partial = np.full((2 * ntasks, D), np.nan, np.float32)
    tail_row = np.full(ntasks, -7, np.int64)
    out[deg == 0] = 0.0
    for t in range(ntasks):
        narrow_task_lanes(t, EPW, indptr, cols, x, scale, n, E, start, first_row, out, partial, tail_row)
    fixup(ntasks, T, indptr, partial, tail_row, out, False)
    return out


@pytest.mark.parametrize("EPW,D", [(2, 64), (2, 36), (4, 32), (4, 20), (8, 16), (8, 4)])
def test_narrow_kernel_lane_emulation(EPW, D):
    rng = np.random.default_rng(950)
    hub = rng.integers(0, 40, (1500, 2))
    hub[:1150, 1] = 3
    gaps = O.chung_lu_edges(300, 900, exponent=0.8, seed=951)
    gaps[:, 1] = gaps[:, 1] // 9 * 9
    for name, n, edges in (("powerlaw", 150, O.chung_lu_edges(150, 1400, exponent=0.9, seed=952)),
                           ("hub", 40, hub), ("gaps", 300, gaps),
                           ("tiny", 5, np.array([[0, 1], [1, 2], [3, 4], [4, 1], [1, 0]]))):
        edges = np.asarray(edges, np.int64)
        x = rng.standard_normal((n, D)).astype(np.float32)
        scale = (rng.random(n) + 0.5).astype(np.float32)
        for T in (32, 2048):
            got = lanes_spmm(edges, n, x, EPW, T, scale)
            want = O.send_u_recv(x * scale[:, None], edges[:, 0], edges[:, 1], "sum")
            assert not np.isnan(got).any(), (name, EPW, D, T)
            err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-12)
            assert err <= 1e-5, (name, EPW, D, T, err)
