"""CPU models of the scheduling logic added to the streaming kernels in round 2 (csrc/spmm_v5.inl, spmm_gat5.inl,
spmm_stream.cu): statement-by-statement restatements of the device code, checked against plain definitions.

  * row_of_slot_from   -- the galloping empty-row jump
  * the mbarrier ring of a persistent warp -- slot and wait parity carried from task to task through `gtot`
  * the device-side task queue -- every task handed out exactly once, in either order
"""
import numpy as np
import pytest


def row_of_slot_from(indptr, n_rows, frm, v):
    """csrc/spmm_stream.cu row_of_slot_from: the row r >= frm with indptr[r] <= v < indptr[r + 1], given
    indptr[frm] <= v < indptr[n_rows]."""
    lo, step = frm, 1
    loads = 0
    while True:
        hi = lo + step
        if hi >= n_rows:
            hi = n_rows
            break
        loads += 1
        if indptr[hi] > v:
            break
        lo = hi
        step <<= 1
    while hi - lo > 1:
        mid = (lo + hi) >> 1
        loads += 1
        if indptr[mid] > v:
            hi = mid
        else:
            lo = mid
    return lo, loads


@pytest.mark.parametrize("seed", range(6))
def test_galloping_row_search_equals_upper_bound(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 3000))
    # degrees with long runs of empty rows (RMAT-like tails) and a few hubs
    deg = rng.integers(0, 4, n) * (rng.random(n) < rng.choice([0.05, 0.3, 0.9]))
    deg[rng.integers(0, n, 3)] += rng.integers(50, 500, 3)
    if deg.sum() == 0:
        deg[rng.integers(0, n)] = 1
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    E = int(indptr[-1])
    worst = 0
    for v in rng.integers(0, E, 400):
        want = int(np.searchsorted(indptr, v, side="right")) - 1
        # every legal starting row: any row whose start is not beyond v
        for frm in {0, want, max(want - 1, 0), max(want - 7, 0), int(rng.integers(0, want + 1))}:
            got, loads = row_of_slot_from(indptr, n, frm, int(v))
            assert got == want and indptr[got] <= v < indptr[got + 1]
            if frm >= want - 7:
                worst = max(worst, loads)
    # the case the kernels hit: the jump starts at the empty row right behind the finished one
    for r in np.flatnonzero(deg[:-1] > 0)[:200]:
        pos = int(indptr[r + 1])
        if pos >= E or deg[r + 1] != 0:
            continue
        got, loads = row_of_slot_from(indptr, n, int(r + 1), pos)
        k = got - (r + 1)                                   # length of the empty run
        assert deg[got] > 0 and (deg[r + 1:got] == 0).all()
        assert loads <= 2 * max(int(np.ceil(np.log2(k + 1))), 1) + 2
    assert worst <= 8   # a nearby start costs a handful of loads, not log2(n)


@pytest.mark.parametrize("NG", [3, 4])
def test_mbarrier_ring_phases_carry_over_tasks(NG):
    """A persistent warp pushes the groups of successive tasks through one ring of NG mbarriers.  Group number
    G = gtot + g uses slot G % NG and waits with parity (G // NG) & 1; a barrier (arrival count 1) flips its phase
    once per use.  The wait of every group must name exactly the phase its own expect_tx / TMA completes."""
    rng = np.random.default_rng(NG)
    LAG = NG - 1
    for _ in range(50):
        completed = [0] * NG          # phases completed so far, per slot
        pending = {}                  # slot -> phase index its outstanding group will complete
        gtot = 0
        for ngroups in rng.integers(0, 23, 12):
            ngroups = int(ngroups)
            for g in range(ngroups + LAG):
                if g < ngroups:
                    s = (gtot + g) % NG
                    assert s not in pending, "slot reissued before its previous group was consumed"
                    pending[s] = completed[s]          # expect_tx arms the barrier's current phase
                if g >= LAG:
                    gc = g - LAG
                    if gc < ngroups:
                        s = (gtot + gc) % NG
                        parity = ((gtot + gc) // NG) & 1
                        phase = pending.pop(s)
                        assert parity == (phase & 1), "wait would watch the wrong phase"
                        completed[s] = phase + 1
            assert not pending                         # a task leaves the ring drained
            gtot += ngroups


@pytest.mark.parametrize("mode", [1, 2])
def test_task_queue_hands_every_task_out_once(mode):
    rng = np.random.default_rng(mode)
    for ntasks in (0, 1, 7, 296 * 13, 50001):
        warps = 296 * 13
        counter = 0
        seen = np.zeros(ntasks, dtype=np.int32)
        alive = list(range(warps))
        while alive:
            nxt = []
            for w in rng.permutation(alive):          # any interleaving of the warps' atomicAdd
                t = counter
                counter += 1
                if t >= ntasks:
                    continue                           # this warp leaves the loop
                task = ntasks - 1 - t if mode == 2 else t
                seen[task] += 1
                nxt.append(w)
            alive = nxt
        assert (seen == 1).all()
        assert counter >= ntasks                       # re-armed to 0 by the next launch's task_plan_kernel


def test_geometry_and_queue_defaults():
    """mirror of v5_geo() / dyn_mode(): groups of 8 from 12 slots per row on; ascending queue above 8 tasks per resident
    warp, descending below (148 SMs x 2 CTAs x 13 warps)."""
    def geo(E, n):
        return 2 if (n > 0 and E >= 12 * n) else 0

    def dyn(ntasks, sms=148):
        return 1 if ntasks >= 8 * sms * 2 * 13 else 2
    assert geo(100_000_000, 10_000_000) == 0 and geo(12_500_044, 1_184_431) == 0      # cfg5 and its 8 x 1 shard
    assert geo(123_718_280, 2_449_029) == 2                                            # cfg4
    assert dyn(48829) == 1 and dyn(16276) == 2


# ---------------------------------------------------------------- csrc/localgraph.cu: the halo plan as flags + prefix sums
def halo_plan_model(edges, n, lo, hi, offsets):
    """halo_mark_kernel -> two inclusive scans -> halo_fill_edges_kernel / halo_fill_nodes_kernel /
    halo_recv_counts_kernel, statement by statement in numpy."""
    src, dst = edges[:, 0], edges[:, 1]
    mine = ((dst >= lo) & (dst < hi)).astype(np.int64)
    mark = np.zeros(n, np.int64)
    remote = (src < lo) | (src >= hi)
    mark[src[(mine == 1) & remote]] = 1
    mine_incl, mark_incl = np.cumsum(mine), np.cumsum(mark)
    e_loc = int(mine_incl[-1]) if len(mine_incl) else 0
    n_halo = int(mark_incl[-1]) if n else 0
    eid = np.empty(e_loc, np.int64)
    dl = np.empty(e_loc, np.int64)
    cl = np.empty(e_loc, np.int64)
    for e in range(len(edges)):
        prev = mine_incl[e - 1] if e else 0
        if mine_incl[e] == prev:
            continue
        p = mine_incl[e] - 1
        eid[p] = e
        dl[p] = dst[e] - lo
        cl[p] = (hi - lo) + mark_incl[src[e]] - 1 if remote[e] else src[e] - lo
    halo = np.empty(n_halo, np.int64)
    for i in range(n):
        prev = mark_incl[i - 1] if i else 0
        if mark_incl[i] != prev:
            halo[mark_incl[i] - 1] = i

    def before(x):
        x = min(max(x, 0), n)
        return int(mark_incl[x - 1]) if x > 0 else 0
    recv = [before(offsets[p + 1]) - before(offsets[p]) for p in range(len(offsets) - 1)]
    return eid, dl, cl, halo, recv


@pytest.mark.parametrize("seed", range(4))
def test_halo_plan_model_equals_the_torch_construction(seed):
    import torch
    from pgl_b200.distributed.halo import HaloPlan, block_offsets
    rng = np.random.default_rng(seed)
    n, e, k = int(rng.integers(20, 400)), int(rng.integers(1, 3000)), int(rng.integers(1, 6))
    edges = rng.integers(0, n, (e, 2)).astype(np.int64)
    offsets = block_offsets(n, k)
    if k > 2 and seed % 2:
        offsets[2] = offsets[1]          # an empty part
    for r in range(k):
        lo, hi = offsets[r], offsets[r + 1]
        eid, dl, cl, halo, recv = halo_plan_model(edges, n, lo, hi, offsets)
        # world 1 per call (no collective): the torch-op construction the gloo tests use
        src, dst = edges[:, 0], edges[:, 1]
        w_eid = np.nonzero((dst >= lo) & (dst < hi))[0]
        s = src[w_eid]
        rem = (s < lo) | (s >= hi)
        w_halo = np.unique(s[rem])
        w_cl = np.where(rem, np.searchsorted(w_halo, s) + (hi - lo), s - lo)
        assert (eid == w_eid).all() and (dl == dst[w_eid] - lo).all() and (cl == w_cl).all() and (halo == w_halo).all()
        assert recv == [int(((w_halo >= offsets[p]) & (w_halo < offsets[p + 1])).sum()) for p in range(k)]
    plan = HaloPlan.build(torch.from_numpy(edges), n, [0, n], 0, 1)
    eid, dl, cl, halo, recv = halo_plan_model(edges, n, 0, n, [0, n])
    assert (plan.eid.numpy() == eid).all() and (plan.col_local.numpy() == cl).all() and plan.n_halo == 0 == len(halo)
