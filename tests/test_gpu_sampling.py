"""GPU tests of the neighbour-sampling path (csrc/sampling.cu, pgl_b200/sampling).  Written blind at the end
of round 1, validated on a B200 in round 2's first GPU call (gate dropped)."""

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pgl():
    import pgl_b200
    return pgl_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_reindex_graph_vs_oracle(pgl):
    from pgl_b200.sampling import reindex_graph
    src, dst, out = reindex_graph(dev(np.array([0, 1, 2])), dev(np.array([8, 9, 0, 4, 7, 6, 7])),
                                  dev(np.array([2, 3, 2])))
    assert src.tolist() == [3, 4, 0, 5, 6, 7, 6] and dst.tolist() == [0, 0, 1, 1, 1, 2, 2]
    assert out.tolist() == [0, 1, 2, 8, 9, 4, 7, 6]
    rng = np.random.default_rng(701)
    for n, big in ((1, 50), (300, 5000), (4000, 200000)):
        x = rng.permutation(big)[:n].astype(np.int64)
        count = rng.integers(0, 12, n)
        nb = rng.integers(0, big, int(count.sum())).astype(np.int64)
        for _ in range(2):  # twice: the lookup table must come back clean
            src, dst, out = reindex_graph(dev(x), dev(nb), dev(count), num_nodes=big)
            ws, wd, wo = O.reindex_graph(x, nb, count)
            assert (src.cpu().numpy() == ws).all() and (dst.cpu().numpy() == wd).all()
            assert (out.cpu().numpy() == wo).all()


def test_sample_neighbors_properties(pgl):
    from pgl_b200.sampling import sample_neighbors
    n, e = 3000, 120000
    edges = O.chung_lu_edges(n, e, exponent=0.9, seed=711)
    g = pgl.Graph(edges=edges, num_nodes=n)
    g.tensor()
    row, colptr = g.adj_dst_index._sorted_v, g.adj_dst_index._indptr
    indptr_h, row_h = colptr.cpu().numpy(), row.cpu().numpy()
    nodes = np.random.default_rng(712).permutation(n)[:700].astype(np.int64)
    for k in (-1, 0, 1, 5, 25, 40, 300):
        nb, cnt = sample_neighbors(row, colptr, dev(nodes), sample_size=k, seed=3)
        msg = O.check_sampled_neighbors(indptr_h, row_h, nodes, k, nb.cpu().numpy(), cnt.cpu().numpy())
        assert msg is None, (k, msg)
        nb2, _ = sample_neighbors(row, colptr, dev(nodes), sample_size=k, seed=3)
        assert torch.equal(nb, nb2)  # pure function of (seed, arguments)
    # eids: the sampled slot's edge id points back at an edge (neighbour -> node)
    nb, cnt, eids = sample_neighbors(row, colptr, dev(nodes), sample_size=5, eids=g.adj_dst_index._sorted_eid,
                                     return_eids=True, seed=9)
    ed = edges[eids.cpu().numpy()]
    assert (ed[:, 0] == nb.cpu().numpy()).all()
    assert (ed[:, 1] == np.repeat(nodes, cnt.cpu().numpy())).all()
    # uniformity on one hub: every neighbour slot is drawn about equally often
    hub = int(np.argmax(np.diff(indptr_h)))
    deg = int(indptr_h[hub + 1] - indptr_h[hub])
    k, trials = 8, 4000
    hits = np.zeros(deg)
    for s in range(trials // 100):
        _, _, eids = sample_neighbors(row, colptr, dev(np.full(100, hub)), sample_size=k,
                                      eids=None, return_eids=True, seed=100 + s)
        np.add.at(hits, eids.cpu().numpy() - indptr_h[hub], 1)
    expect = trials * k / deg
    assert abs(hits.mean() - expect) < 1e-9 and hits.std() < 6 * np.sqrt(expect) + 1


def test_neighbor_sampler_layers(pgl):
    from pgl_b200.sampling import NeighborSampler
    n, e, d = 2000, 40000, 16
    edges = O.chung_lu_edges(n, e, exponent=0.8, seed=721)
    g = pgl.Graph(edges=edges, num_nodes=n)
    g.tensor()
    sampler = NeighborSampler(g, [10, 5], seed=1)
    seeds = dev(np.arange(50, dtype=np.int64) * 7)
    graphs, in_nodes = sampler.sample_neighbors(seeds)
    assert len(graphs) == 2 and graphs[-1][1] == 50
    assert in_nodes[:50].tolist() == seeds.tolist()
    x = torch.randn(n, d, device="cuda")
    h = x[in_nodes]
    have = set(map(tuple, edges.tolist()))
    nodes_l = in_nodes
    for sub, n_target in graphs:
        assert int(sub.num_nodes) == int(h.shape[0])
        ee = sub.edges.cpu().numpy()
        glob = nodes_l.cpu().numpy()
        assert all((int(glob[s]), int(glob[t])) in have for s, t in ee[:200])
        assert ee[:, 1].max() < n_target
        h = sub.send_recv(h, "mean")[:n_target]
        nodes_l = nodes_l[:n_target]
    assert h.shape == (50, d)
