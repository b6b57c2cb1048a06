"""GPU parity of BiGraph (rectangular send/recv, reference pgl/bigraph.py:1051-1226) and of the
legacy COO helper graph_send_recv (reference pgl/utils/helper.py:163-210) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12) if a.size else 0.0


@pytest.fixture(scope="module")
def pgl():
    import pgl_b200
    return pgl_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def bip_edges(n_src, n_dst, e, seed):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n_src, e)
    dst = np.minimum((rng.random(e) ** 3 * n_dst).astype(np.int64), n_dst - 1)  # skewed in-degree
    return np.stack([src, dst], 1).astype(np.int64)


def test_bigraph_structure_and_send_recv(pgl):
    n_src, n_dst, e, d = 900, 300, 7000, 40
    edges = bip_edges(n_src, n_dst, e, 501)
    edges[:, 1][edges[:, 1] == 7] = 8  # an empty destination row
    x = np.random.default_rng(502).standard_normal((n_src, d)).astype(np.float32)
    g = pgl.BiGraph(edges, src_num_nodes=n_src, dst_num_nodes=n_dst).tensor()
    assert g.is_tensor() and int(g.src_num_nodes) == n_src and int(g.dst_num_nodes) == n_dst
    ref_dst = O.build_index(edges[:, 1], edges[:, 0], n_dst)
    ref_src = O.build_index(edges[:, 0], edges[:, 1], n_src)
    assert (g.indegree().cpu().numpy() == ref_dst[0]).all()
    assert (g.outdegree().cpu().numpy() == ref_src[0]).all()
    s, t, eid = g.sorted_edges("dst")
    assert (eid.cpu().numpy() == ref_dst[3]).all() and (t.cpu().numpy() == ref_dst[2]).all()
    assert (s.cpu().numpy() == ref_dst[1]).all()
    sub = np.array([5, 7, 0], np.int64)
    assert (g.indegree(dev(sub)).cpu().numpy() == ref_dst[0][sub]).all()
    for op in ("sum", "mean", "max", "min"):
        out = g.send_recv(dev(x), op).cpu().numpy()
        want = O.send_u_recv(x, edges[:, 0], edges[:, 1], op, out_size=n_dst)
        assert out.shape == (n_dst, d)
        if g.adj_dst_index.max_degree <= 1024:
            np.testing.assert_array_equal(out, want)
        else:
            assert rel_err(out, want) <= RTOL
    assert not out[7].any()
    with pytest.raises(ValueError):
        g.send_recv(dev(x[:10]), "sum")


def test_bigraph_udf_send_recv_both_modes(pgl):
    n_src, n_dst, e, d = 500, 200, 4000, 16
    edges = bip_edges(n_src, n_dst, e, 511)
    rng = np.random.default_rng(512)
    xs = rng.standard_normal((n_src, d)).astype(np.float32)
    xd = rng.standard_normal((n_dst, d)).astype(np.float32)
    ew = rng.random((e, 1)).astype(np.float32)
    g = pgl.BiGraph(dev(edges), src_num_nodes=n_src, dst_num_nodes=n_dst)

    def message(s, t, ee):
        return {"m": s["h"] * ee["w"] + t["h"], "c": s["h"]}

    msg = g.send(message, src_feat={"h": dev(xs)}, dst_feat={"h": dev(xd)}, edge_feat={"w": dev(ew)})
    msg_o = O.send(edges, message, src_feat={"h": xs}, dst_feat={"h": xd}, edge_feat={"w": ew})
    assert rel_err(msg["m"].cpu().numpy(), msg_o["m"]) <= 1e-6
    for mode, rows in (("dst", n_dst), ("src", n_src)):
        for name in ("reduce_sum", "reduce_mean", "reduce_max", "reduce_min"):
            fn = lambda m, name=name: getattr(m, name)(m["m"])  # noqa: E731
            out = g.recv(fn, msg, recv_mode=mode).cpu().numpy()
            want = O.recv(edges, rows, fn, msg_o, recv_mode=mode)
            assert out.shape == (rows, d) and rel_err(out, want) <= RTOL, (mode, name)
        out = g.recv(lambda m: m.reduce_sum(m["c"]), msg, recv_mode=mode).cpu().numpy()
        want = O.recv(edges, rows, lambda m: m.reduce_sum(m["c"]), msg_o, recv_mode=mode)
        assert rel_err(out, want) <= RTOL
    with pytest.raises(TypeError):
        g.recv(lambda m: m, "not a dict")
    with pytest.raises(TypeError):
        g.send(lambda s, t, ee: 1, src_feat={"h": dev(xs)})
    host = pgl.BiGraph(edges, src_num_nodes=n_src, dst_num_nodes=n_dst)
    with pytest.raises(ValueError):
        host.send(message, src_feat={"h": dev(xs)})


def test_bigraph_backward(pgl):
    n_src, n_dst, e, d = 300, 120, 2500, 8
    edges = bip_edges(n_src, n_dst, e, 521)
    g = pgl.BiGraph(dev(edges), src_num_nodes=n_src, dst_num_nodes=n_dst)
    src, dst = dev(edges[:, 0]), dev(edges[:, 1])
    x1 = torch.randn(n_src, d, device="cuda", requires_grad=True)
    x2 = x1.detach().clone().requires_grad_(True)
    go = torch.randn(n_dst, d, device="cuda")
    g.send_recv(x1, "sum").backward(go)
    ref = torch.zeros(n_dst, d, device="cuda").index_add_(0, dst, x2[src])
    ref.backward(go)
    assert rel_err(x1.grad.cpu().numpy(), x2.grad.cpu().numpy()) <= RTOL
    x1.grad = None
    x2.grad = None
    deg = torch.bincount(dst, minlength=n_dst).clamp(min=1).to(torch.float32)[:, None]
    g.send_recv(x1, "mean").backward(go)
    (torch.zeros(n_dst, d, device="cuda").index_add_(0, dst, x2[src]) / deg).backward(go)
    assert rel_err(x1.grad.cpu().numpy(), x2.grad.cpu().numpy()) <= RTOL


def test_graph_send_recv_helper(pgl):
    from pgl_b200.utils.helper import graph_send_recv
    n, e, d = 800, 9000, 32
    edges = O.chung_lu_edges(n, e, exponent=0.8, seed=531)
    x = np.random.default_rng(532).standard_normal((n, d)).astype(np.float32)
    ed = dev(edges)
    for op in ("sum", "mean", "max", "min"):
        out = graph_send_recv(dev(x), ed[:, 0], ed[:, 1], op).cpu().numpy()
        want = O.send_u_recv(x, edges[:, 0], edges[:, 1], op)
        assert rel_err(out, want) <= RTOL, op
    with pytest.raises(AssertionError):
        graph_send_recv(dev(x), ed[:, 0], ed[:, 1], "prod")
    x1 = dev(x).requires_grad_(True)
    x2 = dev(x).requires_grad_(True)
    go = torch.randn(n, d, device="cuda")
    graph_send_recv(x1, ed[:, 0], ed[:, 1], "sum").backward(go)
    torch.zeros(n, d, device="cuda").index_add_(0, ed[:, 1], x2[ed[:, 0]]).backward(go)
    assert rel_err(x1.grad.cpu().numpy(), x2.grad.cpu().numpy()) <= RTOL
