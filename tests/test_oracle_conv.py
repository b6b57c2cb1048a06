"""CPU checks of oracle/oracle_conv.py (the numpy restatement of the composed conv layers):
identities that tie it to the pinned primitives of oracle.py, so the GPU parity tests of
tests/test_gpu_convs.py compare against something that was itself cross-checked."""
import numpy as np

from oracle import oracle as O
from oracle import oracle_conv as OC


def _setup(n=60, e=400, d=6, seed=5):
    edges = O.chung_lu_edges(n, e, exponent=0.7, seed=seed)
    x = np.random.default_rng(seed + 1).standard_normal((n, d)).astype(np.float32)
    return edges, x


def test_propagation_family_reduces_to_gcn():
    n, d = 60, 6
    edges, x = _setup(n=n, d=d)
    eye, zero = np.eye(d, dtype=np.float32), np.zeros(d, np.float32)
    gcn = O.gcn_conv(edges, n, x, eye, zero)
    np.testing.assert_allclose(OC.lightgcn_conv(edges, n, x), gcn, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(OC.sgc_conv(edges, n, x, eye, k_hop=1), gcn, rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(OC.appnp(edges, n, x, k_hop=0), x)
    np.testing.assert_allclose(OC.appnp(edges, n, x, alpha=0.0, k_hop=1), gcn, rtol=1e-6, atol=1e-7)
    # GPR with gamma = e_1 and identity MLP is one propagation step
    temp = np.array([0, 1, 0], np.float32)
    out = OC.gpr_conv(edges, n, np.abs(x), eye, zero, eye, zero, temp, k_hop=2)
    np.testing.assert_allclose(out, O.gcn_conv(edges, n, np.abs(x), eye, zero), rtol=1e-6, atol=1e-7)
    # SSGC: k_hop = 1 -> x + (1-a) P x + a x
    a = np.float32(0.05)
    want = x + (1 - a) * gcn + a * x
    np.testing.assert_allclose(OC.ssgc_conv(edges, n, x, eye, k_hop=1, alpha=0.05), want, rtol=1e-6,
                               atol=1e-6)


def test_self_loop_rewrite():
    edges = np.array([[0, 1], [1, 1], [2, 0], [2, 2], [1, 2]], np.int64)
    got = OC._with_self_loops(edges, 3)
    want = np.array([[0, 0], [1, 1], [2, 2], [0, 1], [2, 0], [1, 2]], np.int64)
    np.testing.assert_array_equal(got, want)


def test_udf_layers_match_their_fused_spelling():
    n, d = 60, 6
    edges, x = _setup(n=n, d=d, seed=9)
    rng = np.random.default_rng(11)
    src, dst = edges[:, 0], edges[:, 1]
    # PinSage: send(h*w) -> recv(reduce) == send_ue_recv(mul)
    w_e = rng.random((edges.shape[0], 1)).astype(np.float32)
    eye, zero = np.eye(d, dtype=np.float32), np.zeros(d, np.float32)
    for aggr in ("sum", "mean", "max", "min"):
        neigh = O.send_ue_recv(x, w_e, src, dst, "mul", aggr)
        want = OC._l2_normalize(x + neigh)
        got = OC.pinsage_conv(edges, n, x, w_e, eye, zero, eye, zero, aggr)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    # FAConv: the gate on the concat splits into two projections
    gw = rng.standard_normal((2 * d, 1)).astype(np.float32)
    gb = rng.standard_normal(1).astype(np.float32)
    norm = OC._norm_of(edges, n, np.float32)
    gate = np.tanh(O.send_uv(x @ gw[:d], x @ gw[d:] + gb, src, dst, "add"))
    alpha = gate * O.send_uv(norm, norm, src, dst, "mul")
    want = O.send_ue_recv(x, alpha, src, dst, "mul", "sum")
    np.testing.assert_allclose(OC.fa_conv(edges, n, x, gw, gb), want, rtol=1e-5, atol=1e-6)
    # Transformer (no edge feature): send/recv pair == send_uv(mul) + edge_softmax + send_ue_recv
    H, Dh = 2, 3
    p = {k: (rng.standard_normal((d, H * Dh)).astype(np.float32) * 0.3,
             rng.standard_normal(H * Dh).astype(np.float32) * 0.3) for k in ("q", "k", "v")}
    got = OC.transformer_conv(edges, n, x, p, H, Dh, concat=True, activation=None)
    q = (x @ p["q"][0] + p["q"][1]).reshape(-1, H, Dh) / np.float32(Dh ** 0.5)
    k = (x @ p["k"][0] + p["k"][1]).reshape(-1, H, Dh)
    v = (x @ p["v"][0] + p["v"][1]).reshape(-1, H, Dh)
    a = O.send_uv(k, q, src, dst, "mul").sum(-1)
    a = O.edge_softmax(edges, n, a, "dst").reshape(-1, H, 1)
    want = O.send_ue_recv(v, a, src, dst, "mul", "sum").reshape(n, H * Dh)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_gin_rgcn_ngcf_shapes_and_values():
    n, d = 60, 6
    edges, x = _setup(n=n, d=d, seed=13)
    eye, zero, one = np.eye(d, dtype=np.float32), np.zeros(d, np.float32), np.ones(d, np.float32)
    out = OC.gin_conv(edges, n, x, eye, zero, eye, zero, one, zero, epsilon=0.5)
    pre = O.send_u_recv(x, edges[:, 0], edges[:, 1], "sum") + np.float32(1.5) * x
    np.testing.assert_allclose(out, OC._layer_norm(pre, one, zero), rtol=1e-6, atol=1e-6)
    w = np.stack([eye, 2 * eye])
    out = OC.rgcn_conv([("a", edges), ("b", edges[:100])], n, x, w)
    want = O.send_u_recv(x, edges[:, 0], edges[:, 1], "mean") + \
        O.send_u_recv(2 * x, edges[:100, 0], edges[:100, 1], "mean")
    np.testing.assert_allclose(out, want, rtol=1e-6, atol=1e-6)
    out = OC.ngcf_conv(edges, n, x, eye, zero, np.zeros((d, d), np.float32), zero)
    norm = OC._norm_of(edges, n, np.float32)
    pre = (O.send_u_recv(x, edges[:, 0], edges[:, 1], "sum") + x) * norm
    np.testing.assert_allclose(out, np.where(pre >= 0, pre, 0.2 * pre), rtol=1e-6, atol=1e-6)
