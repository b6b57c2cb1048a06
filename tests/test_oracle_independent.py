"""The float half of the oracle cannot be checked against Paddle here (not installable), so besides the
reference's golden vectors it is cross-checked against INDEPENDENT CPU implementations of the same
published contracts: scipy.sparse for sum / mean, torch.scatter_reduce (amax / amin, include_self=False)
for max / min, a per-segment torch.softmax for segment_softmax / edge_softmax, and the plain-C loop of
oracle/oracle_c.c.  Two restatements written against the same contract by different routes agreeing is
weaker than running Paddle, and stronger than one restatement alone."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import oracle as O


def _case(seed, n=300, e=4000, d=7):
    edges = O.chung_lu_edges(n, e, exponent=0.8, seed=seed)
    edges[:, 1] = np.where(edges[:, 1] % 11 == 0, (edges[:, 1] + 1) % n, edges[:, 1])  # some empty rows
    x = np.random.default_rng(seed + 1).standard_normal((n, d)).astype(np.float32)
    return edges, x, n


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_send_u_recv_against_scipy_and_torch(seed):
    edges, x, n = _case(seed)
    src, dst = edges[:, 0], edges[:, 1]
    a = sp.csr_matrix((np.ones(len(src), np.float64), (dst, src)), shape=(n, n))
    want_sum = a @ x.astype(np.float64)
    got = O.send_u_recv(x, src, dst, "sum")
    np.testing.assert_allclose(got, want_sum, rtol=1e-5, atol=1e-5)
    deg = np.asarray(a.sum(1)).reshape(-1)
    want_mean = want_sum / np.maximum(deg, 1)[:, None]
    np.testing.assert_allclose(O.send_u_recv(x, src, dst, "mean"), want_mean, rtol=1e-5, atol=1e-5)
    xt, idx = torch.from_numpy(x), torch.from_numpy(dst)[:, None].expand(-1, x.shape[1])
    for op, red in (("max", "amax"), ("min", "amin")):
        ref = torch.zeros(n, x.shape[1]).scatter_reduce(0, idx, xt[torch.from_numpy(src)], red,
                                                        include_self=False).numpy()
        np.testing.assert_array_equal(O.send_u_recv(x, src, dst, op), ref)   # rows without edges stay 0
    # out_size: more rows than nodes pads zeros
    big = O.send_u_recv(x, src, dst, "sum", out_size=n + 5)
    assert big.shape == (n + 5, x.shape[1]) and not big[n:].any()
    np.testing.assert_array_equal(big[:n], got)


@pytest.mark.parametrize("seed", [4, 5])
def test_send_ue_recv_and_send_uv_against_torch(seed):
    edges, x, n = _case(seed, d=6)
    src, dst = torch.from_numpy(edges[:, 0]), torch.from_numpy(edges[:, 1])
    rng = np.random.default_rng(seed + 7)
    y = (rng.random((len(edges), 6)) + 0.5).astype(np.float32)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    ops = {"add": xt[src] + yt, "sub": xt[src] - yt, "mul": xt[src] * yt, "div": xt[src] / yt}
    for name, msg in ops.items():
        ref = torch.zeros(n, 6).index_add_(0, dst, msg).numpy()
        np.testing.assert_allclose(O.send_ue_recv(x, y, edges[:, 0], edges[:, 1], name, "sum"), ref,
                                   rtol=1e-5, atol=1e-5)
    ys = y[:, :1]                                         # broadcast operand [E, 1]
    ref = torch.zeros(n, 6).index_add_(0, dst, xt[src] * torch.from_numpy(ys)).numpy()
    np.testing.assert_allclose(O.send_ue_recv(x, ys, edges[:, 0], edges[:, 1], "mul", "sum"), ref,
                               rtol=1e-5, atol=1e-5)
    z = rng.standard_normal((n, 6)).astype(np.float32)
    for name, fn in (("add", torch.add), ("sub", torch.sub), ("mul", torch.mul), ("div", torch.div)):
        ref = fn(xt[src], torch.from_numpy(z)[dst]).numpy()
        np.testing.assert_allclose(O.send_uv(x, z, edges[:, 0], edges[:, 1], name), ref, rtol=1e-6, atol=1e-6)


def test_softmax_family_against_torch():
    edges, _, n = _case(6)
    rng = np.random.default_rng(13)
    logits = (rng.standard_normal((len(edges), 3)) * 4).astype(np.float32)
    out = O.edge_softmax(edges, n, logits, "dst")
    dst = edges[:, 1]
    for v in np.unique(dst)[:50]:
        m = dst == v
        ref = torch.softmax(torch.from_numpy(logits[m]), dim=0).numpy()
        np.testing.assert_allclose(out[m], ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.add.reduceat(out[np.argsort(dst, kind="stable")],
                                               np.unique(np.sort(dst), return_index=True)[1]), 1.0, rtol=1e-5)
    seg = np.sort(rng.integers(0, 40, 500))
    data = rng.standard_normal((500, 2)).astype(np.float32)
    sm = O.segment_softmax(data, seg)
    for v in np.unique(seg):
        ref = torch.softmax(torch.from_numpy(data[seg == v]), dim=0).numpy()
        np.testing.assert_allclose(sm[seg == v], ref, rtol=1e-5, atol=1e-6)


def test_c_restatement_matches_numpy_restatement_bit_for_bit():
    from oracle import build as obuild
    import ctypes
    lib = ctypes.CDLL(obuild.build_oracle_c())
    edges, x, n = _case(8, n=500, e=9000, d=16)
    src = np.ascontiguousarray(edges[:, 0])
    dst = np.ascontiguousarray(edges[:, 1])
    out = np.full((n, 16), np.nan, np.float32)
    P = ctypes.c_void_p
    lib.orc_send_u_recv_f32.restype = ctypes.c_int
    for op_id, op in ((0, "sum"), (1, "mean"), (2, "max"), (3, "min")):
        rc = lib.orc_send_u_recv_f32(P(x.ctypes.data), P(src.ctypes.data), P(dst.ctypes.data),
                                     ctypes.c_int64(len(src)), ctypes.c_int64(n), ctypes.c_int64(16),
                                     ctypes.c_int(op_id), P(out.ctypes.data))
        assert rc == 0
        np.testing.assert_array_equal(out, O.send_u_recv(x, src, dst, op), err_msg=op)
