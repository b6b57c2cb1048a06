"""CPU restatement (numpy) of the reference's remaining conv layers -- the ones that are pure
compositions of the send/recv primitives.

TEST INFRASTRUCTURE ONLY (same rule as oracle.py: nothing in ``pgl_b200/`` imports it).

Each function follows the reference ``forward`` literally, in eval mode (every dropout is the
identity), through ``oracle.send`` / ``oracle.recv`` / ``oracle.send_u_recv`` ... ; weights are passed
in as numpy arrays with paddle's ``[in, out]`` Linear layout.  Parity status: the reference's
tests for these layers (tests/test_conv.py) check output shapes only, so the float results are
"parity unpinned" beyond the primitives' own golden vectors.
"""
import numpy as np

from . import oracle as O


def _lin(x, w, b=None):
    y = x @ w
    return y if b is None else y + b


def _layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + np.asarray(eps, x.dtype)) * g + b


def _l2_normalize(x, eps=1e-12):
    n = np.sqrt((x * x).sum(axis=1, keepdims=True))
    return x / np.maximum(n, np.asarray(eps, x.dtype))


def _norm_of(edges, n, dtype):
    return O.degree_norm(O.adj_dst_index(edges, n)[0], dtype)


def _prop(edges, x, norm):
    """feature * norm -> send_recv(sum) -> * norm  (the GCN propagation step)."""
    x = x * norm
    x = O.send_u_recv(x, edges[:, 0], edges[:, 1], "sum")
    return x * norm


def _with_self_loops(edges, n):
    """pgl/nn/conv.py:474-484 (APPNP) / :611-623 (GPRConv): drop self loops, prepend one per node."""
    idx = np.arange(n, dtype=np.int64)
    keep = edges[edges[:, 0] != edges[:, 1]]
    return np.concatenate([np.stack([idx, idx], 1), keep], axis=0)


def pinsage_conv(edges, n, nfeat, efeat, w_self, b_self, w_neigh, b_neigh, aggr_func="sum",
                 act=None):
    """pgl/nn/conv.py:152-186."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    msg = O.send(edges, lambda s, d, e: {"msg": s["h"] * e["w"]}, src_feat={"h": nfeat},
                 edge_feat={"w": efeat})
    neigh = O.recv(edges, n, lambda m: getattr(m, "reduce_%s" % aggr_func)(m["msg"]), msg)
    out = _lin(nfeat, w_self, b_self) + _lin(neigh, w_neigh, b_neigh)
    out = O._act(out, act)
    return _l2_normalize(out)


def gatv2_conv(edges, n, feature, w, b, attn, num_heads, hidden, concat=True, activation=None,
               negative_slope=0.2):
    """pgl/nn/conv.py:399-435 (eval mode)."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    f = _lin(feature, w, b).reshape(-1, num_heads, hidden)
    alpha = O.send_uv(f, f, edges[:, 0], edges[:, 1], "add")
    alpha = np.where(alpha >= 0, alpha, alpha * np.float32(negative_slope)).astype(f.dtype)
    alpha = np.sum(alpha * attn, axis=-1)
    alpha = O.edge_softmax(edges, n, alpha, "dst").reshape(-1, num_heads, 1)
    out = O.send_ue_recv(f, alpha, edges[:, 0], edges[:, 1], "mul", "sum")
    out = out.reshape(-1, num_heads * hidden) if concat else out.mean(axis=1)
    return O._act(out, activation)


def appnp(edges, n, feature, alpha=0.2, k_hop=10, self_loop=False, norm=None):
    """pgl/nn/conv.py:460-497."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    if self_loop:
        edges = _with_self_loops(edges, n)
    if norm is None:
        norm = _norm_of(edges, n, feature.dtype.type)
    h0 = feature
    a = feature.dtype.type(alpha)
    for _ in range(k_hop):
        feature = _prop(edges, feature, norm)
        feature = a * h0 + (1 - a) * feature
    return feature


def gpr_conv(edges, n, feature, w1, b1, w2, b2, temp, k_hop=10, activation="relu",
             self_loop=False, norm=None):
    """pgl/nn/conv.py:596-642 (eval mode)."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    if self_loop:
        edges = _with_self_loops(edges, n)
    feature = O._act(_lin(feature, w1, b1), activation)
    feature = _lin(feature, w2, b2)
    if norm is None:
        norm = _norm_of(edges, n, feature.dtype.type)
    hidden = feature * temp[0]
    for k in range(k_hop):
        feature = _prop(edges, feature, norm)
        hidden = hidden + temp[k + 1] * feature
    return hidden


def gcnii(edges, n, feature, ws, bs, activation=None, lambda_l=0.5, alpha=0.2, norm=None):
    """pgl/nn/conv.py:688-721 (eval mode); k_hop = len(ws)."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    if norm is None:
        norm = _norm_of(edges, n, feature.dtype.type)
    h0 = feature
    t = feature.dtype.type
    for i in range(len(ws)):
        beta = t(np.log(1.0 * lambda_l / (i + 1) + 1))
        feature = _prop(edges, feature, norm)
        feature = t(alpha) * h0 + (1 - t(alpha)) * feature
        trans = _lin(feature, ws[i], bs[i])
        feature = beta * trans + (1 - beta) * feature
        feature = O._act(feature, activation)
    return feature


def transformer_conv(edges, n, feature, p, num_heads, hidden, concat=True, edge_feat=None,
                     activation="relu"):
    """pgl/nn/conv.py:808-885 (eval mode).  p: dict with q/k/v (w, b), optional skip (w, b),
    gate (w, b), ln (g, b)."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    q = _lin(feature, *p["q"]).reshape(-1, num_heads, hidden)
    k = _lin(feature, *p["k"]).reshape(-1, num_heads, hidden)
    v = _lin(feature, *p["v"]).reshape(-1, num_heads, hidden)
    q = q / np.asarray(hidden ** 0.5, q.dtype)

    def send_attention(s, d, e):
        if "edge_feat" in e:
            alpha = d["q"] * (s["k"] + e["edge_feat"])
            vv = s["v"] + e["edge_feat"]
        else:
            alpha = d["q"] * s["k"]
            vv = s["v"]
        return {"alpha": alpha.sum(axis=-1), "v": vv}

    def reduce_attention(m):
        alpha = m.reduce_softmax(m["alpha"]).reshape(-1, num_heads, 1)
        f = m["v"] * alpha
        f = f.reshape(-1, num_heads * hidden) if concat else f.mean(axis=1)
        return m.reduce(f, pool_type="sum")

    if edge_feat is not None:
        ef = edge_feat.reshape(-1, num_heads, hidden)
        msg = O.send(edges, send_attention, src_feat={"k": k, "v": v}, dst_feat={"q": q},
                     edge_feat={"edge_feat": ef})
    else:
        msg = O.send(edges, send_attention, src_feat={"k": k, "v": v}, dst_feat={"q": q})
    out = O.recv(edges, n, reduce_attention, msg)
    if "skip" in p:
        skip = _lin(feature, *p["skip"])
        if "gate" in p:
            g = _lin(np.concatenate([skip, out, skip - out], axis=-1), *p["gate"])
            g = 1.0 / (1.0 + np.exp(-g))
            out = g * skip + (1 - g) * out
        else:
            out = skip + out
    if "ln" in p:
        out = _layer_norm(out, *p["ln"])
    return O._act(out.astype(feature.dtype), activation)


def gin_conv(edges, n, feature, w1, b1, w2, b2, ln_g, ln_b, epsilon=0.0, activation=None):
    """pgl/nn/conv.py:934-958."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    neigh = O.send_u_recv(feature, edges[:, 0], edges[:, 1], "sum")
    out = neigh + feature * feature.dtype.type(epsilon + 1.0)
    out = _layer_norm(_lin(out, w1, b1), ln_g, ln_b)
    out = O._act(out, activation)
    return _lin(out, w2, b2)


def rgcn_conv(edges_by_type, n, feat, weight, w_comp=None):
    """pgl/nn/conv.py:998-1024.  edges_by_type: ordered list of (etype, edges)."""
    if w_comp is not None:
        weight = np.einsum("rb,bio->rio", w_comp, weight)
    out = None
    for idx, (_, edges) in enumerate(edges_by_type):
        edges = np.asarray(edges, np.int64).reshape(-1, 2)
        h = feat @ weight[idx]
        h = O.send_u_recv(h, edges[:, 0], edges[:, 1], "mean")
        out = h if out is None else out + h
    return out


def sgc_conv(edges, n, feature, w, k_hop=2, bias=None, activation=None):
    """pgl/nn/conv.py:1064-1101."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    norm = _norm_of(edges, n, feature.dtype.type)
    for _ in range(k_hop):
        feature = _prop(edges, feature, norm)
    out = feature @ w
    if bias is not None:
        out = out + bias
    return O._act(out, activation)


def ssgc_conv(edges, n, feature, w, k_hop=16, alpha=0.05, bias=None, activation=None):
    """pgl/nn/conv.py:1151-1199, read as the paper's formula: sum_feature accumulates out of place
    (it does not alias the input)."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    norm = _norm_of(edges, n, feature.dtype.type)
    t = feature.dtype.type
    ori = feature
    total = feature
    for _ in range(k_hop):
        feature = _prop(edges, feature, norm)
        feature = (1 - t(alpha)) * feature
        total = total + feature
    feature = total / t(k_hop) + t(alpha) * ori
    out = feature @ w
    if bias is not None:
        out = out + bias
    return O._act(out, activation)


def ngcf_conv(edges, n, feature, w1, b1, w2, b2, negative_slope=0.2):
    """pgl/nn/conv.py:1229-1249."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    norm = _norm_of(edges, n, feature.dtype.type)
    neigh = O.send_u_recv(feature, edges[:, 0], edges[:, 1], "sum")
    out = (neigh + feature) * norm
    out = _lin(out, w1, b1) + _lin(feature * out, w2, b2)
    return np.where(out >= 0, out, out * np.float32(negative_slope)).astype(feature.dtype)


def lightgcn_conv(edges, n, feature):
    """pgl/nn/conv.py:1266-1284."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    return _prop(edges, feature, _norm_of(edges, n, feature.dtype.type))


def fa_conv(edges, n, feature, gate_w, gate_b):
    """pgl/nn/conv.py:1306-1341 (eval mode)."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    norm = _norm_of(edges, n, feature.dtype.type)

    def send_attention(s, d, e):
        h = np.concatenate([s["src"], d["dst"]], axis=1)
        h = np.tanh(_lin(h, gate_w, gate_b))
        return {"alpha": h * s["d"] * d["d"], "h": s["src"]}

    msg = O.send(edges, send_attention, src_feat={"src": feature, "d": norm},
                 dst_feat={"dst": feature, "d": norm})
    return O.recv(edges, n, lambda m: m.reduce(m["h"] * m["alpha"], pool_type="sum"), msg)
