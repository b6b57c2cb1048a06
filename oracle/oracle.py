"""CPU restatement (numpy) of PGL's send/recv message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``pgl_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker / reported baseline.

Parity status
-------------
* Integer half (``build_index``, degrees, segment ids, dst-sharding): pinned
  bit-exactly against the reference's own compiled Cython module
  (``oracle/_ref/graph_kernel*.so`` built from ``/root/reference/pgl/graph_kernel.pyx``
  by ``oracle/build.py``) and against the reference's known-answer tests
  (``tests/golden/kat_*.json`` transcribed from ``/root/reference/tests``).
* Float half: the arithmetic lives in PaddlePaddle (``paddle.geometric.*``, PyPI
  ``paddlepaddle``; version NOT pinned by the reference: ``requirements.txt:1-2``
  lists numpy+cython only, ``README.md:146`` says ``>=2.2.0``, the code needs
  ``>=2.4`` for ``paddle.geometric``).  Paddle is absent from this image, so the float
  functions below restate Paddle's published op contract and are pinned by the
  reference's own golden vectors for this path.
  What no reference test pins (mean/max/min of send_u_recv, send_uv, out_size,
  any backward) is "parity unpinned": restatement + documented contract only.

Every function cites the reference file:line it follows (paths relative to
``/root/reference``).
"""
import numpy as np

# --------------------------------------------------------------------------
# Integer side: CSR build, degrees, segment ids
# --------------------------------------------------------------------------


def build_index(u, v, num_nodes):
    """pgl/graph_kernel.pyx:59-88 -- histogram, exclusive scan, STABLE counting sort.

    Returns (degree, sorted_v, sorted_u, sorted_eid, indptr), all int64, edges
    inside one bucket keep ascending edge id.
    """
    u = np.ascontiguousarray(u, dtype=np.int64)
    v = np.ascontiguousarray(v, dtype=np.int64)
    n = int(num_nodes)
    degree = np.bincount(u, minlength=n).astype(np.int64)[:n] if len(u) else np.zeros(n, np.int64)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(degree, out=indptr[1:])
    order = np.argsort(u, kind="stable").astype(np.int64)
    return degree, v[order], u[order], order, indptr


def build_index_loops(u, v, num_nodes):
    """Literal three-loop transcription of pgl/graph_kernel.pyx:76-87 (small inputs only)."""
    h = len(u)
    n = int(num_nodes)
    degree = np.zeros(n, np.int64)
    count = np.zeros(n, np.int64)
    tv = np.zeros(h, np.int64)
    tu = np.zeros(h, np.int64)
    te = np.zeros(h, np.int64)
    indptr = np.zeros(n + 1, np.int64)
    for i in range(h):
        degree[u[i]] += 1
    for i in range(n):
        indptr[i + 1] = indptr[i] + degree[i]
    for i in range(h):
        p = indptr[u[i]] + count[u[i]]
        tv[p] = v[i]
        te[p] = i
        tu[p] = u[i]
        count[u[i]] += 1
    return degree, tv, tu, te, indptr


def adj_dst_index(edges, num_nodes):
    """pgl/graph.py:1319-1328 -- CSR keyed by dst: u := edges[:,1], v := edges[:,0]."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    return build_index(edges[:, 1], edges[:, 0], num_nodes)


def adj_src_index(edges, num_nodes):
    """pgl/graph.py:1308-1317 -- CSR keyed by src: u := edges[:,0], v := edges[:,1]."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    return build_index(edges[:, 0], edges[:, 1], num_nodes)


def sorted_edges(edges, num_nodes, sort_by="src"):
    """pgl/graph.py:392-413 -> (src, dst, eid) sorted by the chosen endpoint."""
    if sort_by not in ("src", "dst"):
        raise ValueError("sort_by should be in 'src' or 'dst'.")
    if sort_by == "src":
        _, v, u, eid, _ = adj_src_index(edges, num_nodes)
        return u, v, eid
    _, v, u, eid, _ = adj_dst_index(edges, num_nodes)
    return v, u, eid


def unique_segment(sorted_key):
    """pgl/utils/helper.py:156-160 -- paddle.unique(x, return_inverse=True, dtype=int64)."""
    uniq, inv = np.unique(np.asarray(sorted_key, np.int64), return_inverse=True)
    return uniq.astype(np.int64), inv.astype(np.int64).reshape(-1)


def maybe_num_nodes(edges):
    """pgl/utils/helper.py:133-153."""
    edges = np.asarray(edges)
    if len(edges) == 0:
        return 0
    return int(np.max(edges)) + 1


def shard_edges_by_dst(edges, world_size, rank):
    """pgl/graph.py:1475-1504 -- DistGPUGraph keeps edges with dst % world == rank."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    mask = (edges[:, 1] % world_size) == rank
    eid = np.arange(edges.shape[0])[mask]
    return edges[eid], eid


def metis_weight_scale(x):
    """pgl/partition.py:25-34."""
    x = np.asarray(x)
    x_min = np.min(x)
    x_max = np.max(x)
    xs = (x - x_min) / (x_max - x_min + 1e-5)
    return (xs * 1000).astype("int64") + 1


# --------------------------------------------------------------------------
# Float side: Paddle op contract restated (paddle.geometric.*, gather, scatter)
# --------------------------------------------------------------------------

_MSG = {
    "add": lambda a, b: a + b,
    "sub": lambda a, b: a - b,
    "mul": lambda a, b: a * b,
    "div": lambda a, b: a / b,
}


def _feat_broadcast(x_rows, y_rows):
    """NumPy broadcasting over the non-leading dims (paddle.geometric.send_ue_recv /
    send_uv docs: "Broadcasting follows NumPy semantics"); a 1-D edge operand acts as
    [E, 1] (pinned by tests/test_dist_graph.py:115-137: efeat shape [E] against x [N,4])."""
    if x_rows.ndim == 1:
        x_rows = x_rows[:, None]
    if y_rows.ndim == 1:
        y_rows = y_rows[:, None]
    nd = max(x_rows.ndim, y_rows.ndim)
    xs = (x_rows.shape[0],) + (1,) * (nd - x_rows.ndim) + x_rows.shape[1:]
    ys = (y_rows.shape[0],) + (1,) * (nd - y_rows.ndim) + y_rows.shape[1:]
    return x_rows.reshape(xs), y_rows.reshape(ys)


def _reduce_rows(msg, dst, n_out, reduce_op):
    """Sequential COO loop in edge order == Paddle phi CPU GraphSendRecvCpuLoop:
    out zero-initialised; sum: out[d] += m; mean: sum then / count where count>0;
    max/min: first message copies, later ones max/min; rows with no message stay 0."""
    out = np.zeros((n_out,) + msg.shape[1:], dtype=msg.dtype)
    if reduce_op in ("sum", "mean"):
        # np.add.at applies updates in index order -> same fp32 summation order as the loop
        np.add.at(out, dst, msg)
        if reduce_op == "mean":
            cnt = np.bincount(dst, minlength=n_out)[:n_out]
            nz = cnt > 0
            out[nz] = out[nz] / cnt[nz].astype(msg.dtype).reshape((-1,) + (1,) * (msg.ndim - 1))
    elif reduce_op in ("max", "min"):
        fill = -np.inf if reduce_op == "max" else np.inf
        tmp = np.full_like(out, fill)
        (np.maximum if reduce_op == "max" else np.minimum).at(tmp, dst, msg)
        cnt = np.bincount(dst, minlength=n_out)[:n_out]
        nz = cnt > 0
        out[nz] = tmp[nz]
    else:
        raise AssertionError("Only support 'sum', 'mean', 'max', 'min' built-in reduce functions.")
    return out


def _out_rows(x, out_size):
    if out_size is None or int(out_size) <= 0:
        return x.shape[0]
    return int(out_size)


def send_u_recv(x, src, dst, reduce_op="sum", out_size=None):
    """paddle.geometric.send_u_recv as called at pgl/graph.py:860,886.
    out[d] = reduce_{e: dst[e]=d} x[src[e]]; rows = out_size if >0 else x.shape[0]."""
    x = np.asarray(x)
    src = np.asarray(src, np.int64)
    dst = np.asarray(dst, np.int64)
    return _reduce_rows(x[src], dst, _out_rows(x, out_size), reduce_op)


def send_ue_recv(x, y, src, dst, message_op="add", reduce_op="sum", out_size=None):
    """paddle.geometric.send_ue_recv as called at pgl/graph.py:930."""
    assert message_op in _MSG
    x = np.asarray(x)
    y = np.asarray(y)
    src = np.asarray(src, np.int64)
    dst = np.asarray(dst, np.int64)
    xr, yr = _feat_broadcast(x[src], y)
    msg = _MSG[message_op](xr, yr).astype(np.result_type(x.dtype, y.dtype))
    return _reduce_rows(msg, dst, _out_rows(x, out_size), reduce_op)


def send_uv(x, y, src, dst, message_op="add"):
    """paddle.geometric.send_uv as called at pgl/graph.py:965: out[e] = x[src[e]] op y[dst[e]]."""
    assert message_op in _MSG
    x = np.asarray(x)
    y = np.asarray(y)
    xr, yr = _feat_broadcast(x[np.asarray(src, np.int64)], y[np.asarray(dst, np.int64)])
    return _MSG[message_op](xr, yr)


def segment_pool(data, segment_ids, pool_type):
    """pgl/math.py:30-46 -> paddle.geometric.segment_{sum,mean,max,min}.
    ids non-decreasing; out rows = ids[-1]+1; absent ids give zero rows."""
    data = np.asarray(data)
    ids = np.asarray(segment_ids, np.int64)
    pt = pool_type.lower()
    if pt not in ("sum", "mean", "max", "min"):
        raise ValueError("We only support sum, mean, max, min pool types in segment_pool function.")
    k = int(ids[-1]) + 1 if len(ids) else 0
    return _reduce_rows(data, ids, k, pt)


def segment_sum(data, segment_ids):
    """pgl/math.py:49-79."""
    return segment_pool(data, segment_ids, "sum")


def segment_mean(data, segment_ids):
    """pgl/math.py:82-113."""
    return segment_pool(data, segment_ids, "mean")


def segment_min(data, segment_ids):
    """pgl/math.py:116-145."""
    return segment_pool(data, segment_ids, "min")


def segment_max(data, segment_ids):
    """pgl/math.py:148-178."""
    return segment_pool(data, segment_ids, "max")


def segment_softmax(data, segment_ids):
    """pgl/math.py:216-224: max -> gather -> sub -> exp -> sum -> gather -> div."""
    data = np.asarray(data)
    ids = np.asarray(segment_ids, np.int64)
    data_max = segment_max(data, ids)[ids]
    e = np.exp(data - data_max)
    s = segment_sum(e, ids)[ids]
    return e / s


# --------------------------------------------------------------------------
# Graph.send / Graph.recv (UDF path), edge_softmax, degree_norm
# --------------------------------------------------------------------------


class RowReader(dict):
    """pgl/utils/op.py:75-87 -- lazy, memoised per-key row gather."""

    def __init__(self, nfeat, index):
        super().__init__()
        self.nfeat = nfeat
        self.loaded = {}
        self.index = index

    def __getitem__(self, key):
        if key not in self.loaded:
            self.loaded[key] = _read_rows(self.nfeat[key], self.index)
        return self.loaded[key]


def _read_rows(data, index):
    """pgl/utils/op.py:24-45."""
    if data is None:
        return None
    if isinstance(data, dict):
        return {k: _read_rows(v, index) for k, v in data.items()}
    return np.asarray(data)[index]


class Message(object):
    """pgl/message.py:19-173."""

    def __init__(self, msg, segment_ids):
        self._segment_ids = segment_ids
        self._msg = msg

    def reduce(self, msg, pool_type="sum"):
        return segment_pool(msg, self._segment_ids, pool_type)

    def reduce_sum(self, msg):
        return segment_sum(msg, self._segment_ids)

    def reduce_mean(self, msg):
        return segment_mean(msg, self._segment_ids)

    def reduce_max(self, msg):
        return segment_max(msg, self._segment_ids)

    def reduce_min(self, msg):
        return segment_min(msg, self._segment_ids)

    def edge_expand(self, msg):
        return np.asarray(msg)[self._segment_ids]

    def reduce_softmax(self, msg):
        return segment_softmax(msg, self._segment_ids)

    def __getitem__(self, key):
        return self._msg[key]


def send(edges, message_func, src_feat=None, dst_feat=None, edge_feat=None, node_feat=None):
    """pgl/graph.py:694-776 (tensor-mode branch)."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    if (src_feat is not None or dst_feat is not None) and node_feat is not None:
        raise ValueError("Can not use src/dst feat and node feat at the same time")
    sft, dft, eft = {}, {}, {}
    if node_feat is not None:
        assert isinstance(node_feat, dict), "The input node_feat must be a dict"
        sft.update(node_feat)
        dft.update(node_feat)
    else:
        if src_feat is not None:
            assert isinstance(src_feat, dict), "The input src_feat must be a dict"
            sft.update(src_feat)
        if dst_feat is not None:
            assert isinstance(dst_feat, dict), "The input dst_feat must be a dict"
            dft.update(dst_feat)
    if edge_feat is not None:
        assert isinstance(edge_feat, dict), "The input edge_feat must be a dict"
        eft.update(edge_feat)
    msg = message_func(RowReader(sft, edges[:, 0]), RowReader(dft, edges[:, 1]), eft)
    if not isinstance(msg, dict):
        raise TypeError("The outputs of the %s function is expected to be a dict, but got %s"
                        % (message_func.__name__, type(msg)))
    return msg


def recv(edges, num_nodes, reduce_func, msg, recv_mode="dst"):
    """pgl/graph.py:778-832: sorted_edges -> RowReader(msg, eid) -> unique -> reduce ->
    zeros([N, D_out]) + scatter(uniq_ind)."""
    if not isinstance(msg, dict):
        raise TypeError("The input of msg should be a dict, but receives a %s" % (type(msg)))
    if not callable(reduce_func):
        raise TypeError("reduce_func should be callable")
    src, dst, eid = sorted_edges(edges, num_nodes, sort_by=recv_mode)
    m = RowReader(msg, eid)
    uniq, seg = unique_segment(dst if recv_mode == "dst" else src)
    output = np.asarray(reduce_func(Message(m, seg)))
    final = np.zeros((int(num_nodes), output.shape[-1]), dtype=output.dtype)
    final[uniq] = output
    return final


def edge_softmax(edges, num_nodes, logits, norm_by="dst"):
    """pgl/nn/functional/graph_op.py:101-123."""
    src, dst, eid = sorted_edges(edges, num_nodes, sort_by=norm_by)
    _, seg = unique_segment(dst if norm_by == "dst" else src)
    lg = np.asarray(logits)[eid]
    score = segment_softmax(lg, seg)
    out = np.zeros_like(score)
    out[eid] = score
    return out


def degree_norm(degree, dtype=np.float32):
    """pgl/nn/functional/graph_op.py:46-55: cast -> clip(min=1) -> pow(-0.5) -> [N,1]."""
    norm = np.asarray(degree).astype(dtype)
    norm = np.clip(norm, 1.0, None)
    norm = np.power(norm, dtype(-0.5)).astype(dtype)
    return norm.reshape(-1, 1)


# --------------------------------------------------------------------------
# Conv layers (forward only), weights supplied by the caller.
# Linear follows paddle.nn.Linear: y = x @ W + b with W [in, out].
# --------------------------------------------------------------------------


def _act(x, act):
    if act is None:
        return x
    if act == "relu":
        return np.maximum(x, 0)
    raise ValueError(act)


def gcn_conv(edges, num_nodes, feature, weight, bias, activation=None, norm=True):
    """pgl/nn/conv.py:218-254.  weight [in, out]."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    in_size, out_size = weight.shape
    nrm = None
    if norm:
        deg = adj_dst_index(edges, num_nodes)[0]
        nrm = degree_norm(deg, feature.dtype.type)
    if in_size > out_size:
        feature = feature @ weight
    if nrm is not None:
        feature = feature * nrm
    output = send_u_recv(feature, edges[:, 0], edges[:, 1], "sum")
    if in_size <= out_size:
        output = output @ weight
    if nrm is not None:
        output = output * nrm
    output = output + bias
    return _act(output, activation)


def graphsage_conv(edges, feature_src, feature_dst, w_self, b_self, w_neigh, b_neigh,
                   aggr_func="sum", act=None, normalize=True):
    """pgl/nn/conv.py:81-115."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    neigh = send_u_recv(feature_src, edges[:, 0], edges[:, 1], aggr_func,
                        out_size=feature_dst.shape[0])
    neigh = neigh @ w_neigh + b_neigh
    self_f = feature_dst @ w_self + b_self
    out = _act(self_f + neigh, act)
    if normalize:
        # paddle.nn.functional.normalize(p=2, axis=1, epsilon=1e-12): x / max(||x||, eps)
        nrm = np.sqrt(np.sum(out * out, axis=1, keepdims=True))
        out = out / np.maximum(nrm, 1e-12)
    return out


def gat_conv(edges, num_nodes, feature, w, b, weight_src, weight_dst, num_heads, hidden,
             concat=True, activation=None, negative_slope=0.2):
    """pgl/nn/conv.py:308-346 with feat_drop = attn_drop = 0 (eval mode)."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    f = (feature @ w + b).reshape(-1, num_heads, hidden)
    attn_src = np.sum(f * weight_src, axis=-1)
    attn_dst = np.sum(f * weight_dst, axis=-1)
    alpha = send_uv(attn_src, attn_dst, edges[:, 0], edges[:, 1], "add")
    alpha = np.where(alpha >= 0, alpha, alpha * np.float32(negative_slope)).astype(f.dtype)
    alpha = edge_softmax(edges, num_nodes, alpha, "dst")
    alpha = alpha.reshape(-1, num_heads, 1)
    out = send_ue_recv(f, alpha, edges[:, 0], edges[:, 1], "mul", "sum")
    if concat:
        out = out.reshape(-1, num_heads * hidden)
    else:
        out = out.mean(axis=1)
    return _act(out, activation)


def reindex_graph(x, neighbors, count):
    """paddle.geometric.reindex_graph as the reference's GPU sampler uses it
    (pgl/sampling/sage.py:146-147): out_nodes = x followed by the ids of ``neighbors`` not in x, in
    first-appearance order; reindex_src = neighbors mapped to positions in out_nodes; reindex_dst =
    position of the owning x entry, repeated count[i] times.  Pinned by the example in Paddle's API
    reference (x=[0,1,2], neighbors=[8,9,0,4,7,6,7], count=[2,3,2])."""
    x = np.asarray(x, np.int64).reshape(-1)
    neighbors = np.asarray(neighbors, np.int64).reshape(-1)
    count = np.asarray(count, np.int64).reshape(-1)
    pos = {int(v): i for i, v in enumerate(x)}
    out_nodes = [int(v) for v in x]
    src = np.empty(len(neighbors), np.int64)
    for p, v in enumerate(neighbors):
        v = int(v)
        if v not in pos:
            pos[v] = len(out_nodes)
            out_nodes.append(v)
        src[p] = pos[v]
    dst = np.repeat(np.arange(len(x), dtype=np.int64), count)
    return src, dst, np.asarray(out_nodes, np.int64)


def check_sampled_neighbors(indptr, row, nodes, sample_size, neighbors, count):
    """Properties every valid output of paddle.geometric.sample_neighbors has (the draw itself is
    random): count = min(deg, k) (deg when k < 0), whole list in order when deg <= k, otherwise a
    sub-multiset of the neighbour list (no slot used twice).  Returns None or a message."""
    indptr, row = np.asarray(indptr), np.asarray(row)
    o = 0
    for i, v in enumerate(np.asarray(nodes).reshape(-1)):
        lst = row[indptr[v]:indptr[v + 1]]
        deg = len(lst)
        want = deg if (sample_size < 0 or deg <= sample_size) else sample_size
        if int(count[i]) != want:
            return "count[%d] = %d, expected %d" % (i, int(count[i]), want)
        got = np.asarray(neighbors[o:o + want])
        if want == deg:
            if not (got == lst).all():
                return "node %d: full list expected in CSR order" % int(v)
        else:
            a, ca = np.unique(got, return_counts=True)
            b, cb = np.unique(lst, return_counts=True)
            m = dict(zip(b.tolist(), cb.tolist()))
            if any(m.get(int(t), 0) < int(c) for t, c in zip(a, ca)):
                return "node %d: sample is not a sub-multiset of its neighbour list" % int(v)
        o += want
    if o != len(neighbors):
        return "neighbors has %d entries, counts add up to %d" % (len(neighbors), o)
    return None


def load_cora(path):
    """BASELINE config 2 from the committed fixture tests/golden/cora.npz (made by
    tests/golden/make_cora.py from the reference's pgl/data/cora with the logic of
    pgl/dataset.py:195-246).  Returns a dict: num_nodes, edges [E,2] int64, x [N,1433] float32
    (row-normalised by the dataset, pgl/dataset.py:213, and again by the training script,
    examples/citation_benchmark/train.py:29-30,44), y, num_classes and the fixed
    train / val / test index split (dataset.py:240-243)."""
    z = np.load(path)
    n, words = int(z["num_nodes"]), int(z["num_words"])
    x = np.zeros((n, words), np.float32)
    x[z["feat_row"].astype(np.int64), z["feat_col"].astype(np.int64)] = 1.0
    x = (x / (x.sum(axis=1, keepdims=True) + np.float32(1e-15))).astype(np.float32)
    x = (x / np.maximum(x.sum(axis=-1, keepdims=True), 1)).astype(np.float32)
    perm = np.arange(n)
    y = z["labels"].astype(np.int64)
    return {"num_nodes": n, "edges": z["edges"].astype(np.int64), "x": x, "y": y,
            "num_classes": int(y.max()) + 1, "train_index": perm[:140], "val_index": perm[200:500],
            "test_index": perm[500:1500]}


# --------------------------------------------------------------------------
# Workload generators shared by tests and bench
# --------------------------------------------------------------------------


def chung_lu_edges(num_nodes, num_edges, exponent=0.8, seed=20240922):
    """Config-5 style power-law graph: w_i ~ (i+1)^-exponent, src,dst ~ Cat(w) iid,
    ids relabelled by a fixed random permutation, duplicates / self loops kept."""
    rng = np.random.default_rng(seed)
    w = np.power(np.arange(1, num_nodes + 1, dtype=np.float64), -exponent)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    src = np.searchsorted(cdf, rng.random(num_edges)).astype(np.int64)
    dst = np.searchsorted(cdf, rng.random(num_edges)).astype(np.int64)
    np.minimum(src, num_nodes - 1, out=src)
    np.minimum(dst, num_nodes - 1, out=dst)
    perm = rng.permutation(num_nodes).astype(np.int64)
    return np.stack([perm[src], perm[dst]], axis=1)
