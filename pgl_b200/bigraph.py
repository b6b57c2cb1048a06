"""BiGraph: rectangular (src nodes x dst nodes) message passing -- mirror of the send/recv part of
reference pgl/bigraph.py (constructor :123-217, indexes :528-547, degrees :639-681,
send_recv :1051-1085, send :1087-1157, recv :1159-1226).

The kernels already take the number of source rows and of output rows separately, so a bipartite
block (a sampled mini-batch layer, a user-item graph) is the same dst-CSR with ``dst_num_nodes``
rows whose columns index ``src_num_nodes`` feature rows.  ``send_recv`` runs the fused aggregation
(the reference gathers [E, D] and scatter-adds it, sum only); mean/max/min come for free.
Sampling, batching (disjoint) and dump/load of BiGraph stay out of scope.
"""
import numpy as np
import torch

from . import ops
from .message import Message
from .utils import op
from .utils.edge_index import EdgeIndex
from .utils.helper import check_is_tensor, maybe_num_nodes, to_tensor

__all__ = ["BiGraph"]


def _as_int(x):
    if isinstance(x, torch.Tensor):
        return int(x.item())
    return int(np.asarray(x).reshape(-1)[0]) if not isinstance(x, int) else x


class BiGraph(object):
    """reference pgl/bigraph.py:36-217.  ``edges[:, 0]`` indexes the source node set
    (``src_num_nodes`` rows), ``edges[:, 1]`` the destination node set (``dst_num_nodes``)."""

    def __init__(self, edges, src_num_nodes=None, dst_num_nodes=None, src_node_feat=None,
                 dst_node_feat=None, edge_feat=None, **kwargs):
        self._src_node_feat = src_node_feat if src_node_feat is not None else {}
        self._dst_node_feat = dst_node_feat if dst_node_feat is not None else {}
        self._edge_feat = edge_feat if edge_feat is not None else {}
        if not check_is_tensor(edges):
            edges = np.array(edges, dtype="int64").reshape(-1, 2)
        self._edges = edges
        self._src_num_nodes = maybe_num_nodes(self._edges[:, 0]) if src_num_nodes is None \
            else src_num_nodes
        self._dst_num_nodes = maybe_num_nodes(self._edges[:, 1]) if dst_num_nodes is None \
            else dst_num_nodes
        self._adj_src_index = kwargs.get("adj_src_index", None)
        self._adj_dst_index = kwargs.get("adj_dst_index", None)
        feats = list(self._src_node_feat.values()) + list(self._dst_node_feat.values()) + \
            list(self._edge_feat.values())
        self._is_tensor = bool(
            check_is_tensor(self._src_num_nodes, self._dst_num_nodes, self._edges, *feats)
            or (self._adj_src_index is not None and self._adj_src_index.is_tensor())
            or (self._adj_dst_index is not None and self._adj_dst_index.is_tensor()))
        self._n_src = _as_int(self._src_num_nodes)
        self._n_dst = _as_int(self._dst_num_nodes)
        if self._is_tensor:
            self._is_tensor = False
            self.tensor(inplace=True)
        self._src_nodes = None
        self._dst_nodes = None

    def __repr__(self):
        return "BiGraph(src_num_nodes=%d, dst_num_nodes=%d, edges_shape=%s)" % (
            self._n_src, self._n_dst, list(self._edges.shape))

    # ------------------------------------------------------------------ mode switch
    def is_tensor(self):
        return self._is_tensor

    def tensor(self, inplace=True):
        """reference bigraph.py:367-402."""
        if self._is_tensor:
            return self
        if not inplace:
            return BiGraph(to_tensor(self._edges), to_tensor(np.int64(self._n_src)),
                           to_tensor(np.int64(self._n_dst)),
                           {k: to_tensor(v) for k, v in self._src_node_feat.items()},
                           {k: to_tensor(v) for k, v in self._dst_node_feat.items()},
                           {k: to_tensor(v) for k, v in self._edge_feat.items()})
        self._src_num_nodes = to_tensor(np.int64(self._n_src))
        self._dst_num_nodes = to_tensor(np.int64(self._n_dst))
        self._edges = to_tensor(self._edges)
        if self._edges.dtype != torch.int64:
            self._edges = self._edges.to(torch.int64)
        self._edges = self._edges.reshape(-1, 2).contiguous()
        for feat in (self._src_node_feat, self._dst_node_feat, self._edge_feat):
            for key in feat:
                feat[key] = to_tensor(feat[key])
        for index in (self._adj_src_index, self._adj_dst_index):
            if index is not None and not index.is_tensor():
                index.tensor(inplace=True)
        self._src_nodes = self._dst_nodes = None
        self._is_tensor = True
        return self

    def numpy(self, inplace=True):
        """reference bigraph.py:427-462."""
        if not self._is_tensor:
            return self

        def host(x):
            return x.cpu().numpy() if isinstance(x, torch.Tensor) else x

        if not inplace:
            return BiGraph(host(self._edges), self._n_src, self._n_dst,
                           {k: host(v) for k, v in self._src_node_feat.items()},
                           {k: host(v) for k, v in self._dst_node_feat.items()},
                           {k: host(v) for k, v in self._edge_feat.items()})
        self._src_num_nodes, self._dst_num_nodes = self._n_src, self._n_dst
        self._edges = host(self._edges)
        for feat in (self._src_node_feat, self._dst_node_feat, self._edge_feat):
            for key in feat:
                feat[key] = host(feat[key])
        for index in (self._adj_src_index, self._adj_dst_index):
            if index is not None and index.is_tensor():
                index.numpy(inplace=True)
        for attr in ("_dst_uniq_ind", "_dst_segment_ids", "_src_uniq_ind", "_src_segment_ids"):
            self.__dict__.pop(attr, None)
        self._src_nodes = self._dst_nodes = None
        self._is_tensor = False
        return self

    # ------------------------------------------------------------------ structure
    @property
    def adj_src_index(self):
        """CSR keyed by source (u = src, v = dst); reference bigraph.py:528-536."""
        if self._adj_src_index is None:
            self._adj_src_index = EdgeIndex.from_edges(u=self._edges[:, 0], v=self._edges[:, 1],
                                                       num_nodes=self._src_num_nodes, v_bound=self._dst_num_nodes)
        return self._adj_src_index

    @property
    def adj_dst_index(self):
        """CSR keyed by destination (u = dst, v = src); reference bigraph.py:539-547."""
        if self._adj_dst_index is None:
            self._adj_dst_index = EdgeIndex.from_edges(u=self._edges[:, 1], v=self._edges[:, 0],
                                                       num_nodes=self._dst_num_nodes, v_bound=self._src_num_nodes)
        return self._adj_dst_index

    @property
    def edge_feat(self):
        return self._edge_feat

    @property
    def src_node_feat(self):
        return self._src_node_feat

    @property
    def dst_node_feat(self):
        return self._dst_node_feat

    @property
    def num_edges(self):
        return self._edges.shape[0]

    @property
    def src_num_nodes(self):
        return self._src_num_nodes

    @property
    def dst_num_nodes(self):
        return self._dst_num_nodes

    @property
    def edges(self):
        return self._edges

    def sorted_edges(self, sort_by="src"):
        """reference bigraph.py:594-615."""
        if sort_by not in ["src", "dst"]:
            raise ValueError("sort_by should be in 'src' or 'dst'.")
        if sort_by == "src":
            src, dst, eid = self.adj_src_index.triples()
        else:
            dst, src, eid = self.adj_dst_index.triples()
        return src, dst, eid

    def _arange(self, n):
        if self._is_tensor:
            return torch.arange(n, dtype=torch.int64, device=self._edges.device)
        return np.arange(n, dtype=np.int64)

    @property
    def src_nodes(self):
        if self._src_nodes is None:
            self._src_nodes = self._arange(self._n_src)
        return self._src_nodes

    @property
    def dst_nodes(self):
        if self._dst_nodes is None:
            self._dst_nodes = self._arange(self._n_dst)
        return self._dst_nodes

    def indegree(self, nodes=None):
        """In-degree of the DESTINATION nodes; reference bigraph.py:639-659."""
        degree = self.adj_dst_index.degree
        if nodes is None:
            return degree
        return ops.gather_rows(degree, to_tensor(nodes)) if self._is_tensor else degree[nodes]

    def outdegree(self, nodes=None):
        """Out-degree of the SOURCE nodes; reference bigraph.py:661-681."""
        degree = self.adj_src_index.degree
        if nodes is None:
            return degree
        return ops.gather_rows(degree, to_tensor(nodes)) if self._is_tensor else degree[nodes]

    # ------------------------------------------------------------------ message passing
    def send_recv(self, feature, reduce_func="sum"):
        """out[dst_num_nodes, D] = reduce over in-edges of feature[src]; reference
        bigraph.py:1051-1085 (sum only there: gather + scatter-add)."""
        assert reduce_func in ["sum", "mean", "max", "min"], \
            "Only support 'sum', 'mean', 'max', 'min' built-in receive function."
        assert isinstance(feature, torch.Tensor), \
            "The input of send_recv method should be tensor."
        if not self._is_tensor:
            raise ValueError("You must call BiGraph.tensor() first")
        ops.require_cuda(feature)
        if int(feature.shape[0]) != self._n_src:
            raise ValueError("feature has %d rows, the graph has %d source nodes"
                             % (int(feature.shape[0]), self._n_src))
        bwd = self.adj_src_index.csr if (feature.requires_grad and torch.is_grad_enabled()) \
            else None
        return ops.aggregate_copy(feature, self.adj_dst_index.csr(), self._n_dst, reduce_func,
                                  bwd=bwd)

    def send(self, message_func, src_feat=None, dst_feat=None, edge_feat=None):
        """reference bigraph.py:1087-1157."""
        if not self._is_tensor:
            raise ValueError("You must call BiGraph.tensor() first")
        src_feat_temp, dst_feat_temp, edge_feat_temp = {}, {}, {}
        if src_feat is not None:
            assert isinstance(src_feat, dict), "The input src_feat must be a dict"
            src_feat_temp.update(src_feat)
        if dst_feat is not None:
            assert isinstance(dst_feat, dict), "The input dst_feat must be a dict"
            dst_feat_temp.update(dst_feat)
        if edge_feat is not None:
            assert isinstance(edge_feat, dict), "The input edge_feat must be a dict"
            edge_feat_temp.update(edge_feat)
        msg = message_func(op.RowReader(src_feat_temp, self._edges[:, 0]),
                           op.RowReader(dst_feat_temp, self._edges[:, 1]), edge_feat_temp)
        if not isinstance(msg, dict):
            raise TypeError(
                "The outputs of the %s function is expected to be a dict, but got %s"
                % (message_func.__name__, type(msg)))
        return msg

    def _segments(self, segment_by):
        attr_u, attr_s = "_%s_uniq_ind" % segment_by, "_%s_segment_ids" % segment_by
        if not hasattr(self, attr_u):
            index = self.adj_dst_index if segment_by == "dst" else self.adj_src_index
            uniq, seg = ops.segment_ids_from_indptr(index._indptr, int(self.num_edges))
            indptr = index._indptr
            compact = torch.cat([indptr.index_select(0, uniq), indptr[-1:]]) if uniq.numel() \
                else torch.zeros(1, dtype=torch.int64, device=indptr.device)
            seg._pglb_indptr = compact
            seg._pglb_max_degree = index.max_degree
            setattr(self, attr_u, uniq)
            setattr(self, attr_s, seg)
        return getattr(self, attr_u), getattr(self, attr_s)

    def recv(self, reduce_func, msg, recv_mode="dst"):
        """reference bigraph.py:1159-1226: rows = dst_num_nodes (recv_mode 'dst') or src_num_nodes
        ('src'); nodes without a message get zeros."""
        if not self._is_tensor:
            raise ValueError("You must call BiGraph.tensor()")
        if not isinstance(msg, dict):
            raise TypeError("The input of msg should be a dict, but receives a %s" % (type(msg)))
        if not callable(reduce_func):
            raise TypeError("reduce_func should be callable")
        _, _, eid = self.sorted_edges(sort_by=recv_mode)
        uniq_ind, segment_ids = self._segments(recv_mode)
        output = reduce_func(Message(op.RowReader(msg, eid), segment_ids))
        rows = self._n_dst if recv_mode == "dst" else self._n_src
        init_output = torch.zeros((rows, output.shape[-1]), dtype=output.dtype,
                                  device=output.device)
        if int(uniq_ind.shape[0]) == 0:
            return init_output
        return ops.scatter_rows(init_output, uniq_ind, output)
