"""NeighborSampler: layer-wise uniform neighbour sampling on the device for mini-batch GraphSAGE
(reference pgl/sampling/sage.py:130-155).  ``sample_neighbors`` / ``reindex_graph`` stand in for
``paddle.geometric.sample_neighbors`` / ``paddle.geometric.reindex_graph`` on the sm_100a kernels of
csrc/sampling.cu (validated on hardware in round 2: tests/test_gpu_sampling.py)."""
import ctypes

import torch

from .. import ops
from .._lib import check, lib
from ..graph import Graph

__all__ = ["NeighborSampler", "sample_neighbors", "reindex_graph"]


def sample_neighbors(row, colptr, input_nodes, sample_size=-1, eids=None, return_eids=False, seed=0):
    """paddle.geometric.sample_neighbors: for every node of ``input_nodes`` up to ``sample_size``
    in-neighbours drawn uniformly without replacement from ``row[colptr[v]:colptr[v+1]]`` (all of them
    when sample_size is -1 or the degree is not larger).  Returns (neighbors, count[, eids])."""
    ops.require_cuda(row, colptr, input_nodes)
    nodes = input_nodes.reshape(-1).to(torch.int64).contiguous()
    n = int(nodes.shape[0])
    dev = nodes.device
    count = torch.empty(n, dtype=torch.int64, device=dev)
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    with torch.cuda.device(dev):
        check(lib.pglb_sample_count(p(colptr), p(nodes), n, int(sample_size), p(count), p(offsets), stream))
    total = int(offsets[n].item())
    neighbors = torch.empty(total, dtype=torch.int64, device=dev)
    out_eids = torch.empty(total, dtype=torch.int64, device=dev) if return_eids else None
    if total > 0:  # sample_size == 0 (or no neighbours at all): nothing to fill, and an empty tensor has no address
        with torch.cuda.device(dev):
            check(lib.pglb_sample_fill(p(colptr), p(row), p(eids), p(nodes), n, int(sample_size),
                                       int(seed) & 0xFFFFFFFFFFFFFFFF, p(offsets), p(neighbors), p(out_eids),
                                       stream))
    neighbors._pglb_offsets = offsets  # rides along for reindex_graph (saves a scan)
    if return_eids:
        return neighbors, count, out_eids
    return neighbors, count


_tables = {}


def _table(dev, num_nodes):
    key = (str(dev), int(num_nodes))
    t = _tables.get(key)
    if t is None:
        t = torch.empty(int(num_nodes), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            check(lib.pglb_reindex_table_init(ctypes.c_void_p(t.data_ptr()), int(num_nodes),
                                              ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        _tables[key] = t
    return t


def reindex_graph(x, neighbors, count, num_nodes=None):
    """paddle.geometric.reindex_graph: compact ids in first-appearance order (``x`` first).
    Returns (reindex_src, reindex_dst, out_nodes).  ``num_nodes`` bounds the ids (size of the dense
    lookup table, cached per device); default max id + 1."""
    ops.require_cuda(x, neighbors, count)
    x = x.reshape(-1).to(torch.int64).contiguous()
    offsets = getattr(neighbors, "_pglb_offsets", None)  # set by sample_neighbors (saves a scan)
    neighbors = neighbors.reshape(-1)
    nb = neighbors.to(torch.int64).contiguous()
    n, m = int(x.shape[0]), int(nb.shape[0])
    dev = x.device
    if offsets is None or int(offsets.shape[0]) != n + 1:
        offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(count.reshape(-1).to(torch.int64), 0, out=offsets[1:])
    if num_nodes is None:
        num_nodes = int(max(int(x.max().item()) if n else -1, int(nb.max().item()) if m else -1)) + 1
    table = _table(dev, num_nodes)
    src = torch.empty(m, dtype=torch.int64, device=dev)
    dst = torch.empty(m, dtype=torch.int64, device=dev)
    out_nodes = torch.empty(n + m, dtype=torch.int64, device=dev)
    num_out = torch.zeros(1, dtype=torch.int64, device=dev)
    need = ctypes.c_size_t(0)
    check(lib.pglb_reindex_graph_ws(m, ctypes.byref(need)))
    ws = ops.workspace(dev, need.value)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    with torch.cuda.device(dev):
        check(lib.pglb_reindex_graph(p(x), n, p(nb), p(offsets), m, p(table), p(src), p(dst), p(out_nodes),
                                     p(num_out), p(ws), ws.numel(),
                                     ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return src, dst, out_nodes[: int(num_out.item())]


class NeighborSampler(object):
    """reference pgl/sampling/sage.py:130-155: ``samples`` = fan-out per layer (outermost first);
    ``sample_neighbors(nodes)`` returns ``(graph_list[::-1], nodes)`` where every entry is
    ``(subgraph, number of target nodes)`` and the returned nodes are the input nodes of the first
    layer (gather the features with them)."""

    def __init__(self, graph, samples, uva=False, seed=0):
        if uva:
            raise ValueError("uva mode is not supported by pgl_b200")
        if not graph.is_tensor():
            graph = graph.tensor(inplace=False)
        self.graph = graph
        self.samples = samples
        self.row = graph.adj_dst_index._sorted_v
        self.colptr = graph.adj_dst_index._indptr
        self.num_nodes = graph._n
        self.seed = int(seed)
        self._calls = 0

    def sample_neighbors(self, nodes):
        graph_list = []
        for size in self.samples:
            self._calls += 1
            neighbors, neighbors_count = sample_neighbors(self.row, self.colptr, nodes, sample_size=size,
                                                          seed=self.seed + self._calls)
            edge_src, edge_dst, sample_index = reindex_graph(nodes, neighbors, neighbors_count,
                                                             num_nodes=self.num_nodes)
            subgraph = Graph(num_nodes=int(sample_index.shape[0]),
                             edges=torch.stack([edge_src, edge_dst], dim=1))
            graph_list.append((subgraph, int(nodes.shape[0])))
            nodes = sample_index
        return graph_list[::-1], nodes
