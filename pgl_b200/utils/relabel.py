"""Node / edge relabelling (mirror of graph_kernel.map_nodes / map_edges, reference
pgl/graph_kernel.pyx:104-138, used by pgl/sampling/custom.py:66-68 and by partition -> local-graph
pipelines).  The reference walks a C++ unordered_map per element; here the mapping is a dense
lookup table new_id[old_id] (numpy on the host; in tensor mode the kernels of csrc/localgraph.cu:
pglb_map_nodes / pglb_map_edges), which is what the partition code needs and is O(1) per element
without hashing."""
import numpy as np
import torch

__all__ = ["dense_table", "map_nodes", "map_edges"]


def dense_table(reindex, size=None, device=None):
    """dict {old: new} (the reference's argument type) or (old_ids, new_ids) arrays -> lookup table."""
    if isinstance(reindex, dict):
        old = np.fromiter(reindex.keys(), dtype=np.int64, count=len(reindex))
        new = np.fromiter(reindex.values(), dtype=np.int64, count=len(reindex))
    else:
        old, new = (np.asarray(a, dtype=np.int64) for a in reindex)
    n = int(size) if size is not None else (int(old.max()) + 1 if len(old) else 0)
    table = np.full(n, -1, dtype=np.int64)
    table[old] = new
    if device is not None:
        return torch.from_numpy(table).to(device)
    return table


def map_nodes(nodes, reindex):
    """new id of every node in `nodes` (reference graph_kernel.pyx:123-138)."""
    if isinstance(nodes, torch.Tensor):
        table = reindex if isinstance(reindex, torch.Tensor) else dense_table(reindex, device=nodes.device)
        if nodes.is_cuda:
            from .. import ops
            return ops.map_nodes(nodes.reshape(-1), table, strict=False).reshape(nodes.shape)
        return table.index_select(0, nodes.reshape(-1)).reshape(nodes.shape)
    nodes = np.asarray(nodes, dtype=np.int64)
    table = reindex if isinstance(reindex, np.ndarray) else dense_table(reindex)
    return table[nodes]


def map_edges(eid, edges, reindex):
    """edges[eid] with both endpoints relabelled (reference graph_kernel.pyx:104-120)."""
    if isinstance(edges, torch.Tensor):
        table = reindex if isinstance(reindex, torch.Tensor) else dense_table(reindex, device=edges.device)
        if edges.is_cuda:
            from .. import ops
            return ops.map_edges(eid, edges, table, strict=False)
        sel = edges.index_select(0, eid)
        return table[sel]
    edges = np.asarray(edges, dtype=np.int64)
    table = reindex if isinstance(reindex, np.ndarray) else dense_table(reindex)
    return table[edges[np.asarray(eid, dtype=np.int64)]]
