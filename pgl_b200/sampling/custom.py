"""``pgl.sampling.subgraph`` (reference pgl/sampling/custom.py:23-83): the subgraph over ``nodes`` holding the edges
``eid`` of the parent graph (or an explicit edge list), nodes renumbered from 0 in the order given.  Host side, numpy
mode only, like the reference.  The reference builds a Python dict {node: position} and hands it to the Cython
``map_edges`` (copied into an unordered_map, walked per edge); here the renumbering is the dense table of
``utils.relabel`` (its device twin is ``pglb_map_edges``, used by the partition pipeline)."""
import numpy as np

from ..utils import relabel

__all__ = ["subgraph"]


def subgraph(graph, nodes, eid=None, edges=None, with_node_feat=True, with_edge_feat=True):
    """Every endpoint of the selected edges must be in ``nodes`` (the reference warns in capitals; an endpoint that is
    not maps to -1 here, to a silently inserted 0 there).  Node features follow ``nodes``, edge features ``eid``."""
    from ..graph import Graph
    assert not graph.is_tensor(), "You must call Graph.numpy() first."
    if eid is None and edges is None:
        raise ValueError("Eid and edges can't be None at the same time.")
    nodes_np = np.asarray(nodes, dtype=np.int64)
    table = np.full(int(graph.num_nodes), -1, dtype=np.int64)
    table[nodes_np] = np.arange(len(nodes_np), dtype=np.int64)
    if edges is None:
        sel_e = np.asarray(eid, dtype=np.int64)
        src_edges = np.asarray(graph._edges)[sel_e]
    else:
        sel_e = None if eid is None else np.asarray(eid, dtype=np.int64)
        src_edges = np.array(edges, dtype="int64").reshape(-1, 2)
    sub_edges = relabel.map_edges(np.arange(len(src_edges), dtype=np.int64), src_edges, table)
    sub_edge_feat = {}
    if with_edge_feat:
        for key, value in graph.edge_feat.items():
            if sel_e is None:
                raise ValueError("Eid can not be None with edge features.")
            sub_edge_feat[key] = value[sel_e]
    sub_node_feat = {}
    if with_node_feat:
        for key, value in graph.node_feat.items():
            sub_node_feat[key] = value[nodes_np]
    return Graph(edges=sub_edges, num_nodes=len(nodes_np), node_feat=sub_node_feat, edge_feat=sub_edge_feat)
