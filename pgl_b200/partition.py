"""Graph partitioning (mirror of reference pgl/partition.py:25-123).

``metis_partition`` calls METIS K-way through the C-ABI (``pglb_metis_partition``), which
dlopen()s a libmetis built with 64-bit indices -- by default the one ``pgl_b200.build`` compiles
from the reference's vendored METIS 5.1.0 sources.  Same inputs (dst-keyed CSR, scaled integer
weights, default options => fixed seed) => same ``part[]`` as the reference.
"""
import ctypes
import logging
import math
import os

import numpy as np

from . import _lib
from .utils.helper import check_is_tensor

log = logging.getLogger("pgl")

__all__ = ["metis_partition", "random_partition", "block_partition"]


def _metis_weight_scale(X):
    """Positive integer weights 1..1001; reference partition.py:25-34."""
    X_min = np.min(X)
    X_max = np.max(X)
    X_scaled = (X - X_min) / (X_max - X_min + 1e-5)
    X_scaled = (X_scaled * 1000).astype("int64") + 1
    assert np.any(X_scaled > 0), "The weight of METIS input must be postive integers"
    return X_scaled


def _np(x):
    if check_is_tensor(x):
        return x.cpu().numpy()
    return np.asarray(x)


def metis_csr(num_nodes, indptr, adjncy, npart, node_weights=None, edge_weights=None,
              recursive=False, libmetis_path=None):
    """Raw call with the argument meaning of graph_kernel.metis_partition
    (reference pgl/graph_kernel.pyx:434-472)."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    adjncy = np.ascontiguousarray(adjncy, dtype=np.int64)
    part = np.zeros(int(num_nodes), dtype=np.int64)
    nw = np.ascontiguousarray(node_weights, dtype=np.int64) if node_weights is not None else None
    ew = np.ascontiguousarray(edge_weights, dtype=np.int64) if edge_weights is not None else None
    path = libmetis_path or os.environ.get("PGLB_LIBMETIS") or _lib.METIS_PATH
    p = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None  # noqa: E731
    _lib.check(_lib.lib.pglb_metis_partition(path.encode(), int(num_nodes), p(indptr), p(adjncy),
                                             int(npart), p(nw), p(ew), 1 if recursive else 0,
                                             p(part)))
    return part


def metis_partition(graph, npart, node_weights=None, edge_weights=None):
    """METIS K-way partition of an (undirected) graph -> int64 part id per node;
    reference partition.py:37-91."""
    log.warning("The input graph of metis_partition should be undirected.")
    n = int(graph._n)
    if npart == 1:
        return np.zeros(n, dtype=np.int64)
    csr = graph.adj_dst_index.numpy(inplace=False)
    indptr = csr._indptr
    v = csr._sorted_v
    sorted_eid = csr._sorted_eid
    if edge_weights is not None:
        edge_weights = _np(edge_weights)[np.asarray(sorted_eid)]
        edge_weights = _metis_weight_scale(edge_weights)
    if node_weights is not None:
        node_weights = _metis_weight_scale(_np(node_weights))
    # K-way only, as the reference (partition.py:81-90: "recursive metis always core dump")
    return metis_csr(n, indptr, v, npart, node_weights=node_weights, edge_weights=edge_weights,
                     recursive=False)


def random_partition(graph, npart):
    """Equal-size random parts; reference partition.py:94-123."""
    n = int(graph._n)
    if npart == 1:
        return np.zeros(n, dtype=np.int64)
    cs = int(math.ceil(n / npart))
    part_id = np.repeat(np.arange(npart, dtype=np.int64), cs)[:n]
    np.random.shuffle(part_id)
    return part_id


def block_partition(num_nodes, npart):
    """Contiguous equal blocks of the node id space (not in the reference): the zero-cost
    partition used for graphs whose ids are already randomly permuted or pre-clustered."""
    n = int(num_nodes)
    cs = int(math.ceil(n / npart)) if npart > 0 else n
    return np.minimum(np.arange(n, dtype=np.int64) // max(cs, 1), npart - 1)
