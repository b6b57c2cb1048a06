"""Segment operators (mirror of reference pgl/math.py:30-224) on the sm_100a kernels.

``paddle.geometric.segment_*`` -> ``pglb_spmm_csr_f32`` with identity columns (a sorted-segment
reduce is a CSR row reduce whose slot j reads row j); ``segment_softmax`` (7 Paddle ops in the
reference, math.py:216-224) -> one fused ``pglb_edge_softmax_csr_f32`` launch.
"""
import torch

from . import ops

__all__ = [
    "segment_pool", "segment_sum", "segment_mean", "segment_max", "segment_min",
    "segment_softmax",
]


def _indptr_of(segment_ids):
    """Compact CSR over the segments.  Graph.get_segment_ids attaches the cached one so the
    UDF recv path never re-derives it (and never syncs)."""
    cached = getattr(segment_ids, "_pglb_indptr", None)
    if cached is not None:
        return cached, getattr(segment_ids, "_pglb_max_degree", -1)
    ids = segment_ids if segment_ids.dtype == torch.int64 else segment_ids.to(torch.int64)
    n = int(ids.shape[0])
    k = int(ids[-1].item()) + 1 if n else 0
    return ops.segment_indptr(ids, k), -1


def segment_pool(data, segment_ids, pool_type, name=None):
    """reference pgl/math.py:30-46."""
    pool_type = pool_type.upper()
    if pool_type not in ("SUM", "MEAN", "MAX", "MIN"):
        raise ValueError(
            "We only support sum, mean, max, min pool types in segment_pool function.")
    ops.require_cuda(data, segment_ids)
    indptr, maxdeg = _indptr_of(segment_ids)
    from .utils.op import LazyRows
    if isinstance(data, LazyRows) and data.is_lazy():
        # untouched message rows: fuse the gather into the reduce (no [E, D] materialisation)
        return ops.segment_reduce(data._lz_base, None, pool_type.lower(), indptr=indptr,
                                  cols=data._lz_index.contiguous(), max_degree=maxdeg)
    return ops.segment_reduce(data, None, pool_type.lower(), indptr=indptr, max_degree=maxdeg)


def segment_sum(data, segment_ids, name=None):
    """reference pgl/math.py:49-79: out_i = sum_j data_j over segment_ids[j] == i."""
    return segment_pool(data, segment_ids, "sum")


def segment_mean(data, segment_ids, name=None):
    """reference pgl/math.py:82-113."""
    return segment_pool(data, segment_ids, "mean")


def segment_min(data, segment_ids, name=None):
    """reference pgl/math.py:116-145."""
    return segment_pool(data, segment_ids, "min")


def segment_max(data, segment_ids, name=None):
    """reference pgl/math.py:148-178."""
    return segment_pool(data, segment_ids, "max")


def segment_softmax(data, segment_ids):
    """reference pgl/math.py:181-224: exp(x - segmax) / segsum, fused into one kernel."""
    ops.require_cuda(data, segment_ids)
    indptr, _ = _indptr_of(segment_ids)
    return ops.edge_softmax_csr(indptr, None, data, int(data.shape[0]))
