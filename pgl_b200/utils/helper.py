"""Host-side helpers mirroring reference pgl/utils/helper.py (tensor = CUDA torch.Tensor)."""
import numpy as np
import torch

from .. import ops


def check_is_tensor(*data):
    """reference pgl/utils/helper.py:23-29 (paddle.Tensor -> torch.Tensor)."""
    for d in data:
        if isinstance(d, torch.Tensor):
            return True
    return False


def cuda_device():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "pgl_b200: tensor mode needs a CUDA device (B200); no CPU tensor mode exists. "
            "Use the numpy-mode Graph for host-side index queries.")
    return torch.device("cuda", torch.cuda.current_device())


def to_tensor(data, device=None):
    """reference to_paddle_tensor (pgl/utils/helper.py:32-43); UVA mode is out of scope."""
    if isinstance(data, torch.Tensor):
        return data if data.is_cuda else data.to(device or cuda_device())
    dev = device or cuda_device()
    arr = np.ascontiguousarray(np.asarray(data))
    return torch.from_numpy(arr).to(dev)


def maybe_num_nodes(edges):
    """reference pgl/utils/helper.py:133-153."""
    if len(edges) == 0:
        return 0
    if check_is_tensor(edges):
        return torch.max(edges) + 1
    return np.max(edges) + 1


def unique_segment(data, dtype="int64"):
    """reference pgl/utils/helper.py:156-160 for a SORTED key vector (the only use on the
    send/recv path): (unique values, dense non-decreasing segment id per element)."""
    ops.require_cuda(data)
    n = int(data.shape[0])
    if n == 0:
        z = torch.zeros(0, dtype=torch.int64, device=data.device)
        return z, z.clone()
    num = int(data[-1].item()) + 1
    indptr = ops.segment_indptr(data, num)
    return ops.segment_ids_from_indptr(indptr, n)


def generate_segment_id_from_index(index):
    """reference pgl/utils/helper.py:116-130: graph_node_id from a [G+1] offset vector (int32)."""
    if check_is_tensor(index):
        counts = (index[1:] - index[:-1]).to(torch.int64)
        return torch.repeat_interleave(
            torch.arange(counts.shape[0], device=index.device, dtype=torch.int32), counts)
    index = np.asarray(index)
    segments = np.zeros(index[-1] + 1, dtype="int32")
    np.add.at(segments, index[:-1], 1)
    return np.cumsum(segments)[:-1] - 1


def scatter(x, index, updates, overwrite=True, name=None):
    """reference pgl/utils/helper.py:46-113 -> paddle.scatter."""
    if overwrite:
        return ops.scatter_rows(x, index, updates)
    out = x.clone()
    out.index_fill_(0, index, 0)
    out.index_add_(0, index, updates)
    return out


def graph_send_recv(x, src_index, dst_index, pool_type="sum"):
    """reference pgl/utils/helper.py:163-210: copy-source send + built-in reduce on a bare COO
    edge list (out rows = x rows).  The reference gathers [E, D] and scatter-adds it (sum only);
    here the dst-CSR is built on the device (stable, so the per-row order is edge order) and the
    fused aggregation kernel runs on it -- mean / max / min come with it.  Nothing is cached:
    callers that reuse the edge list should hold a Graph instead."""
    assert pool_type in ("sum", "mean", "max", "min"), \
        "Only support 'sum', 'mean', 'max', 'min' pool_type."
    ops.require_cuda(x, src_index, dst_index)
    n = int(x.shape[0])
    degree, cols, _, eid, indptr = ops.csr_build(dst_index, src_index, n)
    fwd = {"indptr": indptr, "cols": cols, "eid": eid, "degree": degree, "max_degree": -1}
    bwd = None
    if x.requires_grad and torch.is_grad_enabled():
        def bwd():
            d2, c2, _, e2, ip2 = ops.csr_build(src_index, dst_index, n)
            return {"indptr": ip2, "cols": c2, "eid": e2, "degree": d2, "max_degree": -1}
    return ops.aggregate_copy(x, fwd, n, pool_type, bwd=bwd)
