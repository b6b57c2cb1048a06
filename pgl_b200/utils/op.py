"""Mirror of reference pgl/utils/op.py for the send/recv path."""
import numpy as np
import torch

from .. import ops
from .helper import check_is_tensor


def read_rows(data, index):
    """reference pgl/utils/op.py:24-45: recursive row gather (paddle.gather -> gather_rows)."""
    if data is None:
        return None
    if isinstance(data, dict):
        return {key: read_rows(value, index) for key, value in data.items()}
    return ops.gather_rows(data, index)


def get_index_from_counts(counts):
    """reference pgl/utils/op.py:48-72: [2,3,4] -> [0,2,5,9]."""
    if check_is_tensor(counts):
        return torch.cat([torch.zeros(1, dtype=counts.dtype, device=counts.device),
                          torch.cumsum(counts, 0)], dim=-1)
    index = np.cumsum(counts, dtype="int64")
    return np.insert(index, 0, 0)


class RowReader(dict):
    """reference pgl/utils/op.py:75-87 -- lazy, memoised per-key row gather."""

    def __init__(self, nfeat, index):
        super().__init__()
        self.nfeat = nfeat
        self.loaded_nfeat = {}
        self.index = index

    def __getitem__(self, key):
        if key not in self.loaded_nfeat:
            self.loaded_nfeat[key] = read_rows(self.nfeat[key], self.index)
        return self.loaded_nfeat[key]


class _AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, group):
        import torch.distributed as dist
        ctx.group = group
        out = tensor.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        # the registered gradient of c_allreduce_sum is c_allreduce_sum of the upstream gradient
        import torch.distributed as dist
        g = g.clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


def all_reduce_sum_with_grad(tensor, group=None):
    """reference pgl/utils/op.py:90-122 (c_allreduce_sum on the calc stream, differentiable)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensor
    if tensor.requires_grad and torch.is_grad_enabled():
        return _AllReduceSum.apply(tensor, group)
    out = tensor.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
