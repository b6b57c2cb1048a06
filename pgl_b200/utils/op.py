"""Mirror of reference pgl/utils/op.py for the send/recv path."""
import numpy as np
import torch

from .. import ops
from .helper import check_is_tensor


class LazyRows(torch.Tensor):
    """``base[index]`` that is only gathered when something other than a segment reduce touches
    it.  ``Graph.send`` hands these to the user's message function and ``Graph.recv`` composes
    them with the sorted edge ids, so the README pattern ``send(copy src) -> recv(reduce_sum)``
    lowers to ONE fused gather+reduce kernel instead of materialising two [E, D] tensors
    (reference graph.py:763-768 + 821-830: gather, re-gather by eid, segment op, zero+scatter,
    about 5x the bytes -- SURVEY.md section 3.2).  Any torch op on the object materialises it
    (once) and proceeds on the plain tensor, so user code sees an ordinary tensor."""

    _PASSIVE = None

    @staticmethod
    def __new__(cls, base, index):
        shape = (int(index.shape[0]),) + tuple(base.shape[1:])
        return torch.Tensor._make_wrapper_subclass(cls, shape, dtype=base.dtype, device=base.device,
                                                   requires_grad=False)

    def __init__(self, base, index):
        self._lz_base = base
        self._lz_index = index
        self._lz_value = None

    def materialize(self):
        if self._lz_value is None:
            self._lz_value = ops.gather_rows(self._lz_base, self._lz_index)
        return self._lz_value

    def is_lazy(self):
        return self._lz_value is None

    @classmethod
    def _passive(cls):
        if cls._PASSIVE is None:
            T = torch.Tensor
            cls._PASSIVE = {T.shape.__get__, T.dtype.__get__, T.device.__get__, T.is_cuda.__get__,
                            T.ndim.__get__, T.dim, T.size, T.numel, T.requires_grad.__get__,
                            T.layout.__get__, T.is_sparse.__get__}
        return cls._PASSIVE

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in cls._passive():
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)

        def unwrap(a):
            if isinstance(a, LazyRows):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(unwrap(v) for v in a)
            if isinstance(a, dict):
                return {k: unwrap(v) for k, v in a.items()}
            return a

        with torch._C.DisableTorchFunctionSubclass():
            return func(*unwrap(args), **unwrap(kwargs))

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # only reached if a wrapper slips past __torch_function__: same rule, materialise first
        def unwrap(a):
            if isinstance(a, LazyRows):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(unwrap(v) for v in a)
            return a

        return func(*unwrap(args), **{k: unwrap(v) for k, v in (kwargs or {}).items()})

    def __repr__(self):
        return "LazyRows(shape=%s, materialized=%s)" % (tuple(self.shape), self._lz_value is not None)


def read_rows(data, index):
    """reference pgl/utils/op.py:24-45: recursive row gather (paddle.gather).  float32 CUDA rows
    come back lazy (LazyRows); a lazy input is composed (index of index) instead of gathered."""
    if data is None:
        return None
    if isinstance(data, dict):
        return {key: read_rows(value, index) for key, value in data.items()}
    if isinstance(data, LazyRows) and data.is_lazy():
        idx = data._lz_index.index_select(0, index.contiguous())
        return LazyRows(data._lz_base, idx)
    if isinstance(data, torch.Tensor) and data.is_cuda and data.dtype == torch.float32 \
            and index.dtype == torch.int64 and data.dim() >= 1:
        return LazyRows(data, index)
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
