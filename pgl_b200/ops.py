"""Torch-tensor front end of the C-ABI (include/pglb.h): allocates outputs / workspaces with
torch, passes raw device pointers + the current CUDA stream to libpglb.  CUDA tensors only --
every function raises on CPU tensors; there is no eager / CPU fallback on this path.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import BCAST_FULL, BCAST_HEAD, BCAST_SCALAR, MSG, REDUCE, check, lib

_ws_cache = {}


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "pgl_b200: tensor-mode ops run on CUDA tensors only (got a %s tensor); "
                "there is no CPU fallback on the send/recv path" % t.device)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def workspace(device, nbytes):
    """Per-(device, stream) scratch buffer, grown geometrically; stream-ordered reuse."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        n = max(int(nbytes), 1 << 20)
        if buf is not None:
            n = max(n, int(buf.numel() * 1.5))
        buf = torch.empty(n, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _i64(t):
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    return t


def f64_through(fn):
    """float64 in -> float64 out for the public feature ops (the reference's tests/test_conv.py:43-71 runs GCN,
    GraphSage and GAT in float64).  The kernels are fp32: float64 tensors are rounded to float32 on the way in and
    the result is widened on the way out (autograd flows through both casts), so a float64 caller gets fp32-accurate
    values in its own dtype -- within the path's stated 1e-4 relative tolerance, not fp64 accuracy."""
    import functools

    def is64(t):
        return isinstance(t, torch.Tensor) and t.dtype == torch.float64

    def down(v):
        mat = getattr(v, "materialize", None)
        if mat is not None and getattr(v, "dtype", None) == torch.float64:
            v = mat()
        if is64(v):
            return v.to(torch.float32)
        if isinstance(v, (list, tuple)):
            return type(v)(down(a) for a in v)
        return v

    def any64(v):
        if is64(v) or (hasattr(v, "materialize") and getattr(v, "dtype", None) == torch.float64):
            return True
        return isinstance(v, (list, tuple)) and any(any64(a) for a in v)

    def up(v):
        if isinstance(v, torch.Tensor) and v.dtype == torch.float32:
            return v.to(torch.float64)
        if isinstance(v, tuple):
            return tuple(up(a) for a in v)
        return v

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if not (any(any64(a) for a in args) or any(any64(a) for a in kwargs.values())):
            return fn(*args, **kwargs)
        out = fn(*[down(a) for a in args], **{k: down(a) for k, a in kwargs.items()})
        return up(out)
    return wrapper


def _f32_2d(t):
    """[n, d1, d2, ...] -> contiguous [n, D] float32 view (no copy when already so).  A lazy row
    gather (utils.op.LazyRows) is resolved to its plain tensor first: the wrapper itself carries no
    autograd state, so requires_grad must be read from the gathered result."""
    mat = getattr(t, "materialize", None)
    if mat is not None:
        t = mat()
    if t.dtype != torch.float32:
        raise TypeError("pgl_b200: float32 features expected on the CUDA path, got %s" % t.dtype)
    t2 = t.reshape(t.shape[0], -1) if t.dim() != 2 else t
    if not t2.is_contiguous():
        t2 = t2.contiguous()
    return t2


# ------------------------------------------------------------------------------------------
# index construction
# ------------------------------------------------------------------------------------------


def csr_build(u, v, num_nodes):
    """Device twin of graph_kernel.build_index (reference pgl/graph_kernel.pyx:59-88).
    u, v: 1-D int64 CUDA tensors (strided views such as edges[:, 1] are read in place).
    Returns (degree, sorted_v, sorted_u, sorted_eid, indptr) like the reference."""
    require_cuda(u, v)
    u = _i64(u)
    v = _i64(v)
    E = int(u.shape[0])
    N = int(num_nodes)
    dev = u.device
    degree = torch.empty(N, dtype=torch.int64, device=dev)
    indptr = torch.empty(N + 1, dtype=torch.int64, device=dev)
    su = torch.empty(E, dtype=torch.int64, device=dev)
    sv = torch.empty(E, dtype=torch.int64, device=dev)
    se = torch.empty(E, dtype=torch.int64, device=dev)
    need = ctypes.c_size_t(0)
    check(lib.pglb_csr_build_ws(E, N, ctypes.byref(need)))
    ws = workspace(dev, need.value)
    us = u.stride(0) if E > 0 else 1
    vs = v.stride(0) if E > 0 else 1
    with torch.cuda.device(dev):
        check(lib.pglb_csr_build(_ptr(u), max(us, 1), _ptr(v), max(vs, 1), E, N, _ptr(degree),
                                 _ptr(indptr), _ptr(su), _ptr(sv), _ptr(se), _ptr(ws),
                                 ws.numel(), _stream()))
    return degree, sv, su, se, indptr


# partition -> local graph (csrc/localgraph.cu)
# ------------------------------------------------------------------------------------------


def map_nodes(nodes, table, strict=True):
    """graph_kernel.map_nodes on the device: table[nodes] for a dense table new_id[old_id] (int64 CUDA tensors).
    strict: ids outside the table raise (one 4-byte read back); otherwise they map to -1."""
    require_cuda(nodes, table)
    nodes = _i64(nodes).contiguous()
    table = _i64(table).contiguous()
    out = torch.empty_like(nodes)
    bad = torch.zeros(1, dtype=torch.int32, device=nodes.device)
    with torch.cuda.device(nodes.device):
        check(lib.pglb_map_nodes(_ptr(nodes), nodes.numel(), _ptr(table), table.numel(), _ptr(out), _ptr(bad),
                                 _stream()))
    if strict and int(bad.item()):
        raise IndexError("pgl_b200.map_nodes: node id outside the relabelling table")
    return out


def map_edges(eid, edges, table, strict=True):
    """graph_kernel.map_edges on the device: table[edges[eid]] (eid None = every edge in order)."""
    require_cuda(edges, table)
    edges = _i64(edges).contiguous()
    table = _i64(table).contiguous()
    if edges.dim() != 2 or edges.shape[1] != 2:
        raise ValueError("pgl_b200.map_edges: edges must be [E, 2]")
    if eid is not None:
        eid = _i64(eid).contiguous()
    n = int(eid.numel()) if eid is not None else int(edges.shape[0])
    out = torch.empty((n, 2), dtype=torch.int64, device=edges.device)
    bad = torch.zeros(1, dtype=torch.int32, device=edges.device)
    with torch.cuda.device(edges.device):
        check(lib.pglb_map_edges(_ptr(eid), n, _ptr(edges), int(edges.shape[0]), _ptr(table), table.numel(),
                                 _ptr(out), _ptr(bad), _stream()))
    if strict and int(bad.item()):
        raise IndexError("pgl_b200.map_edges: edge or node id outside its table")
    return out


def partition_relabel(part, num_parts):
    """part[N] (int64 CUDA tensor, values in [0, num_parts)) -> (new_id[N], offsets[num_parts + 1]) such that every
    part is a contiguous range of new ids, stable inside a part (reference apps/GNNAutoScale/graph_partition.py:94-101:
    ``permutation = argsort(part)`` + offsets).  The stable sort is pglb_csr_build's radix sort with u = part."""
    require_cuda(part)
    part = _i64(part).contiguous()
    n = int(part.numel())
    deg, sv, su, perm, offsets = csr_build(part, part, int(num_parts))
    new_id = torch.empty(n, dtype=torch.int64, device=part.device)
    with torch.cuda.device(part.device):
        check(lib.pglb_invert_perm(_ptr(perm), n, _ptr(new_id), _stream()))
    return new_id, offsets


def halo_plan(edges, num_nodes, lo, hi, offsets):
    """One rank's local graph of a contiguous 1-D node partition (pglb_halo_plan_count / _fill).
    edges [E, 2] int64 CUDA (global ids), this rank owns [lo, hi), offsets = python list / tensor of K + 1 part
    starts.  Returns (eid, dst_local, col_local, halo_ids, recv_counts) -- all int64 CUDA tensors."""
    require_cuda(edges)
    edges = _i64(edges).contiguous()
    dev = edges.device
    E, N = int(edges.shape[0]), int(num_nodes)
    off_t = torch.as_tensor(offsets, dtype=torch.int64).to(dev).contiguous()
    K = int(off_t.numel()) - 1
    need = ctypes.c_size_t(0)
    check(lib.pglb_halo_plan_ws(E, N, ctypes.byref(need)))
    ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)  # private: must survive between the two calls
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.pglb_halo_plan_count(_ptr(edges), E, N, int(lo), int(hi), _ptr(counts), _ptr(bad), _ptr(ws),
                                       ws.numel(), _stream()))
        e_loc, n_halo = (int(v) for v in counts.tolist())
        if int(bad.item()):
            raise IndexError("pgl_b200.halo_plan: edge endpoint outside [0, num_nodes)")
        eid = torch.empty(e_loc, dtype=torch.int64, device=dev)
        dst_local = torch.empty(e_loc, dtype=torch.int64, device=dev)
        col_local = torch.empty(e_loc, dtype=torch.int64, device=dev)
        halo_ids = torch.empty(n_halo, dtype=torch.int64, device=dev)
        recv_counts = torch.zeros(K, dtype=torch.int64, device=dev)
        check(lib.pglb_halo_plan_fill(_ptr(edges), E, N, int(lo), int(hi), _ptr(off_t), K, _ptr(eid),
                                      _ptr(dst_local), _ptr(col_local), _ptr(halo_ids), _ptr(recv_counts), _ptr(ws),
                                      ws.numel(), _stream()))
    return eid, dst_local, col_local, halo_ids, recv_counts


def segment_ids_from_indptr(indptr, num_edges):
    """(uniq_ind, segment_ids) of a CSR == paddle.unique(sorted_key, return_inverse=True)
    (reference pgl/utils/helper.py:156-160).  One D2H read of K (the reference's unique syncs too)."""
    require_cuda(indptr)
    N = int(indptr.shape[0]) - 1
    E = int(num_edges)
    dev = indptr.device
    uniq = torch.empty(max(N, 0), dtype=torch.int64, device=dev)
    seg = torch.empty(E, dtype=torch.int64, device=dev)
    k = torch.zeros(1, dtype=torch.int64, device=dev)
    need = ctypes.c_size_t(0)
    check(lib.pglb_segment_ids_ws(N, ctypes.byref(need)))
    ws = workspace(dev, need.value)
    with torch.cuda.device(dev):
        check(lib.pglb_segment_ids(_ptr(indptr), N, E, _ptr(uniq), _ptr(seg), _ptr(k), _ptr(ws),
                                   ws.numel(), _stream()))
    return uniq[: int(k.item())], seg


def segment_indptr(segment_ids, num_segments):
    require_cuda(segment_ids)
    ids = _i64(segment_ids).contiguous()
    E = int(ids.shape[0])
    K = int(num_segments)
    indptr = torch.empty(K + 1, dtype=torch.int64, device=ids.device)
    with torch.cuda.device(ids.device):
        check(lib.pglb_segment_indptr(_ptr(ids), E, K, _ptr(indptr), _stream()))
    return indptr


# ------------------------------------------------------------------------------------------
# aggregation
# ------------------------------------------------------------------------------------------


def _spmm_raw(indptr, cols, x2, n_dst, reduce_op, eid=None, y2=None, y_bcast=BCAST_FULL,
              head_dim=1, msg_op="copy", scale_src=None, scale_dst=None, max_degree=-1,
              num_edges=None, out=None, packed=None, accumulate=False):
    dev = x2.device
    D = int(x2.shape[1])
    E = int(num_edges if num_edges is not None else (cols.shape[0] if cols is not None else 0))
    if out is None:
        out = torch.empty((n_dst, D), dtype=torch.float32, device=dev)
    if isinstance(packed, NarrowPlan):
        if (reduce_op in ("sum", "mean") and msg_op == "copy" and y2 is None and not accumulate and E > 0
                and n_dst > 0 and x2.data_ptr() % 16 == 0 and x2.stride(0) % 4 == 0 and x2.stride(1) == 1
                and out.data_ptr() % 16 == 0 and out.stride(0) % 4 == 0 and int(x2.shape[0]) < (1 << 30)):
            need = ctypes.c_size_t(0)
            check(lib.pglb_spmm_narrow_ws(E, D, ctypes.byref(need)))
            ws = workspace(dev, need.value)
            sval = None
            if scale_src is not None and NARROW_SLOT_SCALE and packed.slot_scale is not None:
                sval = packed.slot_scale(scale_src)     # scale_src[cols[j]] per slot, cached per scale tensor
            with torch.cuda.device(dev):
                check(lib.pglb_spmm_narrow_f32(_ptr(packed.plan), _ptr(packed.nz_row), _ptr(packed.blk_k),
                                               _ptr(indptr), _ptr(x2), x2.stride(0), _ptr(out), out.stride(0),
                                               n_dst, int(x2.shape[0]), E, D, REDUCE[reduce_op], _ptr(scale_src),
                                               _ptr(sval), _ptr(scale_dst), 2 if packed.hints else 0, _ptr(ws),
                                               ws.numel(), _stream()))
            return out
        packed = packed.fallback() if packed.fallback is not None else None
    need = ctypes.c_size_t(0)
    check(lib.pglb_spmm_csr_ws(n_dst, E, D, ctypes.byref(need)))
    ws = workspace(dev, need.value)
    wsn = ws.numel()
    with torch.cuda.device(dev):
        check(lib.pglb_spmm_csr_f32(
            _ptr(indptr), _ptr(cols), _ptr(eid), _ptr(x2), x2.stride(0), _ptr(y2),
            (y2.stride(0) if y2 is not None else 0), y_bcast, _ptr(out), out.stride(0), n_dst,
            int(x2.shape[0]), E, D, head_dim, MSG[msg_op], REDUCE[reduce_op], _ptr(scale_src),
            _ptr(scale_dst), _ptr(packed[0]) if packed is not None else None, int(max_degree),
            (1 if accumulate else 0) | (2 if (packed is not None and packed[1]) else 0), _ptr(ws),
            wsn, _stream()))
    return out


# PGLB_HOST_COPY=kernel (experimental): move the column blocks of HostAggregator with a zero-copy kernel
# instead of cudaMemcpy2DAsync (csrc/host_copy.cu); PGLB_HOST_COPY_CTAS sets its grid (default 16)
HOST_COPY_KERNEL = os.environ.get("PGLB_HOST_COPY") == "kernel"
HOST_COPY_CTAS = int(os.environ.get("PGLB_HOST_COPY_CTAS", "16"))


def _copy2d(dst_ptr, dpitch, src_ptr, spitch, width, height, kind, stream):
    """kind 1: host -> device, 2: device -> host (pglb_memcpy2d_async's convention)."""
    if HOST_COPY_KERNEL:
        check(lib.pglb_copy2d_kernel_async(ctypes.c_void_p(dst_ptr), dpitch, ctypes.c_void_p(src_ptr), spitch,
                                           width, height, HOST_COPY_CTAS, ctypes.c_void_p(stream)))
    else:
        check(lib.pglb_memcpy2d_async(ctypes.c_void_p(dst_ptr), dpitch, ctypes.c_void_p(src_ptr), spitch,
                                      width, height, kind, ctypes.c_void_p(stream)))


class HostAggregator(object):
    """send_u_recv for features that live in (pinned) HOST memory: out_host = aggregate(x_host).

    The graph stays resident on the GPU; per call the [N, D] float32 matrix is streamed through the GPU
    in `chunks` column blocks on three streams -- upload of block c+1, aggregation of block c and
    download of block c-1 run concurrently (PCIe is full duplex, and the columns of a copy-message
    aggregation are independent).

    Two ways to call it:

    * ``agg(x_host, out_host, ...)``  BLOCKING: returns when ``out_host`` holds the result (the host
      thread has waited for the last device-to-host copy), so the caller may read ``out_host`` and
      rewrite ``x_host`` immediately.
    * ``t = agg.submit(x_host, out_host, ...)`` ... ``agg.wait(t)``  PIPELINED across calls: the
      aggregator owns ``depth`` device buffer sets, so the upload of call i+1 overlaps the kernel and
      the download of call i (a stream of feature matrices costs one direction of PCIe traffic per
      matrix).  Contract: ``x_host`` must not be written and ``out_host`` must not be read until
      ``wait(t)`` (or ``t.synchronize()``) has returned."""

    def __init__(self, fwd, n_src, n_dst, dim, device, chunks=2, depth=2):
        self.fwd, self.n_src, self.n_dst, self.dim = fwd, int(n_src), int(n_dst), int(dim)
        self.device = device
        chunks = max(1, min(int(chunks), self.dim // 4 if self.dim >= 4 else 1))
        w = -(-self.dim // chunks)
        w = (w + 3) // 4 * 4
        self.bounds = [(c, min(c + w, self.dim)) for c in range(0, self.dim, w)]
        self.depth = max(1, int(depth))
        self.sets = [None] * self.depth       # device buffer sets, allocated on first use
        self.count = 0
        self.s_in = torch.cuda.Stream(device=device)
        self.s_k = torch.cuda.Stream(device=device)
        self.s_out = torch.cuda.Stream(device=device)

    def _set(self, i):
        st = self.sets[i]
        if st is None:
            st = {"xd": torch.empty((self.n_src, self.dim), dtype=torch.float32, device=self.device),
                  "od": torch.empty((self.n_dst, self.dim), dtype=torch.float32, device=self.device),
                  "kernel_done": None, "d2h_done": None}
            self.sets[i] = st
        return st

    def submit(self, x_host, out_host, reduce_op="sum", scale_src=None, scale_dst=None):
        assert x_host.dtype == torch.float32 and out_host.dtype == torch.float32
        assert tuple(x_host.shape) == (self.n_src, self.dim) and x_host.is_contiguous()
        assert tuple(out_host.shape) == (self.n_dst, self.dim) and out_host.is_contiguous()
        st = self._set(self.count % self.depth)
        self.count += 1
        main = torch.cuda.current_stream(self.device)
        self.s_k.wait_stream(main)  # the scale vectors may have been produced on the caller's stream
        if st["kernel_done"] is not None:
            self.s_in.wait_event(st["kernel_done"])   # the previous user of this xd has been aggregated
        if st["d2h_done"] is not None:
            self.s_k.wait_event(st["d2h_done"])       # the previous result has left this od
        xd, od = st["xd"], st["od"]
        pitch = self.dim * 4
        full = len(self.bounds) == 1
        with torch.cuda.device(self.device):
            ev_in = []
            for lo, hi in self.bounds:
                if full:
                    with torch.cuda.stream(self.s_in):
                        xd.copy_(x_host, non_blocking=True)
                else:
                    _copy2d(xd.data_ptr() + lo * 4, pitch, x_host.data_ptr() + lo * 4, pitch,
                            (hi - lo) * 4, self.n_src, 1, self.s_in.cuda_stream)
                e = torch.cuda.Event()
                e.record(self.s_in)
                ev_in.append(e)
            for i, (lo, hi) in enumerate(self.bounds):
                self.s_k.wait_event(ev_in[i])
                with torch.cuda.stream(self.s_k):
                    _spmm_raw(self.fwd["indptr"], self.fwd["cols"], xd[:, lo:hi], self.n_dst, reduce_op,
                              scale_src=scale_src, scale_dst=scale_dst,
                              max_degree=self.fwd.get("max_degree", -1), out=od[:, lo:hi],
                              packed=_packed_of(self.fwd, xd[:, lo:hi]))
                e = torch.cuda.Event()
                e.record(self.s_k)
                self.s_out.wait_event(e)
                if full:
                    with torch.cuda.stream(self.s_out):
                        out_host.copy_(od, non_blocking=True)
                else:
                    _copy2d(out_host.data_ptr() + lo * 4, pitch, od.data_ptr() + lo * 4, pitch,
                            (hi - lo) * 4, self.n_dst, 2, self.s_out.cuda_stream)
                st["kernel_done"] = e
        done = torch.cuda.Event()
        done.record(self.s_out)
        st["d2h_done"] = done
        return done

    def wait(self, ticket):
        ticket.synchronize()

    def __call__(self, x_host, out_host, reduce_op="sum", scale_src=None, scale_dst=None):
        t = self.submit(x_host, out_host, reduce_op, scale_src, scale_dst)
        torch.cuda.current_stream(self.device).wait_event(t)
        t.synchronize()  # the host may touch out_host / x_host as soon as this returns
        return out_host


class IpcBuffer(object):
    """A cudaMalloc'ed float32 [rows, cols] buffer with a CUDA IPC handle (pglb_ipc_alloc), exposed
    to torch zero-copy through __cuda_array_interface__.  Used for the multi-GPU feature buffer
    whose rows peers pull over NVLink."""

    def __init__(self, rows, cols, device):
        self.shape = (int(rows), int(cols))
        self.device = device
        nbytes = max(self.shape[0] * self.shape[1] * 4, 256)
        ptr = ctypes.c_void_p()
        self.handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            check(lib.pglb_ipc_alloc(nbytes, ctypes.byref(ptr), self.handle))
        self.ptr = int(ptr.value)
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f4",
                                         "data": (self.ptr, False), "version": 2, "strides": None}
        self.tensor = torch.as_tensor(self, device=device)

    def handle_bytes(self):
        return bytes(self.handle.raw)

    def close(self):
        if self.ptr:
            self.tensor = None
            lib.pglb_ipc_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0


def ipc_open(handle_bytes, device):
    """Map a peer's IpcBuffer under `device`; returns the raw device pointer (int)."""
    ptr = ctypes.c_void_p()
    buf = ctypes.create_string_buffer(handle_bytes, 64)
    with torch.cuda.device(device):
        check(lib.pglb_ipc_open(buf, ctypes.byref(ptr)))
    return int(ptr.value)


def gather_rows_ptr(src_ptr, ld, index, out):
    """out[i] = src[index[i]] where src is a raw (possibly peer-mapped) float32 pointer."""
    n = int(index.shape[0])
    if n == 0:
        return out
    D = int(out.shape[1])
    with torch.cuda.device(out.device):
        check(lib.pglb_gather_rows_f32(ctypes.c_void_p(src_ptr), int(ld), _ptr(index),
                                       max(index.stride(0), 1), n, D, _ptr(out), out.stride(0),
                                       _stream()))
    return out


HOT_L2_BYTES = int(float(os.environ.get("PGLB_HOT_MB", "64")) * (1 << 20))


def pack_cols(cols, n_src, row_bytes, budget_bytes=None):
    """(packed uint32 column ids, has_hints) for the wide-row aggregation kernel, or None when
    ids do not fit 31 bits.  With a budget (PGLB_HOT_MB, default 32 MB; 0 = no hints) bit 31 marks the
    most frequently gathered sources, as many as fit in the budget, for an evict_last L2 policy.
    One-off per graph; cached by EdgeIndex."""
    budget = HOT_L2_BYTES if budget_bytes is None else int(budget_bytes)
    n_src = int(n_src)
    E = int(cols.shape[0])
    if n_src >= (1 << 31) - 1 or E == 0 or n_src == 0:
        return None
    dev = cols.device
    count = torch.empty(n_src, dtype=torch.int32, device=dev)
    packed = torch.empty(E, dtype=torch.int32, device=dev)  # uint32 payload
    k = budget // max(int(row_bytes), 1)
    thr = (1 << 62)
    hints = budget > 0 and 0 < k < n_src
    if hints:
        with torch.cuda.device(dev):
            check(lib.pglb_pack_cols(_ptr(cols), E, n_src, _ptr(count), 1 << 62, _ptr(packed), _stream()))
        thr = max(int(torch.topk(count, k, sorted=True).values[-1].item()), 2)
    with torch.cuda.device(dev):
        check(lib.pglb_pack_cols(_ptr(cols), E, n_src, _ptr(count), thr, _ptr(packed), _stream()))
    return packed, hints


class _CopyAgg(torch.autograd.Function):
    """out[d] = reduce_{e -> d} scale_src[src] * x[src]  (* scale_dst[d]); sum / mean.
    Backward is the same kernel on the reverse (src-keyed) CSR."""

    @staticmethod
    def forward(ctx, x2, fwd, bwd, n_dst, reduce_op, scale_src, scale_dst):
        ctx.bwd = bwd
        ctx.reduce_op = reduce_op
        ctx.n_src = int(x2.shape[0])
        ctx.scales = (scale_src, scale_dst)
        ctx.fwd = fwd
        return _spmm_raw(fwd["indptr"], fwd["cols"], x2, n_dst, reduce_op, scale_src=scale_src,
                         scale_dst=scale_dst, max_degree=fwd.get("max_degree", -1),
                         packed=_packed_of(fwd, x2))

    @staticmethod
    def backward(ctx, g):
        bwd = ctx.bwd() if callable(ctx.bwd) else ctx.bwd
        if bwd is None:
            raise RuntimeError("pgl_b200: backward of this aggregation needs the reverse CSR")
        g = g.contiguous()
        scale_src, scale_dst = ctx.scales
        s_in = scale_dst
        if ctx.reduce_op == "mean":
            inv = 1.0 / torch.clamp(ctx.fwd["degree"].to(torch.float32), min=1.0)
            s_in = inv if s_in is None else s_in * inv
        # the reverse CSR has one row per graph node; a feature matrix with MORE rows than the graph has nodes
        # (legal in the forward: the extra rows are never gathered) gets zero gradient rows, never an indptr over-read
        nb = int(bwd["indptr"].shape[0]) - 1
        rows = min(ctx.n_src, nb)
        sdst = scale_src[:rows] if (scale_src is not None and int(scale_src.shape[0]) > rows) else scale_src
        gx = _spmm_raw(bwd["indptr"][: rows + 1] if rows < nb else bwd["indptr"], bwd["cols"], g, rows, "sum",
                       scale_src=s_in, scale_dst=sdst, max_degree=bwd.get("max_degree", -1),
                       packed=_packed_of(bwd, g) if rows == nb else None)
        if ctx.n_src > rows:
            gx = torch.cat([gx, gx.new_zeros((ctx.n_src - rows, gx.shape[1]))], 0)
        return gx, None, None, None, None, None, None


class _MaxMinAgg(torch.autograd.Function):
    """send_u_recv(max|min); backward routes the gradient to every source entry that equals the
    reduced value (pglb_maxmin_bwd_f32 on the reverse CSR)."""

    @staticmethod
    def forward(ctx, x2, fwd, bwd, n_dst, reduce_op):
        out = _spmm_raw(fwd["indptr"], fwd["cols"], x2, n_dst, reduce_op,
                        max_degree=fwd.get("max_degree", -1), packed=_packed_of(fwd, x2))
        ctx.bwd = bwd
        ctx.save_for_backward(x2, out)
        return out

    @staticmethod
    def backward(ctx, g):
        bwd = ctx.bwd() if callable(ctx.bwd) else ctx.bwd
        if bwd is None:
            raise RuntimeError("pgl_b200: backward of this aggregation needs the reverse CSR")
        x2, out = ctx.saved_tensors
        g = g.contiguous()
        gx = torch.empty_like(x2)
        with torch.cuda.device(x2.device):
            check(lib.pglb_maxmin_bwd_f32(_ptr(bwd["indptr"]), _ptr(bwd["cols"]), _ptr(x2), _ptr(out),
                                          _ptr(g), _ptr(gx), int(x2.shape[0]), int(x2.shape[1]),
                                          _stream()))
        return gx, None, None, None, None


# rows of <= 64 floats: spmm_narrow2_kernel by default; PGLB_NARROW=1 PGLB_NARROW2=0 selects round 1's narrow kernel
NARROW2 = os.environ.get("PGLB_NARROW2", "1") != "0"          # spmm_narrow2_kernel for copy-sum / mean over rows of <= 64 floats
NARROW_ROWS = os.environ.get("PGLB_NARROW") == "1"            # round 1's spmm_narrow_kernel (opt-in, superseded)


NARROW_HOT_BYTES = int(float(os.environ.get("PGLB_NARROW_HOT_MB", "64")) * (1 << 20))   # L2 residency budget, narrow rows
NARROW_SLOT_SCALE = os.environ.get("PGLB_NARROW_SLOT_SCALE", "1") != "0"                 # stream cached per-slot norms


class NarrowPlan(object):
    """What pglb_spmm_narrow_f32 streams instead of cols / indptr (see include/pglb.h).  `hints`: bit 31 of the plan
    carries an L2 residency hint.  `slot_scale(scale_src)`: the source scale laid out per CSR slot (cached per scale
    tensor by the EdgeIndex).  `fallback()` gives the packed column ids for calls the narrow kernel does not take."""
    __slots__ = ("plan", "nz_row", "blk_k", "hints", "slot_scale", "fallback")

    def __init__(self, plan, nz_row, blk_k, hints=False, slot_scale=None, fallback=None):
        self.plan, self.nz_row, self.blk_k, self.hints = plan, nz_row, blk_k, bool(hints)
        self.slot_scale, self.fallback = slot_scale, fallback


def narrow_plan(indptr, cols, n_src, row_bytes=64):
    """Build the plan of a dst-CSR once (cached by the EdgeIndex).  Returns (plan, nz_row, blk_k, hints)."""
    require_cuda(indptr, cols)
    dev = indptr.device
    n_dst = int(indptr.shape[0]) - 1
    E = int(cols.shape[0])
    pk = pack_cols(cols, n_src, row_bytes, NARROW_HOT_BYTES) if NARROW_HOT_BYTES > 0 else None
    hints = bool(pk is not None and pk[1])
    plan = torch.empty(E, dtype=torch.int32, device=dev)
    nz_row = torch.empty(n_dst + 2, dtype=torch.int32, device=dev)
    blk_k = torch.empty((E + 31) // 32, dtype=torch.int32, device=dev)
    need = ctypes.c_size_t(0)
    check(lib.pglb_narrow_plan_ws(n_dst, ctypes.byref(need)))
    ws = workspace(dev, need.value)
    with torch.cuda.device(dev):
        check(lib.pglb_narrow_plan(_ptr(indptr), _ptr(cols), _ptr(pk[0]) if hints else None, n_dst, int(n_src), E,
                                   _ptr(plan), _ptr(nz_row), _ptr(blk_k), _ptr(ws), ws.numel(), _stream()))
    return plan, nz_row, blk_k, hints


def _packed_of(csr, x2):
    """Per-(graph, row width) index data from the EdgeIndex cache: packed column ids (+ optional L2 hints) for the
    wide-row kernels (64 < D <= 128), the narrow plan for D <= 64."""
    fn = csr.get("packed")
    D = int(x2.shape[1])
    if fn is None or D > 128 or D % 4 or (D <= 64 and not (NARROW_ROWS or NARROW2)):
        return None
    n_src = int(x2.shape[0])
    if D <= 64 and NARROW2 and csr.get("plan") is not None and 0 < n_src < (1 << 30) and \
            csr["cols"] is not None and int(csr["cols"].shape[0]) > 0:
        plan, nz_row, blk_k, hints = csr["plan"](n_src, D * 4)
        return NarrowPlan(plan, nz_row, blk_k, hints, slot_scale=csr.get("slot_scale"),
                          fallback=(lambda: fn(n_src, D * 4)) if NARROW_ROWS else None)
    if D <= 64 and not NARROW_ROWS:
        return None
    return fn(n_src, D * 4)


@f64_through
def aggregate_copy(x, fwd, n_dst, reduce_op="sum", bwd=None, scale_src=None, scale_dst=None):
    """send_u_recv on a cached dst-CSR.  fwd/bwd: dicts {indptr, cols, degree, max_degree}."""
    require_cuda(x)
    shape = x.shape
    x2 = _f32_2d(x)
    needs_grad = x2.requires_grad and torch.is_grad_enabled()
    if reduce_op in ("sum", "mean") and needs_grad:
        out = _CopyAgg.apply(x2, fwd, bwd, n_dst, reduce_op, scale_src, scale_dst)
    elif needs_grad and scale_src is None and scale_dst is None:
        out = _MaxMinAgg.apply(x2, fwd, bwd, n_dst, reduce_op)
    elif needs_grad:
        # max / min with fused scales has no backward kernel: do not hand back a silently detached result
        raise NotImplementedError("pgl_b200: send_u_recv(%s) with fused scale vectors is not differentiable; apply the "
                                  "scales outside the aggregation" % reduce_op)
    else:
        out = _spmm_raw(fwd["indptr"], fwd["cols"], x2, n_dst, reduce_op, scale_src=scale_src,
                        scale_dst=scale_dst, max_degree=fwd.get("max_degree", -1),
                        packed=_packed_of(fwd, x2))
    return out.reshape((n_dst,) + tuple(shape[1:]))


def classify_bcast(x_shape, y_shape):
    """Map (x feature dims, y feature dims) to a kernel broadcast mode.
    Returns (mode, head_dim) or None when y must be materialised by expand()."""
    xf = tuple(x_shape[1:])
    yf = tuple(y_shape[1:])
    D = 1
    for s in xf:
        D *= s
    ny = 1
    for s in yf:
        ny *= s
    if ny == 1:
        return BCAST_SCALAR, 1
    if len(yf) < len(xf):
        yf = (1,) * (len(xf) - len(yf)) + yf
    if len(yf) != len(xf):
        return None
    if yf == xf:
        return BCAST_FULL, 1
    # [.., H, 1, 1] against [.., H, a, b]: leading dims equal, trailing dims of y all 1
    k = len(xf)
    while k > 0 and yf[k - 1] == 1:
        k -= 1
    if yf[:k] == xf[:k]:
        hd = 1
        for s in xf[k:]:
            hd *= s
        return BCAST_HEAD, hd
    return None


def sddmm_dot(a2, b2, ia, ib, H, Dh):
    """out[e,h] = <a2[ia[e], h, :], b2[ib[e], h, :]>; a2, b2 are [n, H*Dh]."""
    E = int(ia.shape[0])
    out = torch.empty((E, H), dtype=torch.float32, device=a2.device)
    with torch.cuda.device(a2.device):
        check(lib.pglb_sddmm_dot_f32(_ptr(a2), _ptr(b2), _ptr(ia), max(ia.stride(0), 1) if E else 1,
                                     _ptr(ib), max(ib.stride(0), 1) if E else 1, E, H, Dh, _ptr(out),
                                     _stream()))
    return out


class _UeAgg(torch.autograd.Function):
    """send_ue_recv with reduce sum and message mul / add, differentiable in x and y."""

    @staticmethod
    def forward(ctx, x2, y2, fwd, bwd, edges, n_dst, mode, hd, msg_op):
        ctx.meta = (fwd, bwd, edges, mode, hd, msg_op)
        ctx.save_for_backward(x2, y2)
        return _spmm_raw(fwd["indptr"], fwd["cols"], x2, n_dst, "sum", eid=fwd["eid"], y2=y2,
                         y_bcast=mode, head_dim=hd, msg_op=msg_op, max_degree=fwd.get("max_degree", -1))

    @staticmethod
    def backward(ctx, g):
        fwd, bwd, edges, mode, hd, msg_op = ctx.meta
        bwd = bwd() if callable(bwd) else bwd
        x2, y2 = ctx.saved_tensors
        g = g.contiguous()
        n_src, D = int(x2.shape[0]), int(x2.shape[1])
        src, dst = edges[:, 0], edges[:, 1]
        gx = gy = None
        if msg_op == "mul":
            if ctx.needs_input_grad[0]:
                gx = _spmm_raw(bwd["indptr"], bwd["cols"], g, n_src, "sum", eid=bwd["eid"], y2=y2,
                               y_bcast=mode, head_dim=hd, msg_op="mul",
                               max_degree=bwd.get("max_degree", -1))
            if ctx.needs_input_grad[1]:
                if mode == BCAST_FULL:
                    gy = send_uv(x2, g, src, dst, "mul")
                elif mode == BCAST_HEAD:
                    gy = sddmm_dot(x2, g, src, dst, D // hd, hd)
                else:
                    gy = sddmm_dot(x2, g, src, dst, 1, D)
        else:  # add
            if ctx.needs_input_grad[0]:
                gx = _spmm_raw(bwd["indptr"], bwd["cols"], g, n_src, "sum",
                               max_degree=bwd.get("max_degree", -1))
            if ctx.needs_input_grad[1]:
                ge = gather_rows(g, dst)
                if mode == BCAST_FULL:
                    gy = ge
                elif mode == BCAST_HEAD:
                    gy = ge.reshape(ge.shape[0], D // hd, hd).sum(-1)
                else:
                    gy = ge.sum(-1, keepdim=True)
        return gx, gy, None, None, None, None, None, None, None


def _ue_with_grad(x2, y2, fwd, bwd, edges, n_dst, mode, hd, message_op, reduce_op):
    """Differentiable send_ue_recv for every (message_op, reduce_op) of the reference (pgl/graph.py:889-937).
    add / mul with sum run on the fused kernels (_UeAgg: transpose-SpMM for the node operand, SDDMM / gather for
    the edge operand).  The rest is composed from those and from pieces that already carry a backward:
      sub  = add with -y,   div = mul with 1 / y   (autograd differentiates the negation / reciprocal),
      mean = sum / in-degree,
      max / min: the [E, D] message is materialised like the reference does (gather_rows(x, src) op y) and reduced
      by the segment kernel over the dst-CSR slots, whose backward routes the gradient to the selected entries."""
    if message_op == "sub":
        return _ue_with_grad(x2, -y2, fwd, bwd, edges, n_dst, mode, hd, "add", reduce_op)
    if message_op == "div":
        return _ue_with_grad(x2, 1.0 / y2, fwd, bwd, edges, n_dst, mode, hd, "mul", reduce_op)
    if reduce_op == "sum":
        return _UeAgg.apply(x2, y2, fwd, bwd, edges, n_dst, mode, hd, message_op)
    if reduce_op == "mean":
        total = _UeAgg.apply(x2, y2, fwd, bwd, edges, n_dst, mode, hd, message_op)
        deg = (fwd["indptr"][1:] - fwd["indptr"][:-1])[:n_dst].clamp(min=1).to(total.dtype)
        return total / deg.unsqueeze(1)
    # max / min
    D = int(x2.shape[1])
    xe = gather_rows(x2, edges[:, 0])                                     # [E, D], differentiable
    if mode == BCAST_FULL:
        ye = y2
    elif mode == BCAST_HEAD:
        ye = y2.reshape(y2.shape[0], D // hd, 1).expand(-1, -1, hd).reshape(y2.shape[0], D)
    else:
        ye = y2.reshape(-1, 1).expand(-1, D)
    msg = xe * ye if message_op == "mul" else xe + ye
    return segment_reduce(msg, None, reduce_op, indptr=fwd["indptr"][: n_dst + 1], cols=fwd["eid"],
                          max_degree=fwd.get("max_degree", -1))


@f64_through
def aggregate_ue(x, y, fwd, n_dst, message_op="add", reduce_op="sum", bwd=None, edges=None):
    """send_ue_recv on a cached dst-CSR: y is in original edge order, read through eid."""
    require_cuda(x, y)
    xs = tuple(x.shape)
    y_shape_in = tuple(y.shape)
    if y.dim() == 1:
        y = y.reshape(-1, 1)
    out_feat = torch.broadcast_shapes(tuple(x.shape[1:]) if x.dim() > 1 else (1,), tuple(y.shape[1:]))
    if tuple(out_feat) != tuple(xs[1:]):
        x = x.reshape((xs[0],) + (1,) * (len(out_feat) - len(xs[1:])) + tuple(xs[1:]))
        x = x.expand((xs[0],) + tuple(out_feat)).contiguous()
    cls = classify_bcast(x.shape, y.shape)
    if cls is None:
        y = y.expand((y.shape[0],) + tuple(out_feat)).contiguous()
        cls = (BCAST_FULL, 1)
    mode, hd = cls
    x2 = _f32_2d(x)
    y2 = _f32_2d(y)
    needs_grad = torch.is_grad_enabled() and (x2.requires_grad or y2.requires_grad)
    if needs_grad:
        if bwd is None or edges is None:
            raise NotImplementedError("pgl_b200: a differentiable send_ue_recv needs the reverse CSR and the edge list")
        out = _ue_with_grad(x2, y2, fwd, bwd, edges, n_dst, mode, hd, message_op, reduce_op)
    else:
        out = _spmm_raw(fwd["indptr"], fwd["cols"], x2, n_dst, reduce_op, eid=fwd["eid"], y2=y2,
                        y_bcast=mode, head_dim=hd, msg_op=message_op,
                        max_degree=fwd.get("max_degree", -1))
    return out.reshape((n_dst,) + tuple(out_feat))


class _SegmentReduce(torch.autograd.Function):
    """Row reduce over a compact CSR (segment_* / fused gather+segment).  Backward: sum / mean
    expand the gradient back over the slots (then scatter-add through `cols` if the forward
    gathered); max / min route it to the entries equal to the reduced value."""

    @staticmethod
    def forward(ctx, d2, indptr, cols, pool_type, max_degree, n_slots):
        K = int(indptr.shape[0]) - 1
        out = _spmm_raw(indptr, cols, d2, K, pool_type, num_edges=n_slots, max_degree=max_degree)
        ctx.meta = (indptr, cols, pool_type, n_slots)
        ctx.save_for_backward(d2, out)
        return out

    @staticmethod
    def backward(ctx, g):
        indptr, cols, pool_type, n_slots = ctx.meta
        d2, out = ctx.saved_tensors
        g = g.contiguous()
        K = int(indptr.shape[0]) - 1
        counts = indptr[1:] - indptr[:-1]
        seg = torch.repeat_interleave(torch.arange(K, device=g.device), counts)  # slot -> row
        ge = gather_rows(g, seg)  # [n_slots, D]
        if pool_type == "mean":
            ge = ge / counts.clamp(min=1).to(ge.dtype)[seg].unsqueeze(1)
        elif pool_type in ("max", "min"):
            src_rows = d2 if cols is None else gather_rows(d2, cols)
            ge = ge * (src_rows == gather_rows(out, seg)).to(ge.dtype)
        if cols is None:
            return ge, None, None, None, None, None
        gd = torch.zeros_like(d2).index_add_(0, cols, ge)
        return gd, None, None, None, None, None


@f64_through
def segment_reduce(data, segment_ids, pool_type, num_segments=None, indptr=None, cols=None,
                   max_degree=-1):
    """paddle.geometric.segment_* over sorted ids (reference pgl/math.py:36-42).
    `cols` (optional) fuses a row gather: slot j reads data[cols[j]] (lazy messages)."""
    require_cuda(data)
    d2 = _f32_2d(data)
    if indptr is None:
        ids = _i64(segment_ids)
        E = int(ids.shape[0])
        if num_segments is None:
            num_segments = int(ids[-1].item()) + 1 if E > 0 else 0
        indptr = segment_indptr(ids, num_segments)
    else:
        E = int(cols.shape[0]) if cols is not None else int(d2.shape[0])
        num_segments = int(indptr.shape[0]) - 1
    if d2.requires_grad and torch.is_grad_enabled():
        out = _SegmentReduce.apply(d2, indptr, cols, pool_type, max_degree, E)
    else:
        out = _spmm_raw(indptr, cols, d2, int(num_segments), pool_type, num_edges=E,
                        max_degree=max_degree)
    feat = tuple(data.shape[1:])
    return out.reshape((int(num_segments),) + feat)


def _send_uv_raw(x2, y2, src, dst, message_op):
    E = int(src.shape[0])
    D = int(x2.shape[1])
    out = torch.empty((E, D), dtype=torch.float32, device=x2.device)
    with torch.cuda.device(x2.device):
        check(lib.pglb_send_uv_f32(_ptr(x2), _ptr(y2), _ptr(src), max(src.stride(0), 1) if E else 1,
                                   _ptr(dst), max(dst.stride(0), 1) if E else 1, E, D,
                                   MSG[message_op], _ptr(out), _stream()))
    return out


class _SendUV(torch.autograd.Function):
    """send_uv (add / sub / mul) with gradients reduced over the cached src- and dst-CSR."""

    @staticmethod
    def forward(ctx, x2, y2, src, dst, message_op, src_csr, dst_csr):
        ctx.meta = (message_op, src_csr, dst_csr)
        ctx.save_for_backward(x2, y2)
        return _send_uv_raw(x2, y2, src, dst, message_op)

    @staticmethod
    def backward(ctx, g):
        op_, src_csr, dst_csr = ctx.meta
        x2, y2 = ctx.saved_tensors
        g = g.contiguous()
        gx = gy = None
        sc = src_csr() if callable(src_csr) else src_csr
        dc = dst_csr() if callable(dst_csr) else dst_csr
        if op_ in ("add", "sub"):
            if ctx.needs_input_grad[0]:   # sum of g over the out-edges of every source
                gx = _spmm_raw(sc["indptr"], sc["eid"], g, int(x2.shape[0]), "sum",
                               max_degree=sc.get("max_degree", -1))
            if ctx.needs_input_grad[1]:
                gy = _spmm_raw(dc["indptr"], dc["eid"], g, int(y2.shape[0]), "sum",
                               max_degree=dc.get("max_degree", -1))
                if op_ == "sub":
                    gy = -gy
        elif op_ == "mul":
            if ctx.needs_input_grad[0]:   # gx[s] = sum_e g[e] * y[dst[e]]
                gx = _spmm_raw(sc["indptr"], sc["cols"], y2, int(x2.shape[0]), "sum", eid=sc["eid"],
                               y2=g, y_bcast=BCAST_FULL, msg_op="mul", max_degree=sc.get("max_degree", -1))
            if ctx.needs_input_grad[1]:
                gy = _spmm_raw(dc["indptr"], dc["cols"], x2, int(y2.shape[0]), "sum", eid=dc["eid"],
                               y2=g, y_bcast=BCAST_FULL, msg_op="mul", max_degree=dc.get("max_degree", -1))
        else:
            raise NotImplementedError("pgl_b200: send_uv(div) has no backward yet")
        return gx, gy, None, None, None, None, None


@f64_through
def send_uv(x, y, src, dst, message_op="add", src_csr=None, dst_csr=None):
    require_cuda(x, y, src, dst)
    xf = tuple(x.shape[1:]) if x.dim() > 1 else (1,)
    yf = tuple(y.shape[1:]) if y.dim() > 1 else (1,)
    of = tuple(torch.broadcast_shapes(xf, yf))
    if xf != of:
        x = x.reshape((x.shape[0],) + (1,) * (len(of) - len(xf)) + xf).expand((x.shape[0],) + of)
    if yf != of:
        y = y.reshape((y.shape[0],) + (1,) * (len(of) - len(yf)) + yf).expand((y.shape[0],) + of)
    x2 = _f32_2d(x.reshape(x.shape[0], -1) if x.dim() != 2 else x)
    y2 = _f32_2d(y.reshape(y.shape[0], -1) if y.dim() != 2 else y)
    E = int(src.shape[0])
    if torch.is_grad_enabled() and (x2.requires_grad or y2.requires_grad) and src_csr is not None:
        out = _SendUV.apply(x2, y2, src, dst, message_op, src_csr, dst_csr)
    else:
        out = _send_uv_raw(x2, y2, src, dst, message_op)
    return out.reshape((E,) + of)


def _gather_rows_raw(x2, index, out2=None):
    n = int(index.shape[0])
    D = int(x2.shape[1])
    dev = index.device
    if out2 is None:
        out2 = torch.empty((n, D), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.pglb_gather_rows_f32(_ptr(x2), x2.stride(0), _ptr(index),
                                       max(index.stride(0), 1) if n else 1, n, D, _ptr(out2),
                                       out2.stride(0), _stream()))
    return out2


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, index):
        ctx.meta = (index, int(x2.shape[0]))
        return _gather_rows_raw(x2, index)

    @staticmethod
    def backward(ctx, g):
        index, n = ctx.meta
        gx = torch.zeros((n, g.shape[1]), dtype=g.dtype, device=g.device)
        gx.index_add_(0, index.contiguous(), g.contiguous())
        return gx, None


@f64_through
def gather_rows(x, index, out=None):
    """paddle.gather(x, index, axis=0).  The kernel is launched on the device of ``index`` /
    ``out``; ``x`` may live on a peer GPU mapped into this process (NVLink P2P pull)."""
    require_cuda(x, index)
    index = _i64(index)
    n = int(index.shape[0])
    if x.dtype != torch.float32:
        return x.index_select(0, index.contiguous())  # integer gathers (degree subsets): torch device op
    x2 = _f32_2d(x)
    D = int(x2.shape[1])
    if out is None and x2.requires_grad and torch.is_grad_enabled():
        return _GatherRows.apply(x2, index).reshape((n,) + tuple(x.shape[1:]))
    out2 = None
    if out is not None:
        out2 = out.reshape(n, D) if out.dim() != 2 else out
    res = _gather_rows_raw(x2, index, out2)
    if out is not None:
        return out
    return res.reshape((n,) + tuple(x.shape[1:]))


class _ScatterRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, init, index, updates):
        out = init.clone()
        o2 = out.reshape(out.shape[0], -1) if out.dim() != 2 else out
        u2 = _f32_2d(updates)
        n = int(index.shape[0])
        with torch.cuda.device(out.device):
            check(lib.pglb_scatter_rows_f32(_ptr(u2), u2.stride(0), _ptr(index), n, int(u2.shape[1]),
                                            _ptr(o2), o2.stride(0), _stream()))
        ctx.index = index
        return out

    @staticmethod
    def backward(ctx, g):
        index = ctx.index
        g = g.contiguous()
        gu = _gather_rows_raw(g.reshape(g.shape[0], -1), index).reshape((index.shape[0],) + tuple(g.shape[1:]))
        gi = g.clone()
        gi.index_fill_(0, index, 0)
        return gi, None, gu


@f64_through
def scatter_rows(init, index, updates):
    """paddle.scatter(init, index, updates, overwrite=True) for unique indices (out-of-place)."""
    require_cuda(init, index, updates)
    index = _i64(index).contiguous()
    return _ScatterRows.apply(init, index, updates)


def _edge_softmax_raw(indptr, eid, l2, E):
    H = int(l2.shape[1])
    out = torch.empty_like(l2)
    need = ctypes.c_size_t(0)
    check(lib.pglb_edge_softmax_csr_ws(E, ctypes.byref(need)))
    ws = workspace(l2.device, need.value)
    with torch.cuda.device(l2.device):
        check(lib.pglb_edge_softmax_csr_f32(_ptr(indptr), _ptr(eid), _ptr(l2), _ptr(out),
                                            int(indptr.shape[0]) - 1, E, H, _ptr(ws), ws.numel(),
                                            _stream()))
    return out


class _EdgeSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, l2, indptr, eid, E):
        out = _edge_softmax_raw(indptr, eid, l2, E)
        ctx.meta = (indptr, eid, E)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        indptr, eid, E = ctx.meta
        (alpha,) = ctx.saved_tensors
        g = g.contiguous()
        gl = torch.zeros_like(alpha)
        with torch.cuda.device(alpha.device):
            check(lib.pglb_edge_softmax_bwd_csr_f32(_ptr(indptr), _ptr(eid), _ptr(alpha), _ptr(g),
                                                    _ptr(gl), int(indptr.shape[0]) - 1, E,
                                                    int(alpha.shape[1]), _stream()))
        return gl, None, None, None


@f64_through
def edge_softmax_csr(indptr, eid, logits, num_edges):
    """Fused per-row softmax; eid=None -> rows are contiguous slots (segment_softmax)."""
    require_cuda(indptr, logits)
    shape = logits.shape
    l2 = logits.reshape(shape[0], -1) if logits.dim() != 2 else logits
    l2 = _f32_2d(l2)
    E = int(num_edges)
    if l2.requires_grad and torch.is_grad_enabled():
        out = _EdgeSoftmax.apply(l2, indptr, eid, E)
    else:
        out = _edge_softmax_raw(indptr, eid, l2, E)
    return out.reshape(shape)


@f64_through
def gat_attention_csr(csr, attn_src, attn_dst, negative_slope=0.2):
    """alpha[slot, h] = softmax over the dst row of leaky_relu(attn_src[src] + attn_dst[dst]) in CSR
    SLOT order (send_uv + LeakyReLU + edge_softmax fused; inference only)."""
    require_cuda(attn_src, attn_dst)
    a_s = _f32_2d(attn_src)
    a_d = _f32_2d(attn_dst)
    H = int(a_s.shape[1])
    E = int(csr["cols"].shape[0])
    out = torch.empty((E, H), dtype=torch.float32, device=a_s.device)
    need = ctypes.c_size_t(0)
    check(lib.pglb_edge_softmax_csr_ws(E, ctypes.byref(need)))
    ws = workspace(a_s.device, need.value)
    with torch.cuda.device(a_s.device):
        check(lib.pglb_gat_attention_csr_f32(_ptr(csr["indptr"]), _ptr(csr["cols"]), _ptr(a_s), _ptr(a_d),
                                             float(negative_slope), _ptr(out),
                                             int(csr["indptr"].shape[0]) - 1, E, H, _ptr(ws), ws.numel(),
                                             _stream()))
    return out


@f64_through
def gat_fused(csr, f, attn_src, attn_dst, negative_slope=0.2):
    """Single-pass GAT aggregation (online softmax in the wide-row kernel).  f [N, H, Dh] with
    H*Dh <= 128 and Dh % 4 == 0; returns [N, H, Dh], or None when the shape is not supported."""
    require_cuda(f, attn_src, attn_dst)
    n, H, Dh = int(f.shape[0]), int(f.shape[1]), int(f.shape[2])
    if H * Dh > 128 or Dh % 4 or H * Dh <= 64:
        return None
    f2 = _f32_2d(f)
    a_s, a_d = _f32_2d(attn_src), _f32_2d(attn_dst)
    E = int(csr["cols"].shape[0])
    n_dst = int(csr["indptr"].shape[0]) - 1
    out = torch.empty((n_dst, H * Dh), dtype=torch.float32, device=f2.device)
    need = ctypes.c_size_t(0)
    check(lib.pglb_spmm_csr_ws(n_dst, E, H * Dh, ctypes.byref(need)))
    ws = workspace(f2.device, need.value)
    with torch.cuda.device(f2.device):
        check(lib.pglb_gat_fused_csr_f32(_ptr(csr["indptr"]), _ptr(csr["cols"]), _ptr(f2), f2.stride(0),
                                         _ptr(a_s), _ptr(a_d), float(negative_slope), _ptr(out),
                                         out.stride(0), n_dst, n, E, H, Dh, _ptr(ws), ws.numel(),
                                         _stream()))
    return out.reshape(n_dst, H, Dh)


class _HeadDots(torch.autograd.Function):
    """attn_src / attn_dst of GATConv in one pass over the features (pglb_head_dots_f32); the backward is three
    elementwise / reduce torch ops on the gradient path."""

    @staticmethod
    def forward(ctx, f3, w_src, w_dst):
        n, H, Dh = (int(v) for v in f3.shape)
        f2 = f3.reshape(n, H * Dh)
        a_s = torch.empty((n, H), dtype=torch.float32, device=f3.device)
        a_d = torch.empty((n, H), dtype=torch.float32, device=f3.device)
        ws, wd = w_src.contiguous(), w_dst.contiguous()
        with torch.cuda.device(f3.device):
            check(lib.pglb_head_dots_f32(_ptr(f2), f2.stride(0), n, H, Dh, _ptr(ws), _ptr(wd), _ptr(a_s), _ptr(a_d),
                                         _stream()))
        ctx.save_for_backward(f3, w_src, w_dst)
        return a_s, a_d

    @staticmethod
    def backward(ctx, g_s, g_d):
        f3, w_src, w_dst = ctx.saved_tensors
        gf = gw_s = gw_d = None
        if ctx.needs_input_grad[0]:
            gf = g_s.unsqueeze(-1) * w_src + g_d.unsqueeze(-1) * w_dst
        if ctx.needs_input_grad[1]:
            gw_s = (g_s.unsqueeze(-1) * f3).sum(0).reshape(w_src.shape)
        if ctx.needs_input_grad[2]:
            gw_d = (g_d.unsqueeze(-1) * f3).sum(0).reshape(w_dst.shape)
        return gf, gw_s, gw_d


def head_dots(f, w_src, w_dst):
    """(sum(f * w_src, -1), sum(f * w_dst, -1)) for f [N, H, Dh], w [H, Dh] (or [1, H, Dh]) -- GATConv's attention
    projections, reference pgl/nn/conv.py:323-326 -- or None when the shape is outside the kernel (the caller keeps
    the two torch expressions)."""
    if f.dim() != 3 or not f.is_cuda or f.dtype != torch.float32 or w_src.dtype != torch.float32:
        return None
    n, H, Dh = (int(v) for v in f.shape)
    lph = Dh // 4
    if H * Dh > 128 or Dh % 4 or (lph & (lph - 1)) or w_src.numel() != H * Dh or w_dst.numel() != H * Dh:
        return None
    if not f.is_contiguous():
        f = f.contiguous()
    if f.data_ptr() % 16 or w_src.data_ptr() % 16 or w_dst.data_ptr() % 16:
        return None
    return _HeadDots.apply(f, w_src, w_dst)


class _GatFused(torch.autograd.Function):
    """The single-pass GAT aggregation under autograd (reference pgl/nn/conv.py:333-339 differentiated by Paddle over
    four ops).  Forward = pglb_gat_fused_train_csr_f32 (one launch; keeps only lse[N, H]).  Backward = one edge kernel
    (pglb_gat_bwd_edge_f32: alpha rebuilt, SDDMM dot, softmax and LeakyReLU backward) + three reverse-CSR aggregations
    on the existing kernels."""

    @staticmethod
    def forward(ctx, f2, a_s, a_d, fwd, bwd, slope, H, Dh):
        n = int(f2.shape[0])
        E = int(fwd["cols"].shape[0])
        n_dst = int(fwd["indptr"].shape[0]) - 1
        out = torch.empty((n_dst, H * Dh), dtype=torch.float32, device=f2.device)
        lse = torch.zeros((n_dst, H), dtype=torch.float32, device=f2.device)
        need = ctypes.c_size_t(0)
        check(lib.pglb_spmm_csr_ws(n_dst, E, H * Dh, ctypes.byref(need)))
        ws = workspace(f2.device, need.value)
        with torch.cuda.device(f2.device):
            check(lib.pglb_gat_fused_train_csr_f32(_ptr(fwd["indptr"]), _ptr(fwd["cols"]), _ptr(f2), f2.stride(0),
                                                   _ptr(a_s), _ptr(a_d), float(slope), _ptr(out), out.stride(0),
                                                   _ptr(lse), n_dst, n, E, H, Dh, _ptr(ws), ws.numel(), _stream()))
        ctx.meta = (fwd, bwd, float(slope), H, Dh)
        ctx.save_for_backward(f2, a_s, a_d, out, lse)
        return out

    @staticmethod
    def backward(ctx, g):
        fwd, bwd, slope, H, Dh = ctx.meta
        bwd = bwd() if callable(bwd) else bwd
        f2, a_s, a_d, out, lse = ctx.saved_tensors
        g = _f32_2d(g)
        E = int(fwd["cols"].shape[0])
        n_src, n_dst = int(f2.shape[0]), int(out.shape[0])
        alpha = torch.empty((E, H), dtype=torch.float32, device=g.device)
        dz = torch.empty((E, H), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(lib.pglb_gat_bwd_edge_f32(_ptr(fwd["rows"]), _ptr(fwd["cols"]), _ptr(fwd["eid"]), _ptr(f2),
                                            f2.stride(0), _ptr(g), g.stride(0), _ptr(out), out.stride(0), _ptr(a_s),
                                            _ptr(a_d), _ptr(lse), slope, E, H, Dh, _ptr(alpha), _ptr(dz), _stream()))
        gf = gs = gd = None
        if ctx.needs_input_grad[0]:   # grad f[s] = sum over the out-edges of s of alpha[e] * g[dst[e]]
            gf = _spmm_raw(bwd["indptr"], bwd["cols"], g, n_src, "sum", eid=bwd["eid"], y2=alpha, y_bcast=BCAST_HEAD,
                           head_dim=Dh, msg_op="mul", max_degree=bwd.get("max_degree", -1))
        if ctx.needs_input_grad[1]:   # z = as[src] + ad[dst]: send_uv(add)'s backward
            gs = _spmm_raw(bwd["indptr"], bwd["eid"], dz, n_src, "sum", max_degree=bwd.get("max_degree", -1))
        if ctx.needs_input_grad[2]:
            gd = _spmm_raw(fwd["indptr"], fwd["eid"], dz, n_dst, "sum", max_degree=fwd.get("max_degree", -1))
        return gf, gs, gd, None, None, None, None, None


GAT_FUSED_TRAIN = os.environ.get("PGLB_GAT_FUSED_TRAIN", "1") != "0"


def gat_fused_train(fwd, bwd, f, attn_src, attn_dst, negative_slope=0.2):
    """gat_fused with a backward (float32 only).  f [N, H, Dh], 64 < H*Dh <= 128, H % 4 == 0, Dh a power of two >= 4,
    0 <= slope <= 1; returns [N, H, Dh], or None when the shape is outside the fused kernels (the caller keeps the
    op-by-op path).  ``bwd`` is the reverse CSR dict or a callable returning it."""
    require_cuda(f, attn_src, attn_dst)
    n, H, Dh = int(f.shape[0]), int(f.shape[1]), int(f.shape[2])
    if not GAT_FUSED_TRAIN or f.dtype != torch.float32 or H * Dh > 128 or H * Dh <= 64 or H % 4 or H > 32:
        return None
    if Dh < 4 or (Dh & (Dh - 1)) or not (0.0 <= float(negative_slope) <= 1.0):
        return None
    n_dst = int(fwd["indptr"].shape[0]) - 1
    if n_dst != n or int(fwd["cols"].shape[0]) == 0 or fwd.get("rows") is None:
        return None
    f2 = _f32_2d(f)
    a_s, a_d = _f32_2d(attn_src), _f32_2d(attn_dst)
    if f2.data_ptr() % 16 or (f2.stride(0) * 4) % 16 or a_s.data_ptr() % 16:
        return None
    try:
        out = _GatFused.apply(f2, a_s, a_d, fwd, bwd, float(negative_slope), H, Dh)
    except _lib.PglbError as ex:
        if ex.code == -4:   # PGLB_EUNSUPPORTED: e.g. 32 heads, whose attention rows do not fit the kernel's rings
            return None
        raise
    return out.reshape(n_dst, H, Dh)


@f64_through
def aggregate_ue_slots(x, y_slots, fwd, n_dst, message_op="mul", reduce_op="sum"):
    """send_ue_recv whose edge operand is already in CSR slot order (read sequentially)."""
    require_cuda(x, y_slots)
    out_feat = tuple(x.shape[1:])
    cls = classify_bcast(x.shape, y_slots.shape)
    if cls is None:
        raise ValueError("pgl_b200: unsupported broadcast for slot-ordered edge operand")
    mode, hd = cls
    out = _spmm_raw(fwd["indptr"], fwd["cols"], _f32_2d(x), n_dst, reduce_op, eid=None,
                    y2=_f32_2d(y_slots), y_bcast=mode, head_dim=hd, msg_op=message_op,
                    max_degree=fwd.get("max_degree", -1))
    return out.reshape((n_dst,) + out_feat)


TC_GEMM = os.environ.get("PGLB_TC_GEMM", "1") != "0"
TC_GEMM_MIN_ROWS = int(os.environ.get("PGLB_TC_GEMM_MIN_ROWS", "4096"))


def linear_tc_ok(x, weight):
    """True when x @ weight can run on pglb_linear_tf32x3_f32 (tall-skinny fp32, K <= 128 and
    N in {64, 128}); small or odd shapes stay on torch.matmul (cuBLAS), which is plumbing here."""
    if not TC_GEMM or not x.is_cuda or x.dim() != 2 or x.dtype != torch.float32 \
            or weight.dtype != torch.float32:
        return False
    M, K = int(x.shape[0]), int(x.shape[1])
    N = int(weight.shape[1])
    return M >= TC_GEMM_MIN_ROWS and K % 4 == 0 and 4 <= K <= 128 and N in (64, 128) \
        and int(weight.shape[0]) == K


def _linear_tc_raw(x, weight, bias, act):
    x2 = x if (x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0) \
        else x.contiguous()
    w = weight.contiguous()
    M, K = int(x2.shape[0]), int(x2.shape[1])
    N = int(w.shape[1])
    out = torch.empty((M, N), dtype=torch.float32, device=x2.device)
    b = bias.contiguous() if bias is not None else None
    with torch.cuda.device(x2.device):
        check(lib.pglb_linear_tf32x3_f32(_ptr(x2), x2.stride(0), _ptr(w), _ptr(b) if b is not None else None,
                                         _ptr(out), N, M, K, N, 1 if act == "relu" else 0, _stream()))
    return out


class _LinearTC(torch.autograd.Function):
    """Forward on the 3xTF32 tensor-core kernel (bias + ReLU fused); backward is two plain fp32
    matmuls (training plumbing: the reference gets them from Paddle autograd)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        out = _linear_tc_raw(x, weight, bias, act)
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, out if act == "relu" else None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight, out = ctx.saved_tensors
        if ctx.act == "relu":
            g = g * (out > 0)
        gx = g @ weight.t() if ctx.needs_input_grad[0] else None
        gw = x.t() @ g if ctx.needs_input_grad[1] else None
        gb = g.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None


@f64_through
def linear_tc(x, weight, bias=None, act=None):
    """act(x @ weight + bias) with weight [in, out]; act in (None, "relu")."""
    require_cuda(x, weight)
    assert act in (None, "relu")
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad
                                    or (bias is not None and bias.requires_grad)):
        return _LinearTC.apply(x, weight, bias, act)
    return _linear_tc_raw(x, weight, bias, act)


def degree_norm(degree):
    require_cuda(degree)
    degree = _i64(degree).contiguous()
    n = int(degree.shape[0])
    out = torch.empty(n, dtype=torch.float32, device=degree.device)
    with torch.cuda.device(degree.device):
        check(lib.pglb_degree_norm_f32(_ptr(degree), n, _ptr(out), _stream()))
    return out.reshape(-1, 1)


def launch_count():
    return _lib.launch_count()


def stream_kernel_name(num_edges=None, num_rows=None):
    """Which wide-row streaming kernel pglb_spmm_csr_f32 routes copy-sum aggregations to (PGLB_STREAM_V5, PGLB_V5_GEO,
    PGLB_V5_DYN; mirrors csrc/spmm_v5.inl v5_mode() / v5_geo(): without PGLB_V5_GEO, graphs of >= 12 slots per row
    take the groups-of-8 geometry)."""
    v = os.environ.get("PGLB_STREAM_V5", "1")
    g = os.environ.get("PGLB_V5_GEO")
    if g not in ("0", "1", "2"):
        g = "2" if (num_edges is not None and num_rows and num_edges >= 12 * num_rows) else "0"
    geo = {"0": "GRP=4,NG=4,W=13", "1": "GRP=8,NG=4,W=6", "2": "GRP=8,NG=3,W=9"}[g]
    if v == "0":
        return "spmm_stream128_kernel<RK=0,SCALED=1,PK=2,YM=0,CFG=1> (+ task_plan, empty_rows, fix-up kernels)"
    how = "TMA tile::gather4 (UTMALDG.2D.GATHER4)" if v != "2" else "LDGSTS ring"
    q = "" if os.environ.get("PGLB_V5_DYN") == "0" or v == "2" else ", device-side task queue"
    return "spmm_v5_kernel<%s, %s%s> (+ task_plan, empty_rows, fix-up kernels)" % (how, geo, q)
