"""Rr x Rc shard grid for full-batch aggregation on R = Rr * Rc GPUs: rank (r, c) owns destination-row block r and
feature-column block c of EVERY source row (column blocks are replicated across the Rr row groups, the way the
reference's DistGPUGraph replicates whole features, pgl/graph.py:1410-1553).  The aggregation itself needs no exchange
on any grid; what a grid costs is the replication.  ``GridHostAggregator`` is the host-buffer entry of that layout:
features cross PCIe ONCE in total (every rank uploads only its row block of its column block) and the row groups
complete each other's replicas over NVLink."""
import torch
import torch.distributed as dist

from .. import ops

_groups = {}


def column_group(rr, rc, c):
    """Process group of the Rr ranks that hold column block c (ranks q * Rc + c).  Every rank must call this for
    every c in the same order (torch.distributed.new_group is collective); groups are cached."""
    key = (rr, rc)
    if key not in _groups:
        _groups[key] = [dist.new_group([q * rc + cc for q in range(rr)]) for cc in range(rc)]
    return _groups[key][c]


class GridHostAggregator(object):
    """out_host[N_r, Dl] = aggregate(features) for rank (r, c) of an Rr x Rc grid, features in pinned HOST memory.

    ``submit(x_part_host, out_host, ...)``: ``x_part_host`` is this rank's [rows bounds[r]:bounds[r+1], Dl] piece of the
    column block (``bounds`` = the UPLOAD split of the source rows, normally even blocks so that one in-place
    all-gather completes the replica; it is independent of how destination rows are assigned).  Per step:
    upload the piece into its slot of a [N, Dl] device buffer (own stream), all-gather the slots inside the column
    group (NCCL over NVLink; in place), aggregate, download the output block (own stream).  Two buffer sets: the upload
    of step i+1 overlaps the collective / kernel / download of step i.  ``wait(ticket)`` blocks the host until that
    step's output has landed.  Same buffer contract as ops.HostAggregator."""

    def __init__(self, fwd, n_src, n_dst, dim, device, rr, rc, r, c, bounds, depth=2):
        self.fwd, self.n_src, self.n_dst, self.dim, self.device = fwd, int(n_src), int(n_dst), int(dim), device
        self.rr, self.rc, self.r, self.c = int(rr), int(rc), int(r), int(c)
        self.bounds = [int(b) for b in bounds]
        self.lo, self.hi = self.bounds[self.r], self.bounds[self.r + 1]
        self.group = column_group(self.rr, self.rc, self.c) if self.rr > 1 else None
        sizes = [self.bounds[q + 1] - self.bounds[q] for q in range(self.rr)]
        self.even = len(set(sizes)) == 1
        self.depth = max(1, int(depth))
        self.sets = [None] * self.depth
        self.count = 0
        self.s_in = torch.cuda.Stream(device=device)
        self.s_out = torch.cuda.Stream(device=device)

    def _set(self, i):
        st = self.sets[i]
        if st is None:
            st = {"xd": torch.empty((self.n_src, self.dim), dtype=torch.float32, device=self.device),
                  "od": torch.empty((self.n_dst, self.dim), dtype=torch.float32, device=self.device),
                  "kernel_done": None, "d2h_done": None}
            self.sets[i] = st
        return st

    def _complete_replica(self, xd):
        if self.rr == 1:
            return
        if self.even:
            dist.all_gather_into_tensor(xd, xd[self.lo:self.hi], group=self.group)   # in place: my slot is already there
        else:
            for q in range(self.rr):   # uneven row blocks: one broadcast per block (still each row crosses NVLink once per peer)
                dist.broadcast(xd[self.bounds[q]:self.bounds[q + 1]], src=q * self.rc + self.c, group=self.group)

    def submit(self, x_part_host, out_host, reduce_op="sum", scale_src=None, scale_dst=None):
        assert tuple(x_part_host.shape) == (self.hi - self.lo, self.dim) and x_part_host.is_contiguous()
        assert tuple(out_host.shape) == (self.n_dst, self.dim) and out_host.is_contiguous()
        st = self._set(self.count % self.depth)
        self.count += 1
        main = torch.cuda.current_stream(self.device)
        if st["kernel_done"] is not None:
            self.s_in.wait_event(st["kernel_done"])    # the previous user of this xd has been aggregated
        xd, od = st["xd"], st["od"]
        with torch.cuda.stream(self.s_in):
            xd[self.lo:self.hi].copy_(x_part_host, non_blocking=True)
        ev_in = torch.cuda.Event()
        ev_in.record(self.s_in)
        main.wait_event(ev_in)
        if st["d2h_done"] is not None:
            main.wait_event(st["d2h_done"])            # the previous result has left this od
        self._complete_replica(xd)                     # NCCL, ordered on the current stream
        ops._spmm_raw(self.fwd["indptr"], self.fwd["cols"], xd, self.n_dst, reduce_op, scale_src=scale_src,
                      scale_dst=scale_dst, max_degree=self.fwd.get("max_degree", -1), out=od,
                      packed=ops._packed_of(self.fwd, xd))
        ev_k = torch.cuda.Event()
        ev_k.record(main)
        st["kernel_done"] = ev_k
        self.s_out.wait_event(ev_k)
        with torch.cuda.stream(self.s_out):
            out_host.copy_(od, non_blocking=True)
        done = torch.cuda.Event()
        done.record(self.s_out)
        st["d2h_done"] = done
        return done

    def wait(self, ticket):
        ticket.synchronize()
