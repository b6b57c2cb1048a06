"""ShardedGraph: one rank's shard of a partitioned graph on the sm_100a kernels."""
import numpy as np
import torch
import torch.distributed as dist

from .. import ops
from ..utils.edge_index import EdgeIndex
from .halo import HaloPlan, block_offsets, relabel_by_partition


class ShardedGraph(object):
    """Rank-local dst rows + all their in-edges; source features = [own rows | halo rows].

    ``mode="nccl"``: halo rows are packed by a gather kernel and moved with one NCCL
    all-to-all.  ``mode="p2p"``: every rank maps its peers' feature buffers (CUDA IPC, opened
    under its own device) and a gather kernel pulls the halo rows straight out of the owners'
    HBM over NVLink -- no pack pass, no staging buffer, one launch per peer.

    ``overlap=True`` splits the local CSR by source: edges whose source is owned locally are
    aggregated on the compute stream WHILE the halo rows are in flight on a second stream; the
    halo-source edges are then added on top (PGLB_SPMM_ACCUMULATE) with the row epilogue.
    Measured slower than the plain sequence at 2 GPUs on the power-law bench graph (12.0 vs
    8.8 ms/step: two half-passes double the per-row epilogue work and re-read the output), so
    it is off by default.
    """

    def __init__(self, plan, mode="p2p", overlap=False):
        self.plan = plan
        self.mode = mode
        self.overlap = bool(overlap) and plan.world > 1
        self.rank, self.world = plan.rank, plan.world
        self.n_local, self.n_halo = plan.n_local, plan.n_halo
        self.device = plan.dst_local.device
        # local CSR keyed by local dst; columns index the extended feature buffer
        deg, sv, su, se, ip = ops.csr_build(plan.dst_local, plan.col_local, self.n_local)
        self.index = EdgeIndex.from_index(sorted_v=sv, sorted_u=su, sorted_eid=se, degree=deg,
                                          indptr=ip)
        self._csr = self.index.csr()
        self._csr_loc = self._csr_halo = None
        if self.overlap:
            is_loc = plan.col_local < self.n_local
            for name, m in (("_csr_loc", is_loc), ("_csr_halo", ~is_loc)):
                d2, v2, u2, e2, p2 = ops.csr_build(plan.dst_local[m], plan.col_local[m], self.n_local)
                ix = EdgeIndex.from_index(sorted_v=v2, sorted_u=u2, sorted_eid=e2, degree=d2, indptr=p2)
                setattr(self, name, ix.csr())
        self._buffers = {}
        self._ipc = {}
        self._peers = {}
        self._norm = None
        self._norm_ext = None
        self._halo_off = np.concatenate([[0], np.cumsum(plan.recv_counts)]).astype(np.int64)
        self._pull_idx = [plan.halo_ids[int(self._halo_off[p]):int(self._halo_off[p + 1])] - plan.offsets[p]
                          for p in range(self.world)]
        self._comm_stream = torch.cuda.Stream(device=self.device) if self.world > 1 else None
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_global_edges(cls, edges, num_nodes, world, rank, method="block", part=None,
                          mode="p2p", overlap=False, group=None):
        """edges: [E, 2] int64 CUDA tensor with GLOBAL ids, identical on every rank.
        method "block": contiguous id blocks; "metis": pgl.partition.metis_partition of the
        symmetrised graph on rank 0 (host), broadcast, nodes relabelled so parts are
        contiguous; or pass ``part`` ([N] int64) directly."""
        n = int(num_nodes)
        new_id = None
        if part is None and method == "metis":
            part_t = torch.empty(n, dtype=torch.int64, device=edges.device)
            if rank == 0:
                import pgl_b200 as pgl
                e = edges.cpu().numpy()
                sym = np.concatenate([e, e[:, ::-1]], 0)
                sym = np.unique(sym[sym[:, 0] != sym[:, 1]], axis=0)
                g = pgl.Graph(edges=sym, num_nodes=n)
                part_t.copy_(torch.from_numpy(pgl.partition.metis_partition(g, world)))
            if world > 1:
                dist.broadcast(part_t, 0, group=group)
            part = part_t.cpu().numpy()
        if part is not None:
            if edges.is_cuda:   # stable sort by part + inverse permutation + edge relabel: csrc/localgraph.cu
                nid, off_t = ops.partition_relabel(torch.as_tensor(np.asarray(part), dtype=torch.int64).to(edges.device),
                                                   world)
                new_id, offsets = nid.cpu().numpy(), [int(v) for v in off_t.tolist()]
                edges = ops.map_edges(None, edges, nid)
            else:
                new_id, offsets = relabel_by_partition(part, world)
                nid = torch.from_numpy(new_id).to(edges.device)
                edges = nid[edges]
        else:
            offsets = block_offsets(n, world)
        plan = HaloPlan.build(edges, n, offsets, rank, world, group=group)
        self = cls(plan, mode=mode, overlap=overlap)
        self.new_id = new_id
        return self

    def owned_global_ids(self):
        """ORIGINAL (pre-relabel) global id of every owned row, in local row order (int64 tensor on the
        shard's device): row i of this rank's features / outputs is node owned_global_ids()[i]."""
        lo, hi = self.plan.lo, self.plan.hi
        if getattr(self, "new_id", None) is None:
            return torch.arange(lo, hi, dtype=torch.int64, device=self.device)
        perm = np.empty(len(self.new_id), dtype=np.int64)
        perm[self.new_id] = np.arange(len(self.new_id), dtype=np.int64)   # new id -> original id
        return torch.from_numpy(perm[lo:hi]).to(self.device)

    # ------------------------------------------------------------------ feature buffers
    def features(self, dim):
        """(x_ext [n_local + n_halo, dim], x_local view of its first n_local rows).  Keep node
        features in x_local so that no copy is needed before an exchange."""
        if dim not in self._buffers:
            rows = self.n_local + self.n_halo
            x_ext = None
            if self.mode == "p2p" and self.world > 1:
                x_ext = self._alloc_p2p(rows, dim)
            if x_ext is None:
                x_ext = torch.empty((rows, dim), dtype=torch.float32, device=self.device)
            self._buffers[dim] = x_ext
        x_ext = self._buffers[dim]
        return x_ext, x_ext[: self.n_local]

    def _alloc_p2p(self, rows, dim):
        """IPC-shareable buffer + peer mappings; every rank must succeed, otherwise all ranks
        switch to the NCCL all-to-all transport (a transport choice, not a compute fallback)."""
        import warnings
        ok, buf, peers, err = 1, None, None, ""
        try:
            buf = ops.IpcBuffer(max(rows, 1), dim, self.device)
            handle = buf.handle_bytes()
        except Exception as ex:  # pragma: no cover
            ok, handle, err = 0, b"", repr(ex)
        handles = [None] * self.world
        dist.all_gather_object(handles, (ok, handle), group=self.plan.group)
        if all(h[0] for h in handles):
            try:
                peers = [None if p == self.rank else ops.ipc_open(h[1], self.device)
                         for p, h in enumerate(handles)]
            except Exception as ex:  # pragma: no cover
                ok, err = 0, repr(ex)
        else:
            ok = 0
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=self.plan.group)
        if not all(flags):
            warnings.warn("pgl_b200: CUDA IPC peer mapping unavailable (%s); halo exchange uses "
                          "NCCL all-to-all" % err)
            self.mode = "nccl"
            return None
        self._ipc[dim] = buf
        self._peers[dim] = peers
        return buf.tensor[:rows]

    def _sync_peers(self):
        # a 4-byte all-reduce: orders every rank's stream behind every other rank's prior work
        dist.all_reduce(self._flag, group=self.plan.group)

    def _move_halo(self, x_ext, dim):
        """Enqueue the halo transfer on the CURRENT stream."""
        view = x_ext[: self.n_local]
        if self.mode == "p2p":
            self._sync_peers()  # owners' rows are final (stream ordered)
            peers = self._peers[dim]
            for k in range(1, self.world):
                p = (self.rank + k) % self.world  # stagger the peers
                lo, hi = int(self._halo_off[p]), int(self._halo_off[p + 1])
                if hi > lo:
                    ops.gather_rows_ptr(peers[p], dim, self._pull_idx[p],
                                        x_ext[self.n_local + lo: self.n_local + hi])
            self._sync_peers()  # nobody overwrites rows that are still being read
        else:
            self.plan.exchange(view, x_ext=x_ext, pack=ops.gather_rows)

    def exchange(self, x_local):
        """[x_local | halo rows] for the current feature width (blocking in stream order)."""
        dim = int(x_local.shape[1])
        x_ext, view = self.features(dim)
        if x_local.data_ptr() != view.data_ptr():
            view.copy_(x_local)
        if self.world > 1:
            self._move_halo(x_ext, dim)
        return x_ext

    # ------------------------------------------------------------------ aggregation
    def local_norm(self):
        """clip(indegree, 1)^-0.5 of the owned nodes (all in-edges are local)."""
        if self._norm is None:
            self._norm = ops.degree_norm(self.index.degree).reshape(-1)
        return self._norm

    def ext_norm(self):
        if self._norm_ext is None:
            nl = self.local_norm().reshape(-1, 1).contiguous()
            if self.world == 1:
                self._norm_ext = nl.reshape(-1)
            else:
                ext = self.plan.exchange(nl, pack=ops.gather_rows)
                self._norm_ext = ext.reshape(-1).contiguous()
        return self._norm_ext

    def _agg(self, csr, x_ext, reduce_op, scale_src, scale_dst, out=None, accumulate=False):
        return ops._spmm_raw(csr["indptr"], csr["cols"], x_ext, self.n_local, reduce_op,
                             scale_src=scale_src, scale_dst=scale_dst,
                             max_degree=csr["max_degree"], out=out, accumulate=accumulate,
                             packed=ops._packed_of(csr, x_ext))

    def send_recv(self, x_local, reduce_op="sum", scale_src=None, scale_dst=None):
        """Graph.send_recv on the shard: out rows = owned nodes."""
        if self.overlap and reduce_op == "sum":
            dim = int(x_local.shape[1])
            x_ext, view = self.features(dim)
            if x_local.data_ptr() != view.data_ptr():
                view.copy_(x_local)
            main = torch.cuda.current_stream(self.device)
            self._comm_stream.wait_stream(main)
            with torch.cuda.stream(self._comm_stream):
                self._move_halo(x_ext, dim)
            # local-source edges while the halo rows travel
            out = self._agg(self._csr_loc, x_ext, "sum", scale_src, None)
            main.wait_stream(self._comm_stream)
            return self._agg(self._csr_halo, x_ext, "sum", scale_src, scale_dst, out=out,
                             accumulate=True)
        x_ext = self.exchange(x_local)
        return self._agg(self._csr, x_ext, reduce_op, scale_src, scale_dst)

    def gcn_aggregate(self, x_local, norm_local=None):
        """norm * (A (norm * x)) on the shard (the GCNConv aggregation)."""
        return self.send_recv(x_local, "sum", scale_src=self.ext_norm(), scale_dst=self.local_norm())

    # ------------------------------------------------------------------ reporting
    def stats(self):
        s = self.plan.stats()
        s["mode"] = self.mode
        s["overlap"] = self.overlap
        s["max_in_degree"] = int(self._csr["max_degree"])
        if self._csr_loc is not None:
            s["e_local_src"] = int(self._csr_loc["cols"].shape[0])
            s["e_halo_src"] = int(self._csr_halo["cols"].shape[0])
        return s

    def time_split(self, x_local, norm_local=None, iters=3):
        """Device-timed pieces on this rank (ms): halo exchange, full local aggregation and, when
        the source-split CSRs exist, the two half passes of the overlapped variant."""
        ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
        self.gcn_aggregate(x_local)
        torch.cuda.synchronize()
        res = {"rank": self.rank, "exchange_ms": 0.0, "aggregate_ms": 0.0}
        if self._csr_loc is not None:
            res["loc_pass_ms"] = res["halo_pass_ms"] = 0.0
        for _ in range(iters):
            if self.world > 1:
                dist.barrier(group=self.plan.group)
            a, b, c, d, e = ev(), ev(), ev(), ev(), ev()
            a.record()
            x_ext = self.exchange(x_local)
            b.record()
            self._agg(self._csr, x_ext, "sum", self.ext_norm(), self.local_norm())
            c.record()
            if self._csr_loc is not None:
                out = self._agg(self._csr_loc, x_ext, "sum", self.ext_norm(), None)
                d.record()
                self._agg(self._csr_halo, x_ext, "sum", self.ext_norm(), self.local_norm(), out=out,
                          accumulate=True)
                e.record()
            torch.cuda.synchronize()
            res["exchange_ms"] += a.elapsed_time(b) / iters
            res["aggregate_ms"] += b.elapsed_time(c) / iters
            if self._csr_loc is not None:
                res["loc_pass_ms"] += c.elapsed_time(d) / iters
                res["halo_pass_ms"] += d.elapsed_time(e) / iters
        return res
