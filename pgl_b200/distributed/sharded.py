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
    all-to-all.  ``mode="p2p"``: every rank maps its peers' feature buffers (CUDA IPC) and a
    gather kernel pulls the halo rows straight out of the owners' HBM over NVLink -- no pack
    pass, no staging buffer, one launch per peer, overlappable with the interior aggregation.
    """

    def __init__(self, plan, mode="nccl"):
        self.plan = plan
        self.mode = mode
        self.rank, self.world = plan.rank, plan.world
        self.n_local, self.n_halo = plan.n_local, plan.n_halo
        self.device = plan.dst_local.device
        # local CSR keyed by local dst; columns index the extended feature buffer
        deg, sv, su, se, ip = ops.csr_build(plan.dst_local, plan.col_local, self.n_local)
        self.index = EdgeIndex.from_index(sorted_v=sv, sorted_u=su, sorted_eid=se, degree=deg,
                                          indptr=ip)
        self._csr = self.index.csr()
        self._buffers = {}
        self._peers = {}
        self._norm = None
        self._norm_ext = None
        self._halo_off = np.concatenate([[0], np.cumsum(plan.recv_counts)]).astype(np.int64)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_global_edges(cls, edges, num_nodes, world, rank, method="block", part=None,
                          mode="nccl", group=None):
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
            new_id, offsets = relabel_by_partition(part, world)
            nid = torch.from_numpy(new_id).to(edges.device)
            edges = nid[edges]
        else:
            offsets = block_offsets(n, world)
        plan = HaloPlan.build(edges, n, offsets, rank, world, group=group)
        self = cls(plan, mode=mode)
        self.new_id = new_id
        return self

    # ------------------------------------------------------------------ feature buffers
    def features(self, dim):
        """(x_ext [n_local + n_halo, dim], x_local view of its first n_local rows).  Keep node
        features in x_local so that no copy is needed before an exchange."""
        if dim not in self._buffers:
            x_ext = torch.empty((self.n_local + self.n_halo, dim), dtype=torch.float32,
                                device=self.device)
            self._buffers[dim] = x_ext
            if self.mode == "p2p" and self.world > 1:
                self._map_peers(dim, x_ext)
        x_ext = self._buffers[dim]
        return x_ext, x_ext[: self.n_local]

    def _map_peers(self, dim, x_ext):
        from torch.multiprocessing.reductions import reduce_tensor
        fn, args = reduce_tensor(x_ext)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (fn, args), group=self.plan.group)
        peers = []
        for p, (f, a) in enumerate(gathered):
            peers.append(None if p == self.rank else f(*a))
        self._peers[dim] = peers
        # lists of rows to pull from every peer, in that peer's local numbering
        self._pull_idx = []
        for p in range(self.world):
            lo, hi = int(self._halo_off[p]), int(self._halo_off[p + 1])
            self._pull_idx.append(self.plan.halo_ids[lo:hi] - self.plan.offsets[p])
        dist.barrier(group=self.plan.group)

    def exchange(self, x_local):
        """[x_local | halo rows] for the current feature width."""
        dim = int(x_local.shape[1])
        x_ext, view = self.features(dim)
        if x_local.data_ptr() != view.data_ptr():
            view.copy_(x_local)
        if self.world == 1:
            return x_ext
        if self.mode == "p2p":
            dist.barrier(group=self.plan.group)  # owners' rows are final (stream ordered)
            peers = self._peers[dim]
            for k in range(1, self.world):
                p = (self.rank + k) % self.world  # stagger the peers
                lo, hi = int(self._halo_off[p]), int(self._halo_off[p + 1])
                if hi > lo:
                    ops.gather_rows(peers[p][: self.plan.offsets[p + 1] - self.plan.offsets[p]],
                                    self._pull_idx[p], out=x_ext[self.n_local + lo: self.n_local + hi])
            dist.barrier(group=self.plan.group)  # nobody overwrites rows that are still being read
            return x_ext
        return self.plan.exchange(view, x_ext=x_ext, pack=ops.gather_rows)

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

    def send_recv(self, x_local, reduce_op="sum", scale_src=None, scale_dst=None):
        """Graph.send_recv on the shard: out rows = owned nodes."""
        x_ext = self.exchange(x_local)
        return ops._spmm_raw(self._csr["indptr"], self._csr["cols"], x_ext, self.n_local, reduce_op,
                             scale_src=scale_src, scale_dst=scale_dst,
                             max_degree=self._csr["max_degree"])

    def gcn_aggregate(self, x_local, norm_local=None):
        """norm * (A (norm * x)) on the shard (the GCNConv aggregation)."""
        return self.send_recv(x_local, "sum", scale_src=self.ext_norm(), scale_dst=self.local_norm())

    # ------------------------------------------------------------------ reporting
    def stats(self):
        s = self.plan.stats()
        s["mode"] = self.mode
        s["max_in_degree"] = int(self._csr["max_degree"])
        return s

    def time_split(self, x_local, norm_local=None, iters=3):
        """Device-timed exchange vs aggregation on this rank (ms)."""
        ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
        self.gcn_aggregate(x_local)
        torch.cuda.synchronize()
        a, b, c = ev(), ev(), ev()
        te = ta = 0.0
        for _ in range(iters):
            if self.world > 1:
                dist.barrier(group=self.plan.group)
            a.record()
            x_ext = self.exchange(x_local)
            b.record()
            ops._spmm_raw(self._csr["indptr"], self._csr["cols"], x_ext, self.n_local, "sum",
                          scale_src=self.ext_norm(), scale_dst=self.local_norm(),
                          max_degree=self._csr["max_degree"])
            c.record()
            torch.cuda.synchronize()
            te += a.elapsed_time(b)
            ta += b.elapsed_time(c)
        return {"rank": self.rank, "exchange_ms": te / iters, "aggregate_ms": ta / iters}
