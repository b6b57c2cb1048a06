"""HaloPlan: which rows a rank owns, which remote rows it needs, and how they move.

On CUDA tensors the plan is built by the kernels of csrc/localgraph.cu (``ops.halo_plan``: flag arrays + prefix
sums, no sort, no hash; ``ops.partition_relabel`` / ``ops.map_edges`` for the relabelling); on CPU tensors (the gloo
tests of the host logic) by the torch index arithmetic below, which the GPU tests hold the kernels against bit for bit.
``torch.distributed`` collectives (NCCL on GPUs, gloo in the CPU tests) move counts and id lists.  No feature math.
"""
import numpy as np
import torch
import torch.distributed as dist


def block_offsets(num_nodes, world):
    """Contiguous equal blocks: offsets[r] .. offsets[r+1] is rank r's node range."""
    n = int(num_nodes)
    cs = -(-n // world)
    return [min(r * cs, n) for r in range(world)] + [n]


def relabel_by_partition(part, world):
    """part[N] (e.g. from pgl.partition.metis_partition) -> (new_id[N], offsets[world+1]) such
    that every part is a contiguous id range; stable inside a part (the permutation /
    offsets convention of reference apps/GNNAutoScale/graph_partition.py:94-101)."""
    part = np.asarray(part, dtype=np.int64)
    perm = np.argsort(part, kind="stable")
    new_id = np.empty_like(perm)
    new_id[perm] = np.arange(len(perm), dtype=np.int64)
    counts = np.bincount(part, minlength=world)[:world]
    offsets = [0] + np.cumsum(counts).tolist()
    return new_id, offsets


class HaloPlan(object):
    """Rank-local view of a globally numbered graph whose parts are contiguous id ranges.

    Attributes (all torch tensors on the device of ``edges``):
        dst_local [E_r]   destination row in [0, n_local)
        col_local [E_r]   source row in the extended buffer [own rows | halo rows]
        eid       [E_r]   global edge id of every local edge (ascending)
        halo_ids  [H]     sorted distinct remote source ids == rows n_local.. of the buffer
        recv_counts [R]   rows received from each peer (python ints)
        send_counts [R]   rows sent to each peer
        send_idx  [S]     local row of every row to send, grouped by destination peer
    """

    def __init__(self):
        pass

    @classmethod
    def build(cls, edges, num_nodes, offsets, rank, world, group=None, force_torch=False):
        self = cls()
        self.rank, self.world, self.group = int(rank), int(world), group
        self.offsets = [int(o) for o in offsets]
        lo, hi = self.offsets[rank], self.offsets[rank + 1]
        self.lo, self.hi = lo, hi
        self.n_local = hi - lo
        dev = edges.device
        if edges.is_cuda and not force_torch:
            from .. import ops
            self.eid, self.dst_local, self.col_local, halo_ids, recv_counts = ops.halo_plan(
                edges, num_nodes, lo, hi, self.offsets)
            self.halo_ids = halo_ids
            self.n_halo = int(halo_ids.shape[0])
        else:
            src_all, dst_all = edges[:, 0], edges[:, 1]
            mine = (dst_all >= lo) & (dst_all < hi)
            eid = torch.nonzero(mine, as_tuple=False).reshape(-1)
            src = src_all.index_select(0, eid)
            self.eid = eid
            self.dst_local = dst_all.index_select(0, eid) - lo
            remote = (src < lo) | (src >= hi)
            halo_ids = torch.unique(src[remote])  # sorted => grouped by owner
            self.halo_ids = halo_ids
            self.n_halo = int(halo_ids.shape[0])
            pos = torch.searchsorted(halo_ids, src) if self.n_halo else torch.zeros_like(src)
            self.col_local = torch.where(remote, pos + self.n_local, src - lo)
            off_t = torch.tensor(self.offsets, dtype=torch.int64, device=dev)
            bounds = torch.searchsorted(halo_ids, off_t) if self.n_halo else torch.zeros_like(off_t)
            recv_counts = (bounds[1:] - bounds[:-1]).to(torch.int64)
        send_counts = torch.empty_like(recv_counts)
        if world > 1:
            dist.all_to_all_single(send_counts, recv_counts, group=group)
        else:
            send_counts.copy_(recv_counts)
        self.recv_counts = [int(v) for v in recv_counts.tolist()]
        self.send_counts = [int(v) for v in send_counts.tolist()]
        want = torch.empty(sum(self.send_counts), dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_to_all_single(want, halo_ids, output_split_sizes=self.send_counts,
                                   input_split_sizes=self.recv_counts, group=group)
        self.send_idx = want - lo
        assert self.recv_counts[rank] == 0 and self.send_counts[rank] == 0
        return self

    def exchange(self, x_local, x_ext=None, pack=None):
        """x_ext = [x_local | halo rows] via pack + all-to-all.  ``pack(x, idx)`` gathers the
        rows to send (the CUDA product passes the gather kernel; the CPU tests pass indexing)."""
        d = tuple(x_local.shape[1:])
        if x_ext is None:
            x_ext = torch.empty((self.n_local + self.n_halo,) + d, dtype=x_local.dtype,
                                device=x_local.device)
        if x_ext.data_ptr() != x_local.data_ptr():
            x_ext[: self.n_local].copy_(x_local)
        if self.world == 1:
            return x_ext
        sendbuf = pack(x_local, self.send_idx)
        dist.all_to_all_single(x_ext[self.n_local:], sendbuf, output_split_sizes=self.recv_counts,
                               input_split_sizes=self.send_counts, group=self.group)
        return x_ext

    def stats(self):
        return {"rank": self.rank, "n_local": self.n_local, "e_local": int(self.eid.shape[0]),
                "halo_rows": self.n_halo, "send_rows": int(self.send_idx.shape[0])}
