"""Column-sharded full-batch message passing (SURVEY 8e "alternative to measure"; the reference
shards features by column in examples/.../dist_feat.py:31-41).

Every rank holds the WHOLE graph (dst-CSR) and a [N, D/R] column slice of the features.  A copy-message
aggregation is independent per column, so it needs no exchange at all; one all-to-all re-shards
[N, D/R] -> [N/R, D] for the dense transform of a layer (which needs whole rows) and one brings the
result back.  Exchange volume per GPU is N*D*4/R bytes per direction whatever the graph looks like
(0.64 GB at cfg5 / 8 GPUs), against 2.1-2.5 GB of halo rows per rank for the row partition of a power-law
graph.  The price is narrow rows (D/R floats): they need the narrow-row streaming
kernel (csrc/spmm_narrow2.inl, the default for rows of <= 64 floats).

Status: the re-shard logic is covered by a world-2 gloo test on CPU and runs under NCCL in bench.py's full_layer
leg on every grid; the narrow-row aggregation was validated and measured on hardware in round 2 (DESIGN.md
sections 4.2b and 5.1 -- it is why the bench default shards destination rows, not columns).
"""
import torch
import torch.distributed as dist

from .halo import block_offsets


class ColumnShardedGraph(object):
    def __init__(self, graph, dim, world, rank, group=None):
        assert dim % world == 0, "the feature width must be divisible by the number of ranks"
        self.graph = graph
        self.dim, self.world, self.rank, self.group = int(dim), int(world), int(rank), group
        self.d_local = self.dim // self.world
        self.num_nodes = graph._n if graph is not None else None
        self._offsets = None

    # ---- layout helpers (device agnostic; exercised by the gloo test) ---------------------------
    def column_range(self):
        return self.rank * self.d_local, (self.rank + 1) * self.d_local

    def offsets(self, num_nodes):
        return block_offsets(num_nodes, self.world)

    def slice_columns(self, x_full):
        lo, hi = self.column_range()
        return x_full[:, lo:hi].contiguous()

    def to_rows(self, x_cols):
        """[N, D/R] (my columns of every row) -> [N_r, D] (every column of my row block)."""
        n = int(x_cols.shape[0])
        off = self.offsets(n)
        if self.world == 1:
            return x_cols
        sizes = [off[r + 1] - off[r] for r in range(self.world)]
        mine = sizes[self.rank]
        recv = torch.empty((mine * self.world, self.d_local), dtype=x_cols.dtype, device=x_cols.device)
        dist.all_to_all_single(recv, x_cols.contiguous(), output_split_sizes=[mine] * self.world,
                               input_split_sizes=sizes, group=self.group)
        # chunk q = my rows, rank q's columns  ->  interleave the chunks along the column axis
        return recv.reshape(self.world, mine, self.d_local).permute(1, 0, 2).reshape(mine, self.dim)

    def to_cols(self, x_rows, num_nodes):
        """[N_r, D] -> [N, D/R]: inverse of to_rows."""
        if self.world == 1:
            return x_rows
        off = self.offsets(num_nodes)
        sizes = [off[r + 1] - off[r] for r in range(self.world)]
        mine = sizes[self.rank]
        send = x_rows.reshape(mine, self.world, self.d_local).permute(1, 0, 2).contiguous()
        recv = torch.empty((int(num_nodes), self.d_local), dtype=x_rows.dtype, device=x_rows.device)
        dist.all_to_all_single(recv, send.reshape(mine * self.world, self.d_local),
                               output_split_sizes=sizes, input_split_sizes=[mine] * self.world,
                               group=self.group)
        return recv

    # ---- CUDA path ---------------------------------------------------------------------------------
    def send_recv(self, x_cols, reduce_op="sum", scale_src=None, scale_dst=None):
        """Aggregation of my column slice over the whole graph: no communication."""
        return self.graph._send_u_recv(x_cols, reduce_op, None, scale_src=scale_src, scale_dst=scale_dst)

    def gcn_aggregate(self, x_cols, norm):
        nv = norm.reshape(-1)
        return self.send_recv(x_cols, "sum", scale_src=nv, scale_dst=nv)

    def gcn_layer(self, x_cols, norm, weight, bias=None, activation=None, aggregate=None):
        """One GCNConv layer (reference pgl/nn/conv.py:218-254, aggregate-then-transform order) on the
        column-sharded layout: aggregate my columns (no exchange) -> all-to-all to whole rows ->
        dense transform of my row block -> all-to-all back to columns.  ``weight`` is the full
        [D, D_out] matrix (replicated; D_out divisible by the world size).  Returns [N, D_out / R].
        ``aggregate(x_cols, norm)`` can be injected (the gloo test passes the oracle)."""
        n = int(x_cols.shape[0])
        agg = (aggregate or self.gcn_aggregate)(x_cols, norm)
        rows = self.to_rows(agg)
        out = rows @ weight
        if bias is not None:
            out = out + bias
        if activation is not None:
            out = activation(out)
        d_out = int(weight.shape[1])
        assert d_out % self.world == 0
        back = ColumnShardedGraph(None, d_out, self.world, self.rank, self.group)
        return back.to_cols(out, n)
