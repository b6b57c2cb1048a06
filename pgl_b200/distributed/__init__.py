"""Multi-GPU full-batch message passing: 1-D node partition + halo exchange.

The reference's multi-GPU mode (``DistGPUGraph``, pgl/graph.py:1410-1553) replicates every
node feature on every GPU and all-reduces the whole [N, D] output after each aggregation
(5.1 GB per layer per GPU at 10M x 128).  Here every rank owns a contiguous block of nodes
(after an optional METIS relabelling, the convention of apps/GNNAutoScale/graph_partition.py:
94-101), all in-edges of its nodes, and the features of its own nodes only; before an
aggregation it fetches just the distinct remote source rows it needs ("halo").

* ``HaloPlan``      pure index logic (torch ops, device agnostic) + the exchange; covered by
                    world-size-2 gloo tests on CPU.
* ``ShardedGraph``  the CUDA product class: local CSR on the sm_100a kernels, halo rows moved
                    either by NCCL all-to-all (``mode="nccl"``) or pulled straight out of the
                    peers' HBM over NVLink by a gather kernel on IPC-mapped pointers
                    (``mode="p2p"``).
"""
from .halo import HaloPlan, block_offsets, relabel_by_partition  # noqa: F401
from .colshard import ColumnShardedGraph  # noqa: F401
from .sharded import ShardedGraph  # noqa: F401
from .grid import GridHostAggregator, column_group  # noqa: F401
