"""GPU neighbour sampling (mirror of reference pgl/sampling/sage.py:130-155): csrc/sampling.cu, validated on
hardware in round 2 (tests/test_gpu_sampling.py)."""
from .custom import subgraph  # noqa: F401
from .sage import NeighborSampler, reindex_graph, sample_neighbors  # noqa: F401
