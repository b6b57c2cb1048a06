"""GPU neighbour sampling (mirror of reference pgl/sampling/sage.py:130-155).  EXPERIMENTAL: the
kernels behind it were written after round 1's GPU budget was spent and have not run on hardware
yet."""
from .sage import NeighborSampler, reindex_graph, sample_neighbors  # noqa: F401
