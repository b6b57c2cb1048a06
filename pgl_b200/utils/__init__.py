from . import helper, op, edge_index, relabel  # noqa: F401
