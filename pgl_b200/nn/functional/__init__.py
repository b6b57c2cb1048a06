from .graph_op import *  # noqa: F401,F403
