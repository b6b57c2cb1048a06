"""Message: the object handed to the user's reduce function by Graph.recv
(mirror of reference pgl/message.py:19-173)."""
from . import math, ops


class Message(object):
    """Args:
        msg: the (lazily row-gathered) message dict produced by ``Graph.send``.
        segment_ids: dense, non-decreasing id of the receiving node per (sorted) edge.
    """

    def __init__(self, msg, segment_ids):
        self._segment_ids = segment_ids
        self._msg = msg

    def reduce(self, msg, pool_type="sum"):
        """reference pgl/message.py:34-53."""
        return math.segment_pool(msg, self._segment_ids, pool_type=pool_type)

    def reduce_sum(self, msg):
        """reference pgl/message.py:55-66."""
        return math.segment_sum(msg, self._segment_ids)

    def reduce_mean(self, msg):
        """reference pgl/message.py:68-79."""
        return math.segment_mean(msg, self._segment_ids)

    def reduce_max(self, msg):
        """reference pgl/message.py:81-92."""
        return math.segment_max(msg, self._segment_ids)

    def reduce_min(self, msg):
        """reference pgl/message.py:94-105."""
        return math.segment_min(msg, self._segment_ids)

    def edge_expand(self, msg):
        """reference pgl/message.py:107-157: inverse of reduce (gather by segment id)."""
        return ops.gather_rows(msg, self._segment_ids)

    def reduce_softmax(self, msg):
        """reference pgl/message.py:159-170."""
        return math.segment_softmax(msg, self._segment_ids)

    def __getitem__(self, key):
        return self._msg[key]
