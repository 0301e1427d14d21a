"""Pool protocol of the readers (petastorm/workers_pool/__init__.py:16-26)."""


class EmptyResultError(RuntimeError):
    """No result is queued and none will come unless more work is ventilated."""


class TimeoutWaitingForResultError(RuntimeError):
    """A timeout elapsed while waiting for a result."""


class VentilatedItemProcessedMessage(object):
    """Marker: a worker finished one ventilated item."""
