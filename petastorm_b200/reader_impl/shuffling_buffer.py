"""Row-at-a-time shuffling buffers (protocol and state machine of petastorm/reader_impl/shuffling_buffer.py:23-180).

``can_add`` while ``size < capacity`` and not finished; ``can_retrieve`` once ``size >= min_after_retrieve`` or, after
``finish()``, while anything is left; ``retrieve`` draws a uniform index from the *global* ``np.random`` state and
swaps the last item into the hole.  Not thread safe (single consumer), like upstream.
"""
import abc
from collections import deque

import numpy as np


class ShufflingBufferBase(abc.ABC):
    @abc.abstractmethod
    def add_many(self, items):
        """Append items."""

    @abc.abstractmethod
    def retrieve(self):
        """Remove and return one item."""

    @abc.abstractmethod
    def can_add(self):
        """May add_many be called now?"""

    @abc.abstractmethod
    def can_retrieve(self):
        """May retrieve be called now?"""

    @property
    @abc.abstractmethod
    def size(self):
        """Items currently stored."""

    @abc.abstractmethod
    def finish(self):
        """No more items will be added: allow draining below min_after_retrieve."""


class NoopShufflingBuffer(ShufflingBufferBase):
    """FIFO pass-through."""

    def __init__(self):
        self.store = deque()

    def add_many(self, items):
        self.store.extend(items)

    def retrieve(self):
        return self.store.popleft()

    def can_retrieve(self):
        return len(self.store) > 0

    def can_add(self):
        return True

    @property
    def size(self):
        return len(self.store)

    def finish(self):
        pass


class RandomShufflingBuffer(ShufflingBufferBase):
    def __init__(self, shuffling_buffer_capacity, min_after_retrieve, extra_capacity=1000):
        self._capacity = shuffling_buffer_capacity
        self._min_after_retrieve = min_after_retrieve
        self._extra_capacity = extra_capacity
        self._items = []
        self._done_adding = False

    def add_many(self, items):
        if self._done_adding:
            raise RuntimeError('Can not call add_many after done_adding() was called.')
        if not self.can_add():
            raise RuntimeError('Can not enqueue. Check the return value of "can_enqueue()" to check if more '
                               'items can be added.')
        expected = len(self._items) + len(items)
        limit = self._capacity + self._extra_capacity
        if expected > limit:
            raise RuntimeError('Attempt to enqueue more elements than the capacity allows. '
                               'Current size: {}, new size {}, maximum allowed: {}'.format(len(self._items), expected, limit))
        self._items.extend(items)

    def retrieve(self):
        if not self._done_adding and not self.can_retrieve():
            raise RuntimeError('Can not dequeue. Check the return value of "can_dequeue()" to check if any '
                               'items are available.')
        i = np.random.randint(0, len(self._items))
        value = self._items[i]
        last = self._items.pop()
        if i < len(self._items):
            self._items[i] = last
        return value

    def can_add(self):
        return len(self._items) < self._capacity and not self._done_adding

    def can_retrieve(self):
        return len(self._items) >= self._min_after_retrieve or (self._done_adding and len(self._items) > 0)

    @property
    def size(self):
        return len(self._items)

    def finish(self):
        self._done_adding = True
