"""Batched shuffling buffers over torch tensors (state machine of
petastorm/reader_impl/pytorch_shuffling_buffer.py:85-279).  ``add_many`` takes a list of column tensors with a common
first dimension, ``retrieve`` returns one batch (list of tensors).  On CUDA tensors the random batch gather and the
compaction of the survivors run through the hand-written row-gather kernel (K14 = K13 ``pst_gather_rows``); the
permutation itself comes from ``torch.randperm`` on the buffer's device, exactly like upstream (global torch RNG)."""
import abc

import torch


def _take(t, idx):
    if t.is_cuda:
        from petastorm_b200 import device_ops
        return device_ops.gather_rows(t.contiguous(), idx)
    return t[idx]


class BatchedShufflingBufferBase(abc.ABC):
    def __init__(self, batch_size=1):
        self._keys = None
        self.batch_size = batch_size

    def add_many(self, items):
        items = [torch.as_tensor(v) for v in items]
        return self._add_many(items)

    @abc.abstractmethod
    def _add_many(self, items):
        pass

    @abc.abstractmethod
    def retrieve(self):
        pass

    @abc.abstractmethod
    def can_add(self):
        pass

    @abc.abstractmethod
    def can_retrieve(self):
        pass

    @property
    @abc.abstractmethod
    def size(self):
        pass

    @abc.abstractmethod
    def finish(self):
        pass


class BatchedNoopShufflingBuffer(BatchedShufflingBufferBase):
    """FIFO: concatenates the leftover of the previous row-group with the next one and slices batches off the front."""

    def __init__(self, batch_size=1):
        super(BatchedNoopShufflingBuffer, self).__init__(batch_size=batch_size)
        self._size = 0
        self._buffer = []
        self._done_adding = False
        self._start = 0

    def _add_many(self, items):
        n_new = len(items[0])
        if not self._buffer or self._size == 0:
            self._buffer = list(items)
        else:
            self._buffer = [torch.cat([old[self._start:], new], 0) for old, new in zip(self._buffer, items)]
        self._size += n_new
        self._start = 0

    def retrieve(self):
        take = min(self._size, self.batch_size)
        batch = [v[self._start:self._start + take] for v in self._buffer]
        self._start += take
        self._size -= take
        return batch

    def can_retrieve(self):
        return self._size >= self.batch_size if not self._done_adding else self._size > 0

    def can_add(self):
        return True

    @property
    def size(self):
        return self._size

    def finish(self):
        self._done_adding = True


class BatchedRandomShufflingBuffer(BatchedShufflingBufferBase):
    """Uniform sampling without replacement from a bounded buffer: a permutation of the live rows is drawn lazily and
    consumed ``batch_size`` indices at a time; the next ``add_many`` first compacts the unconsumed rows to the front
    (their permuted order) and appends the new rows, doubling the allocation when needed."""

    def __init__(self, shuffling_buffer_capacity, min_after_retrieve, extra_capacity=1000, batch_size=1):
        super(BatchedRandomShufflingBuffer, self).__init__(batch_size=batch_size)
        self._extra_capacity = extra_capacity
        self._items = None
        self._capacity = shuffling_buffer_capacity
        self._min_after_retrieve = min_after_retrieve
        self._size = 0
        self._done_adding = False
        self._perm = None
        self._head = 0

    def _add_many(self, items):
        if self._done_adding:
            raise RuntimeError('Can not call add_many after done_adding() was called.')
        if not self.can_add():
            raise RuntimeError('Can not enqueue. Check the return value of "can_enqueue()" to check if more '
                               'items can be added.')
        expected = self._size + len(items[0])
        limit = self._capacity + self._extra_capacity
        if expected > limit:
            raise RuntimeError('Attempt to enqueue more elements than the capacity allows. '
                               'Current size: {}, new size {}, maximum allowed: {}'.format(self._size, expected, limit))
        new_capacity = max(self._capacity, 1)
        while new_capacity < expected:
            new_capacity *= 2
        if self._items is None:
            self._items = [torch.empty((new_capacity,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
                           for v in items]
        if self._head > 0:
            survivors = self._perm[self._head:]
            for k, v in enumerate(self._items):
                v[:self._size] = _take(v, survivors)
        self._perm = None
        self._head = 0
        if new_capacity > self._items[0].shape[0]:
            for k, v in enumerate(self._items):
                grown = torch.empty((new_capacity,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
                grown[:self._size] = v[:self._size]
                self._items[k] = grown
        for k, v in enumerate(items):
            self._items[k][self._size:expected] = v
        self._size = expected

    def retrieve(self):
        if not self._done_adding and not self.can_retrieve():
            raise RuntimeError('Can not dequeue. Check the return value of "can_dequeue()" to check if any '
                               'items are available.')
        take = min(self.batch_size, self._size)
        if self._perm is None:
            self._head = 0
            self._perm = torch.randperm(int(self._size), device=self._items[0].device)
        idx = self._perm[self._head:self._head + take]
        self._head += take
        sample = [_take(v, idx) for v in self._items]
        self._size -= take
        return sample

    def can_add(self):
        return self._size < self._capacity and not self._done_adding

    def can_retrieve(self):
        return self._size >= self._min_after_retrieve + self.batch_size - 1 or (self._done_adding and self._size > 0)

    @property
    def size(self):
        return self._size

    def finish(self):
        self._done_adding = True
